// lacosmic.cu -- LACosmic cosmic-ray detection, astroscrappy.detect_cosmics 1.0.8 semantics
// (sepmed=False, fsmode='median', cleantype='medmask', gain=1, pssl=0, satlevel=inf); the
// reference calls it at blackbox.py:4323-4332.  The CPU statement of the same algorithm, which
// these kernels must match bit for bit, is oracle/csrc/bbo.c (bbo_detect_cosmics).
//
// Per iteration (all float32, one IEEE operation per step, library built with -fmad=false):
//   stage1:  L+ = rebin(max(0, laplace(subsample2(clean))))   (fused: the 2x image never exists)
//            m5 = med5(clean) floored at 1e-5; noise = sqrt(m5 + rn^2); s = L+ / (2 noise)
//            f3 = med3(clean)
//   stage2:  sp = s - med5(s);  f = max((f3 - med7(f3)) / noise, 0.01)
//            flags: c0 = sp > sigclip & good & sp/f > objlim;  b1 = good & sp > sigclip;
//                   b2 = good & sp > sigcliplow
//   grow:    c1 = dilate3(c0) & b1;  c2 = dilate3(c1) & b2;  crmask |= c2;  count += c2
//   clean:   crmask pixels of the interior get the lower median of the 5x5 neighbours that
//            are neither masked nor cosmic, else the background level
// Median filters copy their input in the 1/2/3-pixel frame; dilate3 copies its input in the
// 1-pixel frame; the Laplacian drops neighbours outside the image.
//
// The iteration loop is enqueued up front; a device-side flag turns the kernels of later
// iterations into no-ops once an iteration finds nothing (the reference's `break`), so no
// host synchronisation is needed.
#include "lacosmic_common.cuh"

template <int K>
__global__ void __launch_bounds__(256)
medfilt_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < W && y < H) out[(size_t)y * W + x] = median_at<K>(in, H, W, y, x);
}

__global__ void __launch_bounds__(256)
laplace_plus_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < W && y < H) out[(size_t)y * W + x] = laplace_plus_at(in, H, W, y, x);
}

// --------------------------------------------------------------------------------------------
// iteration kernels
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lac_stage1_kernel(const float *__restrict__ clean, float *__restrict__ s, float *__restrict__ noise,
                  float *__restrict__ f3, int H, int W, LacParams prm, const long long *__restrict__ info)
{
    if (!info[INFO_ACTIVE]) return;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    float m5 = median_at<5>(clean, H, W, y, x);
    if (m5 < 0.00001f) m5 = 0.00001f;
    float nz = m5 + lac_rn2(prm);
    nz = sqrtf(nz);
    const float lp = laplace_plus_at(clean, H, W, y, x);
    const float den = 2.0f * nz;
    s[i] = lp / den;
    noise[i] = nz;
    f3[i] = median_at<3>(clean, H, W, y, x);
}

__global__ void __launch_bounds__(256)
lac_stage2_kernel(const float *__restrict__ s, const float *__restrict__ noise, const float *__restrict__ f3,
                  const uint8_t *__restrict__ inmask, uint8_t *__restrict__ flags, int H, int W, LacParams prm,
                  const long long *__restrict__ info, float *__restrict__ dump_sp, float *__restrict__ dump_f)
{
    if (!info[INFO_ACTIVE]) return;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * W + x;
    const float sp = s[i] - median_at<5>(s, H, W, y, x);
    float f = f3[i] - median_at<7>(f3, H, W, y, x);
    f = f / noise[i];
    if (f < 0.01f) f = 0.01f;
    const bool good = inmask ? (inmask[i] == 0) : true;
    const float ratio = sp / f;
    uint8_t fl = 0;
    if (good && sp > prm.sigclip) { fl |= 2; if (ratio > prm.objlim) fl |= 1; }
    if (good && sp > prm.sigcliplow) fl |= 4;
    flags[i] = fl;
    if (dump_sp) dump_sp[i] = sp;
    if (dump_f) dump_f[i] = f;
}

__device__ __forceinline__ bool lac_c1(const uint8_t *__restrict__ flags, int H, int W, int y, int x)
{
    const uint8_t f = flags[(size_t)y * W + x];
    if (!(f & 2)) return false;
    if (y == 0 || y == H - 1 || x == 0 || x == W - 1) return (f & 1) != 0;   // frame copies c0
    bool any = false;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) any |= (flags[(size_t)(y + dy) * W + (x + dx)] & 1) != 0;
    return any;
}

__global__ void __launch_bounds__(256)
lac_grow_kernel(const uint8_t *__restrict__ flags, uint8_t *__restrict__ crmask, int H, int W, int iter,
                long long *__restrict__ info)
{
    if (!info[INFO_ACTIVE]) return;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    int c2 = 0;
    if (x < W && y < H) {
        const size_t i = (size_t)y * W + x;
        if (flags[i] & 4) {
            if (y == 0 || y == H - 1 || x == 0 || x == W - 1) {
                c2 = lac_c1(flags, H, W, y, x);
            } else {
                bool any = false;
                for (int dy = -1; dy <= 1 && !any; dy++)
                    for (int dx = -1; dx <= 1 && !any; dx++) any = lac_c1(flags, H, W, y + dy, x + dx);
                c2 = any;
            }
            if (c2) crmask[i] = 1;
        }
    }
    const int tot = __syncthreads_count(c2);
    if (threadIdx.x == 0 && tot) atomicAdd((unsigned long long *)&info[INFO_NCR + iter], (unsigned long long)tot);
}

__global__ void lac_init_kernel(long long *info, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) info[i] = (i == INFO_ACTIVE) ? 1 : 0;
}

// runs after grow: decides whether this iteration cleans and whether later ones run
__global__ void lac_control_kernel(long long *info, int iter)
{
    if (!info[INFO_ACTIVE]) return;
    info[INFO_ITERS] = iter + 1;
    if (info[INFO_NCR + iter] == 0) info[INFO_ACTIVE] = 0;
}

__global__ void __launch_bounds__(256)
lac_clean_kernel(float *clean, const uint8_t *__restrict__ crmask, const uint8_t *__restrict__ inmask,
                 int H, int W, const float *__restrict__ background, const long long *__restrict__ info)
{
    if (!info[INFO_ACTIVE]) return;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < 2 || x >= W - 2 || y < 2 || y >= H - 2) return;
    const size_t i = (size_t)y * W + x;
    if (!crmask[i]) return;
    float v[25];
    int n = 0;
    for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
            const size_t j = (size_t)(y + dy) * W + (x + dx);
            const bool bad = crmask[j] || (inmask && inmask[j]);
            if (!bad) v[n++] = clean[j];
        }
    float r = *background;
    if (n > 0) {
        for (int a = 1; a < n; a++) {                     // insertion sort (few pixels, tiny n)
            const float key = v[a];
            int b = a - 1;
            while (b >= 0 && v[b] > key) { v[b + 1] = v[b]; b--; }
            v[b + 1] = key;
        }
        r = v[(n - 1) / 2];
    }
    clean[i] = r;
}

// --------------------------------------------------------------------------------------------
// exact rank selection over the unmasked pixels (radix select on order-preserving keys,
// 11 + 11 + 10 bits): astroscrappy's background level = lower median a[(n-1)/2]
// --------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(512)
select_hist_kernel(const float *__restrict__ img, const uint8_t *__restrict__ inmask, size_t n, SelState *st)
{
    __shared__ unsigned int h[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const unsigned int prefix = st->prefix;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        if (inmask && inmask[p]) continue;
        const unsigned int key = f32_key(img[p]);
        if (PASS == 0) atomicAdd(&h[key >> 21], 1u);
        else if (PASS == 1) { if ((key >> 21) == (prefix >> 21)) atomicAdd(&h[(key >> 10) & 0x7ffu], 1u); }
        else { if ((key >> 10) == (prefix >> 10)) atomicAdd(&h[key & 0x3ffu], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&st->hist[PASS][i], h[i]);
}

template <int PASS>
__global__ void __launch_bounds__(256) select_scan_kernel(SelState *st, float *out)
{
    unsigned long long k = st->k;
    if (PASS == 0) {
        __shared__ unsigned long long s_tot[33];
        unsigned long long mine = 0;
        for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) mine += st->hist[0][i];
        const unsigned long long total = block_sum(mine, s_tot);
        if (total == 0) { if (threadIdx.x == 0) { *out = 0.0f; st->k = ~0ull; } return; }
        k = (total - 1) / 2;
    }
    if (k == ~0ull) return;
    unsigned long long below;
    const int b = select_find_bin(st->hist[PASS], (PASS == 2) ? 1024 : SEL_BINS, k, below);
    if (threadIdx.x != 0) return;
    st->k = k - below;
    if (PASS == 0) st->prefix = (unsigned int)b << 21;
    else if (PASS == 1) st->prefix |= (unsigned int)b << 10;
    else { st->prefix |= (unsigned int)b; *out = key_f32(st->prefix); }
}

extern "C" size_t bbx_select_work_bytes(void) { return sizeof(SelState); }

extern "C" int bbx_masked_lower_median(const float *img, const uint8_t *inmask, size_t n, void *work,
                                       float *out, void *stream)
{
    BBX_REQUIRE(img && work && out, "bbx_masked_lower_median: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    SelState *st = (SelState *)work;
    BBX_CUDA(cudaMemsetAsync(st, 0, sizeof(SelState), s));
    const int blocks = BBX_SM_COUNT * 4;
    select_hist_kernel<0><<<blocks, 512, 0, s>>>(img, inmask, n, st);
    select_scan_kernel<0><<<1, 256, 0, s>>>(st, out);
    select_hist_kernel<1><<<blocks, 512, 0, s>>>(img, inmask, n, st);
    select_scan_kernel<1><<<1, 256, 0, s>>>(st, out);
    select_hist_kernel<2><<<blocks, 512, 0, s>>>(img, inmask, n, st);
    select_scan_kernel<2><<<1, 256, 0, s>>>(st, out);
    BBX_CHECK_LAUNCH("bbx_masked_lower_median");
    return 0;
}

// --------------------------------------------------------------------------------------------
// host entry points
// --------------------------------------------------------------------------------------------
static size_t lac_align(size_t v) { return (v + 255) / 256 * 256; }

size_t lac_dense_work_bytes(int H, int W)
{
    const size_t n = (size_t)H * W;
    return 3 * lac_align(n * sizeof(float)) + lac_align(n) + lac_align(sizeof(SelState)) + 256;
}

struct LacWork {
    float *s, *noise, *f3;
    uint8_t *flags;
    SelState *sel;
    float *background;
};

static LacWork carve_lac_work(void *work, size_t n)
{
    LacWork w;
    uint8_t *p = (uint8_t *)work;
    w.s = (float *)p; p += lac_align(n * sizeof(float));
    w.noise = (float *)p; p += lac_align(n * sizeof(float));
    w.f3 = (float *)p; p += lac_align(n * sizeof(float));
    w.flags = p; p += lac_align(n);
    w.sel = (SelState *)p; p += lac_align(sizeof(SelState));
    w.background = (float *)p;
    return w;
}

LacParams lac_make_params(float sigclip, float sigfrac, float objlim, float readnoise, const double *readnoise_dev)
{
    LacParams prm;
    prm.sigclip = sigclip;
    prm.sigcliplow = sigfrac * sigclip;      // float32 product, as the reference's C float
    prm.objlim = objlim;
    prm.readnoise = readnoise;
    prm.readnoise_dev = readnoise_dev;
    return prm;
}

int lac_dense_iteration(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, const LacParams &prm,
                        int it, void *work, long long *info, cudaStream_t st)
{
    const LacWork w = carve_lac_work(work, (size_t)H * W);
    const dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
    lac_stage1_kernel<<<grid, 256, 0, st>>>(img, w.s, w.noise, w.f3, H, W, prm, info);
    lac_stage2_kernel<<<grid, 256, 0, st>>>(w.s, w.noise, w.f3, inmask, w.flags, H, W, prm, info, nullptr, nullptr);
    lac_grow_kernel<<<grid, 256, 0, st>>>(w.flags, crmask, H, W, it, info);
    lac_control_kernel<<<1, 1, 0, st>>>(info, it);
    lac_clean_kernel<<<grid, 256, 0, st>>>(img, crmask, inmask, H, W, w.background, info);
    BBX_CHECK_LAUNCH("lac_dense_iteration");
    return 0;
}

int lac_dense_begin(const float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, int niter, void *work,
                    long long *info, cudaStream_t st)
{
    const size_t n = (size_t)H * W;
    const LacWork w = carve_lac_work(work, n);
    BBX_CUDA(cudaMemsetAsync(crmask, 0, n, st));
    lac_init_kernel<<<1, 32, 0, st>>>(info, INFO_NCR + niter);
    BBX_CHECK_LAUNCH("lac_init_kernel");
    return bbx_masked_lower_median(img, inmask, n, w.sel, w.background, (void *)st) ? -2 : 0;
}

extern "C" int bbx_medfilt(const float *in, float *out, int H, int W, int ksize, void *stream)
{
    BBX_REQUIRE(in && out && in != out, "bbx_medfilt: null or aliased argument");
    const dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
    cudaStream_t s = (cudaStream_t)stream;
    if (ksize == 3) medfilt_kernel<3><<<grid, 256, 0, s>>>(in, out, H, W);
    else if (ksize == 5) medfilt_kernel<5><<<grid, 256, 0, s>>>(in, out, H, W);
    else if (ksize == 7) medfilt_kernel<7><<<grid, 256, 0, s>>>(in, out, H, W);
    else BBX_REQUIRE(false, "bbx_medfilt: kernel size %d not in {3, 5, 7}", ksize);
    BBX_CHECK_LAUNCH("bbx_medfilt");
    return 0;
}

extern "C" int bbx_laplace_plus(const float *in, float *out, int H, int W, void *stream)
{
    BBX_REQUIRE(in && out && in != out, "bbx_laplace_plus: null or aliased argument");
    const dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
    laplace_plus_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W);
    BBX_CHECK_LAUNCH("bbx_laplace_plus");
    return 0;
}
