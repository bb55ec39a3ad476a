// apply.cu -- fused per-pixel pass of the reduction chain (K3) and small elementwise helpers
//
// reduce_apply_kernel: one pass raw -> reduced
//   v = f32(raw) * gain[chan]                 gain_corr      blackbox.py:7460
//   v = f32(f64(v) - vos_fit[chan][row])      os_corr        blackbox.py:6553 / 6556
//   v = f32(f64(v) - oscan[chan][col])        os_corr        blackbox.py:6844
//   crop to the data sections                 os_corr        blackbox.py:6847
//   v = v - mbias                             master bias    blackbox.py:1679
//   non-finite -> 0, 'bad' if unmasked        mask_init      blackbox.py:4408-4414
//   sat = f64(v) >= satlevel[chan]            mask_init      blackbox.py:4494, 4538
//   v = v / mflat                             master flat    blackbox.py:1825
// Every step is a separately rounded IEEE operation (library built with -fmad=false), so the
// result is bit-identical to numpy's given identical fit vectors.
//
// HBM-bound: per output pixel 2 B (u16 raw) + 4 (mbias) + 4 (mflat) + 1 (bpm) read, 4 + 1
// written.  reduce_apply_strip_kernel (the vectorised path): a thread owns 4 consecutive columns
// of the reduced frame and walks down APPLY_ROWS rows, so everything that depends on the column
// only -- channel, gain, the 4 overscan values, the saturation threshold -- lives in registers
// and the inner loop is 8-byte raw / 16-byte f32 / 4-byte mask streaming accesses plus one
// broadcast load of the row's vertical-overscan fit value.  (The first version recomputed the
// index arithmetic with integer divisions per pixel group and converted to float64 for the
// saturation test: ncu showed the XU pipe at 63 % and DRAM at 53 %.)  The saturation test
// f64(v) >= level is done in float32 against the smallest float32 >= level, which is the same
// predicate.
#include "apply_common.cuh"
#include "bg_track.cuh"

template <typename T>
__device__ __forceinline__ void apply_px(float &v, uint8_t &m, bool have_mask, float gn, double fitv,
                                         double osc, bool has_bias, float mb, bool has_flat, float mf,
                                         bool has_sat, double satl, int bit_bad, int bit_sat)
{
    v = v * gn;
    v = sub_f64(v, fitv);
    v = sub_f64(v, osc);
    if (has_bias) v = v - mb;
    if (have_mask) {
        if (!isfinite(v)) { v = 0.f; if (m == 0) m |= (uint8_t)bit_bad; }
        if (has_sat && (double)v >= satl) m |= (uint8_t)(bit_sat | BBX_TMP_SAT);
    }
    if (has_flat) v = v / mf;
}

// VEC = 4: xsize_chan % 4 == 0 and all row starts suitably aligned; VEC = 1: generic
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
reduce_apply_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain, ApplyArgs a)
{
    const int RW = g.nx * g.xsize_chan;                 // reduced width
    const int RH = g.ny * g.ysize_chan;
    const int groups_per_row = RW / VEC;
    const long long total = (long long)RH * groups_per_row;
    const bool have_mask = a.out_mask != nullptr;
    for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total;
         gi += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(gi / groups_per_row);
        const int x = (int)(gi - (long long)y * groups_per_row) * VEC;
        const int r = y / g.ysize_chan, ly = y - r * g.ysize_chan;
        const int c = x / g.xsize_chan, lx = x - c * g.xsize_chan;
        const int ch = r * g.nx + c;
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + ly;
        const size_t ro = (size_t)rr * g.W + (size_t)c * g.dx + lx;
        const size_t oo = (size_t)y * RW + x;
        const float gn = gain.v[ch];
        const double fitv = a.vos_fit ? a.vos_fit[(size_t)ch * g.dy + (rr - r * g.dy)] : 0.0;
        const double satl = a.satlevel ? a.satlevel[ch] : 0.0;
        float v[VEC], mb[VEC], mf[VEC];
        uint8_t m[VEC];
        if (VEC == 4) {
            RawVec4<T>::load(raw + ro, v);
            if (a.mbias) { const uint4 u = ld_stream_u4(a.mbias + oo); mb[0] = __uint_as_float(u.x); mb[1] = __uint_as_float(u.y); mb[2] = __uint_as_float(u.z); mb[3] = __uint_as_float(u.w); }
            if (a.mflat) { const uint4 u = ld_stream_u4(a.mflat + oo); mf[0] = __uint_as_float(u.x); mf[1] = __uint_as_float(u.y); mf[2] = __uint_as_float(u.z); mf[3] = __uint_as_float(u.w); }
            uint32_t mm = 0;
            if (a.bpm) mm = *reinterpret_cast<const uint32_t *>(a.bpm + oo);
#pragma unroll
            for (int k = 0; k < 4; k++) m[k] = (uint8_t)(mm >> (8 * k));
        } else {
            v[0] = raw_to_f32<T>(raw[ro]);
            if (a.mbias) mb[0] = a.mbias[oo];
            if (a.mflat) mf[0] = a.mflat[oo];
            m[0] = a.bpm ? a.bpm[oo] : 0;
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const double osc = a.oscan ? a.oscan[(size_t)ch * g.xsize_chan + lx + k] : 0.0;
            apply_px<T>(v[k], m[k], have_mask, gn, fitv, osc, a.mbias != nullptr, mb[k], a.mflat != nullptr, mf[k],
                        a.satlevel != nullptr, satl, a.bit_bad, a.bit_sat);
        }
        if (a.seeds && have_mask) {
            // pixels found saturated (type 0) and pixels whose bad-pixel mask already carries a
            // saturated / saturated-connected bit (type 1, bit 31) seed the sparse morphology
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const bool sat = (m[k] & BBX_TMP_SAT) != 0;
                if (sat || (m[k] & a.seed_bits)) {
                    const unsigned int slot = atomicAdd(a.seed_count, 1u);
                    if (slot < a.seed_cap) a.seeds[slot] = (unsigned int)(oo + k) | (sat ? 0u : 0x80000000u);
                }
            }
        }
        if (VEC == 4) {
            st_stream_u4(a.out_img + oo, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
            if (have_mask) st_stream_u32(a.out_mask + oo, (uint32_t)m[0] | ((uint32_t)m[1] << 8) | ((uint32_t)m[2] << 16) | ((uint32_t)m[3] << 24));
        } else {
            a.out_img[oo] = v[0];
            if (have_mask) a.out_mask[oo] = m[0];
        }
    }
}

#define APPLY_THREADS 128
#define APPLY_ROWS 16

// STATS: also take the statistics of LACosmic's background level (the count of unmasked pixels, of
// those below the bracket, the histogram inside it: lacosmic_sparse.cu) while the values are in
// registers -- against the seed mask; the mask morphology that follows corrects them for the pixels
// it masks (bg_track.cuh).  The dense Laplacian scan then reads neither the mask nor computes keys.
// FULL: every optional input and output is there (overscan vectors, master bias and flat, bad-pixel
// mask, saturation levels, mask and seed list: the pipeline's case) -- a compile-time fact, so the
// per-pixel code carries no tests of pointers (a tenth of the generic kernel's instructions).
template <typename T, bool STATS, bool FULL>
__global__ void __launch_bounds__(APPLY_THREADS, 8)
reduce_apply_strip_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain, ApplyArgs a, BgState *bg,
                          unsigned int *__restrict__ bghist)
{
    const int RW = g.nx * g.xsize_chan, RH = g.ny * g.ysize_chan;
    const int x = (blockIdx.x * APPLY_THREADS + threadIdx.x) * 4;
    __shared__ unsigned int s_stat[2];
    if (STATS) {
        if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
        __syncthreads();
    }
    if (x >= RW) return;
    const int c = x / g.xsize_chan, lx = x - c * g.xsize_chan;
    const int ya = blockIdx.y * APPLY_ROWS, yb = min(ya + APPLY_ROWS, RH);
    const bool have_mask = FULL || a.out_mask != nullptr;
    const bool has_bias = FULL || a.mbias != nullptr, has_flat = FULL || a.mflat != nullptr;
    const bool has_bpm = FULL || a.bpm != nullptr, has_fit = FULL || a.vos_fit != nullptr;
    const bool has_osc = FULL || a.oscan != nullptr, has_seeds = FULL || a.seeds != nullptr;
    int r_cur = -1;
    double osc[4] = {0.0, 0.0, 0.0, 0.0};
    float gn = 1.0f, satl = 0.0f;
    bool has_sat = false;
    unsigned int key_a = 0, width = 0, n_valid = 0, n_below = 0;
    if (STATS) { key_a = bg->key_a; width = bg->width; }
    // bracket at a non-negative value: the order-preserving keys are the raw float bits (sp_scan_kernel)
    const bool raw_bits = key_a >= 0x80000000u;
    const int lo_bits = (int)(key_a & 0x7fffffffu);

    struct RowIn { float v[4]; float4 mb, mf; uint32_t mm; double fitv; };
    // everything a row needs from memory, issued back to back (two rows are in flight at a time)
    auto load_row = [&](int y, RowIn &in) {
        const int r = (y >= g.ysize_chan) ? 1 : 0;          // ny == 2
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + (y - r * g.ysize_chan);
        const size_t ro = (size_t)rr * g.W + (size_t)c * g.dx + lx;
        const size_t oo = (size_t)y * RW + x;
        RawVec4<T>::load(raw + ro, in.v);
        in.mb = make_float4(0.f, 0.f, 0.f, 0.f);
        in.mf = make_float4(1.f, 1.f, 1.f, 1.f);
        in.mm = 0;
        if (has_bias) in.mb = __ldcs(reinterpret_cast<const float4 *>(a.mbias + oo));
        if (has_flat) in.mf = __ldcs(reinterpret_cast<const float4 *>(a.mflat + oo));
        if (has_bpm) in.mm = __ldcs(reinterpret_cast<const unsigned int *>(a.bpm + oo));
        in.fitv = has_fit ? a.vos_fit[(size_t)(r * g.nx + c) * g.dy + (rr - r * g.dy)] : 0.0;
    };
    auto finish_row = [&](int y, const RowIn &in) {
        const int r = (y >= g.ysize_chan) ? 1 : 0;
        if (r != r_cur) {                          // once per strip (twice if it straddles the CCD halves)
            r_cur = r;
            const int ch = r * g.nx + c;
            gn = gain.v[ch];
            if (has_osc) {
#pragma unroll
                for (int k = 0; k < 4; k++) osc[k] = a.oscan[(size_t)ch * g.xsize_chan + lx + k];
            }
            has_sat = have_mask && (FULL || a.satlevel != nullptr);
            if (has_sat) {
                const double lv = a.satlevel[ch];
                if (lv != lv) has_sat = false;     // NaN level: the comparison is never true
                else satl = f32_ceil_of(lv);
            }
        }
        const size_t oo = (size_t)y * RW + x;
        const float mb[4] = {in.mb.x, in.mb.y, in.mb.z, in.mb.w}, mf[4] = {in.mf.x, in.mf.y, in.mf.z, in.mf.w};
        float v[4];
        uint32_t mout = 0;
        bool any_seed = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float w = in.v[k] * gn;
            w = sub_f64(w, in.fitv);
            w = sub_f64(w, osc[k]);
            if (has_bias) w = w - mb[k];
            uint32_t m = (in.mm >> (8 * k)) & 0xffu;
            if (have_mask) {
                if (!isfinite(w)) { w = 0.f; if (m == 0) m |= (uint32_t)a.bit_bad; }
                if (has_sat && w >= satl) m |= (uint32_t)(a.bit_sat | BBX_TMP_SAT);
            }
            if (has_flat) w = w / mf[k];
            v[k] = w;
            mout |= m << (8 * k);
        }
        // (one test for the four mask bytes of the row)
        any_seed = have_mask && (mout & (((uint32_t)(BBX_TMP_SAT | a.seed_bits) & 0xffu) * 0x01010101u)) != 0;
        if (has_seeds && any_seed) {
            // pixels found saturated (type 0) and pixels whose bad-pixel mask already carries a
            // saturated / saturated-connected bit (type 1, bit 31) seed the sparse morphology
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t m = (mout >> (8 * k)) & 0xffu;
                const bool sat = (m & BBX_TMP_SAT) != 0;
                if (sat || (m & a.seed_bits)) {
                    const unsigned int slot = atomicAdd(a.seed_count, 1u);
                    if (slot < a.seed_cap) a.seeds[slot] = (unsigned int)(oo + k) | (sat ? 0u : 0x80000000u);
                }
            }
        }
        __stcs(reinterpret_cast<float4 *>(a.out_img + oo), make_float4(v[0], v[1], v[2], v[3]));
        if (have_mask) __stcs(reinterpret_cast<unsigned int *>(a.out_mask + oo), mout);
        if (STATS) {
            if (mout == 0 && raw_bits) {
                n_valid += 4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int bits = __float_as_int(v[k]);
                    n_below += bits < lo_bits;
                    const unsigned int d = (unsigned int)bits - (unsigned int)lo_bits;
                    if (d < width) atomicAdd(&bghist[d], 1u);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if ((mout >> (8 * k)) & 0xffu) continue;
                    const unsigned int key = f32_key(v[k]);
                    n_valid++;
                    if (key < key_a) n_below++;
                    else if (key - key_a < width) atomicAdd(&bghist[key - key_a], 1u);
                }
            }
        }
    };

    int y = ya;
    for (; y + 1 < yb; y += 2) {
        RowIn in0, in1;
        load_row(y, in0);
        load_row(y + 1, in1);
        finish_row(y, in0);
        finish_row(y + 1, in1);
    }
    if (y < yb) {
        RowIn in0;
        load_row(y, in0);
        finish_row(y, in0);
    }
    if (STATS) {
        // per warp (the lanes beyond the frame have left), per CTA in shared memory, then two atomics
        // per CTA: one atomic per thread on the same two words cost 3 ms per frame
        const unsigned int m = __activemask();
        const unsigned int tv = __reduce_add_sync(m, n_valid), tb = __reduce_add_sync(m, n_below);
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) {
            if (tv) atomicAdd(&s_stat[0], tv);
            if (tb) atomicAdd(&s_stat[1], tb);
        }
        __syncthreads();
        if (threadIdx.x == 0) {                       // (thread 0 of a CTA is always inside the frame)
            if (s_stat[0]) atomicAdd(&bg->n_valid, (unsigned long long)s_stat[0]);
            if (s_stat[1]) atomicAdd(&bg->n_below, (unsigned long long)s_stat[1]);
        }
    }
}

__global__ void satlevels_kernel(ChanF64 sat_e, const double *__restrict__ biasm, double *__restrict__ out)
{
    const int i = threadIdx.x;
    if (i < BBX_NCHAN) out[i] = sat_e.v[i] - biasm[i];
}

__global__ void __launch_bounds__(256) gain_corr_kernel(float *__restrict__ raw, bbx_geom g, ChanF32 gain)
{
    const long long total = (long long)g.H * g.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / g.W), x = (int)(i - (long long)y * g.W);
        const int ch = (y / g.dy) * g.nx + x / g.dx;
        raw[i] = raw[i] * gain.v[ch];
    }
}

__global__ void __launch_bounds__(256) binary_inplace_kernel(float *__restrict__ a, const float *__restrict__ b, size_t n, int op)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a[i] = (op == 0) ? a[i] - b[i] : a[i] / b[i];
}

__global__ void __launch_bounds__(256) mask_or_kernel(uint8_t *__restrict__ mask, const uint8_t *__restrict__ flag, size_t n, int bit)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (flag[i]) mask[i] |= (uint8_t)bit;
}

static inline bool aligned_to(const void *p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; }

extern "C" int bbx_reduce_apply(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                                const double *vos_fit, const double *oscan, const float *mbias, const float *mflat,
                                const uint8_t *bpm, const double *satlevel, const bbx_maskbits *bits,
                                float *out_img, uint8_t *out_mask, unsigned int *seeds, unsigned int *seed_count,
                                unsigned int seed_cap, void *stream)
{
    return apply_launch(raw, raw_type, g, gain_h, vos_fit, oscan, mbias, mflat, bpm, satlevel, bits, out_img, out_mask,
                        seeds, seed_count, seed_cap, nullptr, nullptr, stream);
}

// bg / bghist != null: the strip kernel also takes LACosmic's background statistics (the 4-pixel
// aligned layout is then required)
int apply_launch(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                 const double *vos_fit, const double *oscan, const float *mbias, const float *mflat,
                 const uint8_t *bpm, const double *satlevel, const bbx_maskbits *bits,
                 float *out_img, uint8_t *out_mask, unsigned int *seeds, unsigned int *seed_count,
                 unsigned int seed_cap, void *bg_state, unsigned int *bghist, void *stream)
{
    BBX_REQUIRE(g && raw && out_img, "bbx_reduce_apply: null raw / geometry / output");
    BBX_REQUIRE(g->ny == 2 && g->nx * g->ny == BBX_NCHAN, "bbx_reduce_apply: expected 2 x 8 channels");
    BBX_REQUIRE(out_mask == nullptr || bits != nullptr, "bbx_reduce_apply: mask output needs the mask bit values");
    BBX_REQUIRE((const void *)out_img != raw, "bbx_reduce_apply: output must not alias the raw frame");
    ChanF32 gn;
    for (int i = 0; i < BBX_NCHAN; i++) gn.v[i] = gain_h ? gain_h[i] : 1.0f;
    BBX_REQUIRE((seeds == nullptr) == (seed_count == nullptr), "bbx_reduce_apply: seeds and seed_count go together");
    BBX_REQUIRE((long long)g->ny * g->ysize_chan * g->nx * g->xsize_chan < 2147483647LL || seeds == nullptr,
                "bbx_reduce_apply: frame too large for 31-bit seed indices");
    ApplyArgs a = {vos_fit, oscan, mbias, mflat, bpm, satlevel, out_img, out_mask, bits ? bits->bad : 0,
                   bits ? bits->saturated : 0, seeds, seed_count, seed_cap,
                   bits ? (unsigned int)(bits->saturated | bits->satcon) : 0u};
    if (seeds) BBX_CUDA(cudaMemsetAsync(seed_count, 0, sizeof(unsigned int), (cudaStream_t)stream));
    const size_t esz = raw_type == BBX_RAW_U16 ? 2 : 4;
    const long long RW = (long long)g->nx * g->xsize_chan, RH = (long long)g->ny * g->ysize_chan;
    const bool vec4 = (g->xsize_chan % 4 == 0) && (g->dx % 4 == 0) && (g->W % 4 == 0) &&
                      aligned_to(raw, 4 * esz) && aligned_to(mbias, 16) && aligned_to(mflat, 16) &&
                      aligned_to(out_img, 16) && aligned_to(bpm, 4) && aligned_to(out_mask, 4);
    const long long groups = RH * RW / (vec4 ? 4 : 1);
    const int blocks = (int)((groups + 255) / 256 < (long long)BBX_SM_COUNT * 16 ? (groups + 255) / 256 : BBX_SM_COUNT * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec4 && (RH + APPLY_ROWS - 1) / APPLY_ROWS <= 65535) {
        const dim3 grid((unsigned int)((RW / 4 + APPLY_THREADS - 1) / APPLY_THREADS), (unsigned int)((RH + APPLY_ROWS - 1) / APPLY_ROWS));
        BgState *bg = (BgState *)bg_state;
        if (bg) {
            BBX_REQUIRE(out_mask != nullptr && bghist != nullptr, "bbx_reduce_apply: statistics need the mask output");
            const bool full = vos_fit && oscan && mbias && mflat && bpm && satlevel && seeds;
            if (raw_type == BBX_RAW_U16) {
                if (full) reduce_apply_strip_kernel<uint16_t, true, true><<<grid, APPLY_THREADS, 0, s>>>((const uint16_t *)raw, *g, gn, a, bg, bghist);
                else reduce_apply_strip_kernel<uint16_t, true, false><<<grid, APPLY_THREADS, 0, s>>>((const uint16_t *)raw, *g, gn, a, bg, bghist);
            } else reduce_apply_strip_kernel<float, true, false><<<grid, APPLY_THREADS, 0, s>>>((const float *)raw, *g, gn, a, bg, bghist);
        } else {
            const bool full = vos_fit && oscan && mbias && mflat && bpm && satlevel && seeds && out_mask;
            if (raw_type == BBX_RAW_U16) {
                if (full) reduce_apply_strip_kernel<uint16_t, false, true><<<grid, APPLY_THREADS, 0, s>>>((const uint16_t *)raw, *g, gn, a, nullptr, nullptr);
                else reduce_apply_strip_kernel<uint16_t, false, false><<<grid, APPLY_THREADS, 0, s>>>((const uint16_t *)raw, *g, gn, a, nullptr, nullptr);
            } else reduce_apply_strip_kernel<float, false, false><<<grid, APPLY_THREADS, 0, s>>>((const float *)raw, *g, gn, a, nullptr, nullptr);
        }
        BBX_CHECK_LAUNCH("bbx_reduce_apply");
        return 0;
    }
    BBX_REQUIRE(bg_state == nullptr, "bbx_reduce_apply: the statistics need the 4-pixel aligned layout");
    if (raw_type == BBX_RAW_U16) {
        if (vec4) reduce_apply_kernel<uint16_t, 4><<<blocks, 256, 0, s>>>((const uint16_t *)raw, *g, gn, a);
        else reduce_apply_kernel<uint16_t, 1><<<blocks, 256, 0, s>>>((const uint16_t *)raw, *g, gn, a);
    } else {
        if (vec4) reduce_apply_kernel<float, 4><<<blocks, 256, 0, s>>>((const float *)raw, *g, gn, a);
        else reduce_apply_kernel<float, 1><<<blocks, 256, 0, s>>>((const float *)raw, *g, gn, a);
    }
    BBX_CHECK_LAUNCH("bbx_reduce_apply");
    return 0;
}

extern "C" int bbx_satlevels(const double *sat_e_h, const double *biasm, double *out_satlevel, void *stream)
{
    BBX_REQUIRE(sat_e_h && biasm && out_satlevel, "bbx_satlevels: null argument");
    ChanF64 se;
    for (int i = 0; i < BBX_NCHAN; i++) se.v[i] = sat_e_h[i];
    satlevels_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(se, biasm, out_satlevel);
    BBX_CHECK_LAUNCH("bbx_satlevels");
    return 0;
}

extern "C" int bbx_gain_corr(float *raw, const bbx_geom *g, const float *gain_h, void *stream)
{
    BBX_REQUIRE(raw && g && gain_h, "bbx_gain_corr: null argument");
    ChanF32 gn;
    for (int i = 0; i < BBX_NCHAN; i++) gn.v[i] = gain_h[i];
    gain_corr_kernel<<<BBX_SM_COUNT * 16, 256, 0, (cudaStream_t)stream>>>(raw, *g, gn);
    BBX_CHECK_LAUNCH("bbx_gain_corr");
    return 0;
}

extern "C" int bbx_binary_inplace(float *a, const float *b, size_t n, int op, void *stream)
{
    BBX_REQUIRE(a && b, "bbx_binary_inplace: null argument");
    BBX_REQUIRE(op == 0 || op == 1, "bbx_binary_inplace: op %d (0 = subtract, 1 = divide)", op);
    if (n == 0) return 0;
    binary_inplace_kernel<<<BBX_SM_COUNT * 16, 256, 0, (cudaStream_t)stream>>>(a, b, n, op);
    BBX_CHECK_LAUNCH("bbx_binary_inplace");
    return 0;
}

extern "C" int bbx_mask_or(uint8_t *mask, const uint8_t *flag, size_t n, int bit, void *stream)
{
    BBX_REQUIRE(mask && flag, "bbx_mask_or: null argument");
    if (n == 0) return 0;
    mask_or_kernel<<<BBX_SM_COUNT * 16, 256, 0, (cudaStream_t)stream>>>(mask, flag, n, bit);
    BBX_CHECK_LAUNCH("bbx_mask_or");
    return 0;
}

// BIASMEAN / RDNOISE = np.nanmean over the 16 channel values, summed in numpy's pairwise order
// for 16 float64 (8 accumulators a[j] + a[j+8], then a balanced tree); blackbox.py:6865-6868
__global__ void header_means_kernel(const double *__restrict__ biasm, const double *__restrict__ std_vos,
                                    double *__restrict__ out)
{
    if (threadIdx.x >= 2) return;
    const double *a = threadIdx.x == 0 ? biasm : std_vos;
    double r[8];
    int cnt = 0;
    for (int j = 0; j < 8; j++) {
        const double x = a[j], y = a[j + 8];
        const bool nx = (x != x), ny = (y != y);
        cnt += !nx + !ny;
        r[j] = (nx ? 0.0 : x) + (ny ? 0.0 : y);
    }
    const double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    out[threadIdx.x] = cnt ? s / (double)cnt : NAN;
}

extern "C" int bbx_header_means(const double *biasm, const double *std_vos, double *out, void *stream)
{
    BBX_REQUIRE(biasm && std_vos && out, "bbx_header_means: null argument");
    header_means_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(biasm, std_vos, out);
    BBX_CHECK_LAUNCH("bbx_header_means");
    return 0;
}
