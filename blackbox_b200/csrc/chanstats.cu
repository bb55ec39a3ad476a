// chanstats.cu -- exact per-channel medians of a reduced frame and the edge-pixel fill of
// blackbox_reduce (reference: blackbox.py:1958-1974: "set edge pixels to median of
// corresponding channel": data[sec_chan][mask_edge_chan] = np.median(data[sec_chan])).
//
// np.median of a float32 channel (6 969 600 pixels, an even count) is the float32 mean of the two
// middle order statistics; any NaN in the channel makes it NaN.  Both ranks are found by a
// three-pass radix select on the order-preserving 32-bit key of a float32 (11 + 11 + 10 bits),
// all 16 channels at once: every pass reads the frame once (446 MB) and histograms, per channel,
// the next key digit of the pixels that still match the prefix of either rank.  A thread only
// touches the shared-memory histogram when the digit differs from that of its previous pixel
// (pass 0: practically every pixel of a sky-dominated frame has the same top 11 bits, so plain
// atomics would serialise 32-fold).
#include "bbx_common.cuh"

#define CS_BINS 2048
#define CS_THREADS 256

struct ChanSel {
    unsigned long long k[2];      // rank still to find inside the current prefix, per target
    unsigned int prefix[2];       // key bits fixed so far (right-aligned)
    unsigned int two;             // 1: two targets (even count), 0: one
    unsigned int nan_count;
    unsigned long long n;         // pixels of the channel
    unsigned int ignore_nan, pad; // 1: np.nanmedian (NaNs do not count), 0: np.median (any NaN -> NaN)
    unsigned int hist[2][CS_BINS];
};

__device__ __forceinline__ unsigned int cs_key(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float cs_unkey(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void cs_init_kernel(ChanSel *sel, unsigned long long n, unsigned int ignore_nan)
{
    ChanSel &s = sel[blockIdx.x];
    for (int i = threadIdx.x; i < 2 * CS_BINS; i += blockDim.x) (&s.hist[0][0])[i] = 0;
    if (threadIdx.x == 0) {
        s.two = (n % 2 == 0) ? 1u : 0u;
        s.k[0] = s.two ? n / 2 - 1 : n / 2;
        s.k[1] = n / 2;
        s.prefix[0] = s.prefix[1] = 0;
        s.nan_count = 0;
        s.n = n;
        s.ignore_nan = ignore_nan; s.pad = 0;
    }
}

// PASS 0: digit = key >> 21 (one histogram); PASS 1: (key >> 10) & 2047 where key >> 21 matches;
// PASS 2: key & 1023 where key >> 10 matches.  grid = (chunks of a channel row, rows, channels).
template <int PASS>
__global__ void __launch_bounds__(CS_THREADS)
cs_hist_kernel(const float *__restrict__ img, int W, int ysc, int xsc, int rows_per_block, ChanSel *sel)
{
    __shared__ unsigned int h[2][CS_BINS];
    __shared__ unsigned int s_nan;
    const int ch = blockIdx.z, r = ch >> 3, c = ch & 7;
    ChanSel &cs = sel[ch];
    for (int i = threadIdx.x; i < 2 * CS_BINS; i += CS_THREADS) (&h[0][0])[i] = 0;
    if (threadIdx.x == 0) s_nan = 0;
    __syncthreads();
    const unsigned int p0 = cs.prefix[0], p1 = cs.prefix[1];
    const bool two = PASS > 0 && cs.two && p1 != p0;
    const int y0 = blockIdx.y * rows_per_block, y1 = min(y0 + rows_per_block, ysc);
    // coalesced float4 reads: consecutive threads take consecutive groups of 4 pixels of a channel
    // row (when the channel rows start 16-byte aligned, i.e. xsc % 4 == 0; scalar reads otherwise); a
    // thread's successive pixels are 4 KB apart but almost always fall into the same bin in pass
    // 0, which is all the run-length trick needs
    unsigned int nan = 0;
    int cur = -1, cur_t = 0;
    unsigned int cnt = 0;
    auto consume = [&](float v) {
        if (PASS == 0 && v != v) { nan++; return; }
        const unsigned int key = cs_key(v);
        int bin, tgt = 0;
        if (PASS == 0) bin = (int)(key >> 21);
        else if (PASS == 1) {
            const unsigned int pre = key >> 21;
            if (pre == p0) tgt = 0; else if (two && pre == p1) tgt = 1; else return;
            bin = (int)((key >> 10) & 2047u);
        } else {
            const unsigned int pre = key >> 10;
            if (pre == p0) tgt = 0; else if (two && pre == p1) tgt = 1; else return;
            bin = (int)(key & 1023u);
        }
        if (bin == cur && tgt == cur_t) { cnt++; return; }
        if (cnt) atomicAdd(&h[cur_t][cur], cnt);
        cur = bin; cur_t = tgt; cnt = 1;
    };
    if (xsc % 4 == 0 && ((uintptr_t)img & 15) == 0) {
        const int g4 = xsc / 4;
        const int total = (y1 - y0) * g4;
        for (int t = blockIdx.x * CS_THREADS + threadIdx.x; t < total; t += gridDim.x * CS_THREADS) {
            const int yy = y0 + t / g4, xg = t % g4;
            const float4 q = *reinterpret_cast<const float4 *>(img + (size_t)(r * ysc + yy) * W + (size_t)c * xsc + 4 * xg);
            consume(q.x); consume(q.y); consume(q.z); consume(q.w);
        }
    } else {                                                  // any width / alignment
        const int total = (y1 - y0) * xsc;
        for (int t = blockIdx.x * CS_THREADS + threadIdx.x; t < total; t += gridDim.x * CS_THREADS) {
            const int yy = y0 + t / xsc, x = t % xsc;
            consume(img[(size_t)(r * ysc + yy) * W + (size_t)c * xsc + x]);
        }
    }
    if (cnt) atomicAdd(&h[cur_t][cur], cnt);
    if (PASS == 0 && nan) atomicAdd(&s_nan, nan);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS_BINS; i += CS_THREADS) {
        const unsigned int v = (&h[0][0])[i];
        if (v) atomicAdd(&(&cs.hist[0][0])[i], v);
    }
    if (PASS == 0 && threadIdx.x == 0 && s_nan) atomicAdd(&cs.nan_count, s_nan);
}

// one block per channel: locate the bin of each target rank, extend the prefixes, clear the
// histograms; after the last pass write the median
template <int PASS>
__global__ void __launch_bounds__(CS_THREADS)
cs_find_kernel(ChanSel *sel, float *out_med)
{
    __shared__ unsigned long long part[CS_THREADS];
    ChanSel &cs = sel[blockIdx.x];
    const int nb = PASS == 2 ? 1024 : CS_BINS, per = nb / CS_THREADS;
    if (PASS == 0 && cs.ignore_nan) {                           // the ranks among the non-NaN pixels
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long m = cs.n - cs.nan_count;
            cs.two = (m % 2 == 0 && m > 0) ? 1u : 0u;
            cs.k[0] = m == 0 ? 0 : (cs.two ? m / 2 - 1 : m / 2);
            cs.k[1] = m / 2;
        }
        __syncthreads();
    }
    const int ntgt = cs.two ? 2 : 1;
    const bool shared_hist = PASS == 0 || cs.prefix[0] == cs.prefix[1];
    unsigned int newp[2] = {0, 0};
    unsigned long long newk[2] = {0, 0};
    for (int t = 0; t < ntgt; t++) {
        const unsigned int *hist = cs.hist[shared_hist ? 0 : t];
        unsigned long long mine = 0;
        for (int j = 0; j < per; j++) mine += hist[threadIdx.x * per + j];
        part[threadIdx.x] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            // NaNs sort after everything: a rank beyond the finite pixels ends in the last bin
            unsigned long long k = cs.k[t], acc = 0;
            int seg = CS_THREADS - 1;
            for (int q = 0; q < CS_THREADS; q++) {
                if (acc + part[q] > k) { seg = q; break; }
                acc += part[q];
            }
            int bin = seg * per + per - 1;
            for (int j = 0; j < per; j++) {
                const unsigned long long hv = hist[seg * per + j];
                if (acc + hv > k) { bin = seg * per + j; break; }
                acc += hv;
            }
            newp[t] = (cs.prefix[t] << (PASS == 2 ? 10 : 11)) | (unsigned int)bin;
            newk[t] = k - acc;
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CS_BINS; i += CS_THREADS) (&cs.hist[0][0])[i] = 0;
    if (threadIdx.x == 0) {
        if (ntgt == 1) { newp[1] = newp[0]; newk[1] = newk[0]; }
        cs.prefix[0] = newp[0]; cs.prefix[1] = newp[1];
        cs.k[0] = newk[0]; cs.k[1] = newk[1];
        if (PASS == 2) {
            float med;
            if (cs.ignore_nan ? (cs.n == cs.nan_count) : (cs.nan_count != 0)) med = __int_as_float(0x7fc00000);
            else if (cs.two) { const float s = cs_unkey(newp[0]) + cs_unkey(newp[1]); med = s / 2.0f; }
            else med = cs_unkey(newp[0]);
            out_med[blockIdx.x] = med;
        }
    }
}

__global__ void __launch_bounds__(256)
cs_fill_edge_kernel(float *img, const uint8_t *__restrict__ mask, int H, int W, int ysc, int xsc,
                    unsigned int edge_bit, const float *__restrict__ med)
{
    const size_t n4 = (size_t)H * W / 4;
    const unsigned int e4 = edge_bit * 0x01010101u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned int m = reinterpret_cast<const unsigned int *>(mask)[i] & e4;
        if (!m) continue;
        const size_t p = i * 4;
        const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((m >> (8 * k)) & 0xffu) img[p + k] = med[(y / ysc) * 8 + (x + k) / xsc];
    }
}

extern "C" size_t bbx_chanmed_work_bytes(void) { return sizeof(ChanSel) * BBX_NCHAN; }

extern "C" int bbx_channel_medians(const float *img, int H, int W, int ysize_chan, int xsize_chan, int ignore_nan,
                                   void *work, float *out_med, void *stream)
{
    BBX_REQUIRE(img && work && out_med, "bbx_channel_medians: null argument");
    BBX_REQUIRE(ysize_chan > 0 && xsize_chan > 0 && H == 2 * ysize_chan && W == 8 * xsize_chan,
                "bbx_channel_medians: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    cudaStream_t st = (cudaStream_t)stream;
    ChanSel *sel = (ChanSel *)work;
    const unsigned long long n = (unsigned long long)ysize_chan * xsize_chan;
    const int rows_per_block = 64;
    const int groups = rows_per_block * (xsize_chan / 4);
    dim3 grid((unsigned int)max(1, min(4, (groups + CS_THREADS * 8 - 1) / (CS_THREADS * 8))),
              (unsigned int)ceil_div(ysize_chan, rows_per_block), BBX_NCHAN);
    cs_init_kernel<<<BBX_NCHAN, CS_THREADS, 0, st>>>(sel, n, ignore_nan ? 1u : 0u);
    cs_hist_kernel<0><<<grid, CS_THREADS, 0, st>>>(img, W, ysize_chan, xsize_chan, rows_per_block, sel);
    cs_find_kernel<0><<<BBX_NCHAN, CS_THREADS, 0, st>>>(sel, out_med);
    cs_hist_kernel<1><<<grid, CS_THREADS, 0, st>>>(img, W, ysize_chan, xsize_chan, rows_per_block, sel);
    cs_find_kernel<1><<<BBX_NCHAN, CS_THREADS, 0, st>>>(sel, out_med);
    cs_hist_kernel<2><<<grid, CS_THREADS, 0, st>>>(img, W, ysize_chan, xsize_chan, rows_per_block, sel);
    cs_find_kernel<2><<<BBX_NCHAN, CS_THREADS, 0, st>>>(sel, out_med);
    BBX_CHECK_LAUNCH("bbx_channel_medians");
    return 0;
}

extern "C" int bbx_fill_edge(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                             int edge_bit, const float *med, void *stream)
{
    BBX_REQUIRE(img && mask && med, "bbx_fill_edge: null argument");
    BBX_REQUIRE(ysize_chan > 0 && xsize_chan > 0 && H == 2 * ysize_chan && W == 8 * xsize_chan,
                "bbx_fill_edge: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    BBX_REQUIRE(W % 4 == 0 && ((uintptr_t)mask & 3) == 0, "bbx_fill_edge: width / mask alignment must be a multiple of 4");
    BBX_REQUIRE(edge_bit > 0 && edge_bit < 256, "bbx_fill_edge: edge bit %d", edge_bit);
    cs_fill_edge_kernel<<<BBX_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(img, mask, H, W, ysize_chan, xsize_chan,
                                                                          (unsigned int)edge_bit, med);
    BBX_CHECK_LAUNCH("bbx_fill_edge");
    return 0;
}
