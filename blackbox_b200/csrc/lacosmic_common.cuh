// lacosmic_common.cuh -- pieces shared by the dense (lacosmic.cu) and the sparse / lazy
// (lacosmic_sparse.cu) LACosmic implementations.  See lacosmic.cu for the algorithm statement.
#pragma once
#include "bbx_common.cuh"
#include "median_networks.cuh"
#include "bg_track.cuh"

// out_info layout (int64): [0] iterations run, [1] active flag, [2] status bits,
// [3] reserved, [4+k] new CR pixels of iteration k
#define INFO_ITERS 0
#define INFO_ACTIVE 1
#define INFO_STATUS 2
#define INFO_NCR 4
#define LAC_STATUS_OVERFLOW 1   // a candidate list overflowed: result incomplete, redo densely
#define LAC_STATUS_NEED_BG 2    // a CR pixel without usable neighbours needs the background level

// --------------------------------------------------------------------------------------------
// building blocks
// --------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ float median_at(const float *__restrict__ in, int H, int W, int y, int x)
{
    constexpr int R = K / 2;
    if (y < R || y >= H - R || x < R || x >= W - R) return in[(size_t)y * W + x];
    float v[K * K];
#pragma unroll
    for (int dy = 0; dy < K; dy++)
#pragma unroll
        for (int dx = 0; dx < K; dx++) v[dy * K + dx] = in[(size_t)(y + dy - R) * W + (x + dx - R)];
    if constexpr (K == 3) return bbx_med9(v);
    else if constexpr (K == 5) return bbx_med25(v);
    else return bbx_med49(v);
}

// L+ of pixel (y,x): the four Laplacian values of its 2x2 sub-pixels, clipped at 0, averaged
__device__ __forceinline__ float laplace_plus_at(const float *__restrict__ in, int H, int W, int y, int x)
{
    const size_t i = (size_t)y * W + x;
    const float c = in[i];
    const bool hl = x > 0, hr = x + 1 < W, hu = y > 0, hd = y + 1 < H;   // u = previous row
    const float l = hl ? in[i - 1] : 0.f, r = hr ? in[i + 1] : 0.f;
    const float u = hu ? in[i - W] : 0.f, d = hd ? in[i + W] : 0.f;
    const float c4 = 4.0f * c;
    // order of subtraction: right, left, next row, previous row (missing neighbours skipped)
    float s00 = c4 - c; if (hl) s00 = s00 - l; s00 = s00 - c; if (hu) s00 = s00 - u;
    float s01 = c4; if (hr) s01 = s01 - r; s01 = s01 - c; s01 = s01 - c; if (hu) s01 = s01 - u;
    float s10 = c4 - c; if (hl) s10 = s10 - l; if (hd) s10 = s10 - d; s10 = s10 - c;
    float s11 = c4; if (hr) s11 = s11 - r; s11 = s11 - c; if (hd) s11 = s11 - d; s11 = s11 - c;
    s00 = s00 < 0.f ? 0.f : s00; s01 = s01 < 0.f ? 0.f : s01;
    s10 = s10 < 0.f ? 0.f : s10; s11 = s11 < 0.f ? 0.f : s11;
    float p = s00 + s01;
    p = p + s10;
    p = p + s11;
    return p / 4.0f;
}

struct LacParams {
    float sigclip, sigcliplow, objlim;
    float readnoise;                 // used when readnoise_dev == nullptr
    const double *readnoise_dev;     // device scalar (e.g. RDNOISE computed on the GPU), or null
};

__device__ __forceinline__ float lac_rn2(const LacParams &p)
{
    const float rn = p.readnoise_dev ? (float)(*p.readnoise_dev) : p.readnoise;
    return rn * rn;
}


// s = L+ / (2 noise) of one pixel, noise from the 5x5 median of the image (all border rules
// included).  Also returns the noise.
__device__ __forceinline__ float lac_s_at(const float *__restrict__ img, int H, int W, int y, int x, float rn2,
                                          float &noise_out)
{
    float m5 = median_at<5>(img, H, W, y, x);
    if (m5 < 0.00001f) m5 = 0.00001f;
    float nz = m5 + rn2;
    nz = sqrtf(nz);
    noise_out = nz;
    const float lp = laplace_plus_at(img, H, W, y, x);
    const float den = 2.0f * nz;
    return lp / den;
}

// exact rank selection by radix select on order-preserving keys (11 + 11 + 10 bits)
#define SEL_BINS 2048
struct SelState {
    unsigned long long k;       // rank still to find inside the current prefix
    unsigned int prefix;        // key bits fixed so far
    unsigned int pad;
    unsigned int hist[3][SEL_BINS];
};

// Block-wide search (256 threads): first bin b of hist[0..nb) with cumulative count > k.
// Returns b (clamped to nb-1) and the count below it in `below`; identical in all threads.
__device__ __forceinline__ int select_find_bin(const unsigned int *hist, int nb, unsigned long long k,
                                               unsigned long long &below)
{
    __shared__ unsigned int s_h[SEL_BINS];
    __shared__ unsigned long long s_part[256];
    __shared__ int s_bin;
    __shared__ unsigned long long s_below;
    const int per = SEL_BINS / 256;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) s_h[i] = i < nb ? hist[i] : 0u;
    if (threadIdx.x == 0) { s_bin = nb - 1; s_below = 0; }
    __syncthreads();
    unsigned long long mine = 0;
    for (int i = 0; i < per; i++) mine += s_h[threadIdx.x * per + i];
    s_part[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0, total = 0;
        for (int t = 0; t < 256; t++) total += s_part[t];
        if (k >= total) {                       // rank beyond the data: last bin
            unsigned long long a2 = 0;
            for (int i = 0; i < nb - 1; i++) a2 += s_h[i];
            s_below = a2;
        } else {
            int t = 0;
            for (; t < 256; t++) { if (acc + s_part[t] > k) break; acc += s_part[t]; }
            int b = t * per;
            for (; b < nb - 1; b++) { if (acc + s_h[b] > k) break; acc += s_h[b]; }
            s_bin = b;
            s_below = acc;
        }
    }
    __syncthreads();
    below = s_below;
    return s_bin;
}

// dense implementation (lacosmic.cu)
struct LacWork;
size_t lac_dense_work_bytes(int H, int W);
int lac_dense_begin(const float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, int niter, void *work,
                    long long *info, cudaStream_t st);
int lac_dense_iteration(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, const LacParams &prm,
                        int it, void *work, long long *info, cudaStream_t st);
LacParams lac_make_params(float sigclip, float sigfrac, float objlim, float readnoise, const double *readnoise_dev);
