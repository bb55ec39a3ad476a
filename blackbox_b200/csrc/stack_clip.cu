// stack_clip.cu -- optional sigma-clipped master combine (see include/bbx.h: bbx_stack_clipped_median)
#include "bbx_common.cuh"
#include "median_networks.cuh"
#include "stack_common.cuh"

// --------------------------------------------------------------------------------------------
// Sigma-clipped variant (NOT what master_prep does -- blackbox.py:4984 is the plain median above;
// this is the combine BASELINE.json's north star words, kept as an option):
//   clipped = astropy.stats.sigma_clip(cube, sigma, maxiters, cenfunc='median', stdfunc='std', axis=0)
//   out     = np.ma.median(clipped, axis=0)          ((a + b) / 2 in float32; NaN if nothing is left)
// Per pixel, as astropy's C loop does it: non-finite values never take part; repeat: mean and
// population std of the survivors (float64, sums in frame order), centre = their median (mean of
// the two middle values in float64), keep lo <= x <= hi; stop when nothing was rejected or after
// `maxiters` rounds.  The final mask comes from the last bounds applied to all values.
//
// A thread owns a pixel; its N values live in a shared-memory column (conflict-free: consecutive
// threads, consecutive banks) next to their ranks (rank = position in the sorted order, ties by
// frame index; computed once with N^2 comparisons).  Survivors are always a contiguous range of
// ranks, so every median is two rank lookups.
#define CLIP_THREADS 128

__global__ void __launch_bounds__(CLIP_THREADS)
stack_clipmed_kernel(StackArgs a, int n, size_t npix, double sigma, int maxiters, int flat_fix,
                     const uint8_t *__restrict__ bpm, int edge_value, float *__restrict__ out)
{
    extern __shared__ unsigned char clip_smem[];
    float *val = reinterpret_cast<float *>(clip_smem);                                  // [n][CLIP_THREADS]
    unsigned char *rnk = clip_smem + sizeof(float) * (size_t)n * CLIP_THREADS;          // [n][CLIP_THREADS]
    const int t = threadIdx.x;
    for (size_t p0 = (size_t)blockIdx.x * CLIP_THREADS; p0 < npix; p0 += (size_t)gridDim.x * CLIP_THREADS) {
        const size_t p = p0 + t;
        if (p < npix) {
            int nfin = 0;
            for (int k = 0; k < n; k++) {
                float v = __ldcs(a.frames[k] + p);
                v = v / (a.scale[k] != 0.0f ? a.scale[k] : 1.0f);
                val[k * CLIP_THREADS + t] = v;
                nfin += (fabsf(v) <= 3.402823466e+38f);
            }
            // ranks among the finite values (non-finite ones get 255: never looked up)
            for (int k = 0; k < n; k++) {
                const float v = val[k * CLIP_THREADS + t];
                int r = 0;
                if (fabsf(v) <= 3.402823466e+38f) {
                    for (int j = 0; j < n; j++) {
                        const float u = val[j * CLIP_THREADS + t];
                        r += (fabsf(u) <= 3.402823466e+38f) && ((u < v) || (u == v && j < k));
                    }
                } else r = 255;
                rnk[k * CLIP_THREADS + t] = (unsigned char)r;
            }
            auto by_rank = [&](int r) {
                float v = 0.f;
                for (int k = 0; k < n; k++)
                    if (rnk[k * CLIP_THREADS + t] == r) v = val[k * CLIP_THREADS + t];
                return v;
            };
            // survivors = finite values inside [LO, HI] = ranks [ra, ra + count)
            double LO = -INFINITY, HI = INFINITY, lo = NAN, hi = NAN;
            int count = nfin, ra = 0, iteration = 0;
            while (count > 0) {
                double mean = 0.0, sd = 0.0;
                for (int k = 0; k < n; k++) {
                    const double x = (double)val[k * CLIP_THREADS + t];
                    if (fabs(x) <= 3.402823466e+38 && x >= LO && x <= HI) mean += x;
                }
                mean /= (double)count;
                for (int k = 0; k < n; k++) {
                    const double x = (double)val[k * CLIP_THREADS + t];
                    if (fabs(x) <= 3.402823466e+38 && x >= LO && x <= HI) { const double d = mean - x; sd += d * d; }
                }
                sd = sqrt(sd / (double)count);
                const double m0 = (double)by_rank(ra + (count - 1) / 2), m1 = (double)by_rank(ra + count / 2);
                const double cen = (count & 1) ? m0 : (m0 + m1) / 2.0;
                lo = cen - sigma * sd;
                hi = cen + sigma * sd;
                LO = fmax(LO, lo);
                HI = fmin(HI, hi);
                int kept = 0, below = 0;
                for (int k = 0; k < n; k++) {
                    const double x = (double)val[k * CLIP_THREADS + t];
                    const bool fin = fabs(x) <= 3.402823466e+38;
                    kept += fin && x >= LO && x <= HI;
                    below += fin && x < LO;
                }
                ra = below;
                if (kept == count) break;
                count = kept;
                iteration++;
                if (maxiters >= 0 && iteration >= maxiters) break;
            }
            // final mask: the LAST bounds applied to all finite values
            int m = 0, below = 0;
            for (int k = 0; k < n; k++) {
                const double x = (double)val[k * CLIP_THREADS + t];
                const bool fin = fabs(x) <= 3.402823466e+38;
                m += fin && !(x < lo) && !(x > hi);
                below += fin && x < lo;
            }
            float r = NAN;
            if (m > 0) {
                const float f0 = by_rank(below + (m - 1) / 2), f1 = by_rank(below + m / 2);
                const float s2 = f0 + f1;
                r = s2 / 2.0f;
            }
            if (flat_fix) {
                const bool edge = bpm != nullptr && bpm[p] == (uint8_t)edge_value;
                if (edge || r <= 0.0f) r = 1.0f;
            }
            out[p] = r;
        }
    }
}

// Register variant for N <= 32: the N values in frame order (for the sums) and a sorted copy (for
// the medians) both live in registers; same arithmetic, same results as the kernel above.
#define CLIP_REG_MAX 32

template <int N>
__global__ void __launch_bounds__(128, 4)
stack_clipmed_reg_kernel(StackArgs a, size_t npix, double sigma, int maxiters, int scaled, int flat_fix,
                         const uint8_t *__restrict__ bpm, int edge_value, float *__restrict__ out)
{
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
        float v[N], s[N];
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = __ldcs(a.frames[k] + p);
        if (scaled) {                                    // uniform: bias stacks skip the N divisions
#pragma unroll
            for (int k = 0; k < N; k++) v[k] = v[k] / (a.scale[k] != 0.0f ? a.scale[k] : 1.0f);
        }
        unsigned int member = 0;                         // bit k: value k takes part in the next round
#pragma unroll
        for (int k = 0; k < N; k++) {
            const bool fin = fabsf(v[k]) <= 3.402823466e+38f;
            member |= fin ? (1u << k) : 0u;
            s[k] = fin ? v[k] : INFINITY;                // non-finite values sort to the end, never looked up
        }
        const unsigned int finite = member;
        StackSort<N>::run(s);
        auto by_rank = [&](int r) {
            float x = 0.f;
#pragma unroll
            for (int k = 0; k < N; k++) x = (k == r) ? s[k] : x;
            return x;
        };
        // A float32 value v satisfies v >= LO (double) exactly when v >= RU(LO) in float32, and
        // v <= HI when v <= RD(HI): the membership tests run on the float32 pipe
        double LO = -INFINITY, HI = INFINITY, lo = NAN, hi = NAN;
        int count = __popc(member), ra = 0, iteration = 0;
        while (count > 0) {
            double mean = 0.0, sd = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++)
                if ((member >> k) & 1u) mean += (double)v[k];
            mean /= (double)count;
#pragma unroll
            for (int k = 0; k < N; k++)
                if ((member >> k) & 1u) { const double d = mean - (double)v[k]; sd += d * d; }
            sd = sqrt(sd / (double)count);
            const double m0 = (double)by_rank(ra + (count - 1) / 2), m1 = (double)by_rank(ra + count / 2);
            const double cen = (count & 1) ? m0 : (m0 + m1) / 2.0;
            lo = cen - sigma * sd;
            hi = cen + sigma * sd;
            LO = fmax(LO, lo);
            HI = fmin(HI, hi);
            const float LOf = __double2float_ru(LO), HIf = __double2float_rd(HI);
            unsigned int keep = 0;
            int below = 0;
#pragma unroll
            for (int k = 0; k < N; k++) {
                const bool fin = (finite >> k) & 1u;
                keep |= (fin && v[k] >= LOf && v[k] <= HIf) ? (1u << k) : 0u;
                below += fin && v[k] < LOf;
            }
            ra = below;
            const int kept = __popc(keep);
            member = keep;
            if (kept == count) break;
            count = kept;
            iteration++;
            if (maxiters >= 0 && iteration >= maxiters) break;
        }
        // final mask: the LAST bounds applied to all finite values
        const float lof = __double2float_ru(lo), hif = __double2float_rd(hi);
        int m = 0, below = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const bool fin = (finite >> k) & 1u;
            m += fin && !(v[k] < lof) && !(v[k] > hif);
            below += fin && v[k] < lof;
        }
        float r = NAN;
        if (m > 0) {
            const float f0 = by_rank(below + (m - 1) / 2), f1 = by_rank(below + m / 2);
            const float s2 = f0 + f1;
            r = s2 / 2.0f;
        }
        if (flat_fix) {
            const bool edge = bpm != nullptr && bpm[p] == (uint8_t)edge_value;
            if (edge || r <= 0.0f) r = 1.0f;
        }
        out[p] = r;
    }
}

template <int N>
static int dispatch_clip(int n, const StackArgs &a, size_t npix, double sigma, int maxiters, int flat_fix,
                         const uint8_t *bpm, int edge_value, float *out, cudaStream_t s)
{
    if (n == N) {
        const size_t want = (npix + 127) / 128;
        const int blocks = (int)(want < (size_t)BBX_SM_COUNT * 32 ? want : (size_t)BBX_SM_COUNT * 32);
        int scaled = 0;
        for (int k = 0; k < N; k++) scaled |= a.scale[k] != 0.0f;
        stack_clipmed_reg_kernel<N><<<blocks, 128, 0, s>>>(a, npix, sigma, maxiters, scaled, flat_fix, bpm, edge_value, out);
        BBX_CHECK_LAUNCH("stack_clipmed_reg_kernel");
        return 0;
    }
    if constexpr (N > 1) return dispatch_clip<N - 1>(n, a, npix, sigma, maxiters, flat_fix, bpm, edge_value, out, s);
    else return -1;
}

extern "C" int bbx_stack_clipped_median(const float *const *frames_h, const float *scale_h, int N, size_t npix,
                                        double sigma, int maxiters, int flat_fix, const uint8_t *bpm,
                                        int edge_value, float *out, void *stream)
{
    BBX_REQUIRE(frames_h && out, "bbx_stack_clipped_median: null argument");
    BBX_REQUIRE(N >= 1 && N <= STACK_MAX, "bbx_stack_clipped_median: %d frames not in 1..%d", N, STACK_MAX);
    BBX_REQUIRE(sigma > 0.0, "bbx_stack_clipped_median: sigma %g", sigma);
    StackArgs a;
    for (int k = 0; k < STACK_MAX; k++) {
        a.frames[k] = k < N ? frames_h[k] : nullptr;
        a.scale[k] = (k < N && scale_h) ? scale_h[k] : 0.0f;
        BBX_REQUIRE(k >= N || a.frames[k] != nullptr, "bbx_stack_clipped_median: frame %d is null", k);
    }
    if (npix == 0) return 0;
    if (N <= CLIP_REG_MAX)
        return dispatch_clip<CLIP_REG_MAX>(N, a, npix, sigma, maxiters, flat_fix, bpm, edge_value, out, (cudaStream_t)stream);
    const size_t smem = (size_t)N * CLIP_THREADS * (sizeof(float) + 1);
    const size_t want = (npix + CLIP_THREADS - 1) / CLIP_THREADS;
    const int blocks = (int)(want < (size_t)BBX_SM_COUNT * 16 ? want : (size_t)BBX_SM_COUNT * 16);
    stack_clipmed_kernel<<<blocks, CLIP_THREADS, smem, (cudaStream_t)stream>>>(a, N, npix, sigma, maxiters, flat_fix, bpm,
                                                                             edge_value, out);
    BBX_CHECK_LAUNCH("stack_clipmed_kernel");
    return 0;
}

