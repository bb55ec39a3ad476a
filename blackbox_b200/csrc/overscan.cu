// overscan.cu -- overscan statistics and fits of os_corr (reference: blackbox.py:6407-6879)
//
// Kernels (all HBM/L2-bound reductions over the overscan strips, fp64 accumulation):
//   vos_rowstats_kernel  K1  warp per strip row, values in registers, <=5 clip iterations
//   vos_fit_kernel       F1  block per channel, Forsythe orthogonal-polynomial least squares
//   hos_satcount_kernel  K2c BlackGEM saturated-column counts near the horizontal overscan
//   hos_stats_kernel     K2a block per channel: dlevel, strip masking, column clipped stats
//   vos_std_kernel       K2b 8-CTA cluster per channel, iterative global clip via DSMEM
//   hos_fit_kernel       F2  block per channel: errors, pre-clean, 3x polynomial fit + reject
//
// Sigma clipping follows astropy's fast path (oracle/csrc/bbo.c bbo_clip_bounds): mean and
// population std of the survivors in float64, keep lo <= x <= hi, stop when nothing is
// rejected or after maxiters bound computations; final membership comes from the final
// bounds applied to all unmasked values.  Because every iteration keeps an interval, the
// survivor set after k iterations is "valid AND inside the intersection of all bounds so
// far", so no per-element state is needed.
#include <cooperative_groups.h>
#include "bbx_common.cuh"

namespace cg = cooperative_groups;

#define MASKED_ZERO_TOL 1e-8f   // np.ma.masked_values(data, 0): |x| <= 1e-8 is masked

__device__ __forceinline__ bool vos_valid(float v) { return isfinite(v) && !(fabsf(v) <= MASKED_ZERO_TOL); }

// ============================================================================================
// K1: row-wise clipped mean of the vertical-overscan strips
// ============================================================================================
#define VOS_MAXV 8   // values per lane -> strip width <= 256

template <typename T>
__global__ void __launch_bounds__(256)
vos_rowstats_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain, double sigma,
                    int maxiters, double *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= BBX_NCHAN * g.dy) return;
    const int ch = warp / g.dy, trow = warp - ch * g.dy;
    const int r = ch / g.nx, c = ch - r * g.nx;
    const T *p = raw + (size_t)(r * g.dy + trow) * g.W + (size_t)c * g.dx + g.vos_x0;
    const float gn = gain.v[ch];

    double x[VOS_MAXV];
    bool ok[VOS_MAXV];
#pragma unroll
    for (int k = 0; k < VOS_MAXV; k++) {
        const int j = lane + 32 * k;
        float v = 0.f;
        bool valid = false;
        if (j < g.vos_w) {
            v = raw_to_f32<T>(p[j]) * gn;
            valid = vos_valid(v);
        }
        x[k] = (double)v;
        ok[k] = valid;
    }

    double LO = -INFINITY, HI = INFINITY, flo = NAN, fhi = NAN;
    int cnt = 0;
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < VOS_MAXV; k++) if (ok[k]) { cnt++; sum += x[k]; }
    cnt = warp_sum(cnt);
    sum = warp_sum(sum);
    for (int it = 0; it < maxiters && cnt > 0; it++) {
        const double mean = sum / (double)cnt;
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < VOS_MAXV; k++)
            if (ok[k] && x[k] >= LO && x[k] <= HI) { const double d = mean - x[k]; ss += d * d; }
        ss = warp_sum(ss);
        const double sd = sqrt(ss / (double)cnt);
        flo = mean - sigma * sd;
        fhi = mean + sigma * sd;
        LO = fmax(LO, flo);
        HI = fmin(HI, fhi);
        int ncnt = 0;
        double nsum = 0.0;
#pragma unroll
        for (int k = 0; k < VOS_MAXV; k++)
            if (ok[k] && x[k] >= LO && x[k] <= HI) { ncnt++; nsum += x[k]; }
        ncnt = warp_sum(ncnt);
        nsum = warp_sum(nsum);
        const bool done = (ncnt == cnt);
        cnt = ncnt;
        sum = nsum;
        if (done) break;
    }
    // statistics of everything inside the FINAL bounds
    int fc = 0;
    double fs = 0.0;
#pragma unroll
    for (int k = 0; k < VOS_MAXV; k++)
        if (ok[k] && !(x[k] < flo) && !(x[k] > fhi)) { fc++; fs += x[k]; }
    fc = warp_sum(fc);
    fs = warp_sum(fs);
    if (lane == 0) out[warp] = fc > 0 ? fs / (double)fc : NAN;
}

// ============================================================================================
// block-level helpers: clipped statistics over values produced by an accessor, and
// least-squares polynomial fit on a regular grid by Forsythe's three-term recurrence
// ============================================================================================

// get(i, x) -> bool valid.  Returns count of survivors of the final bounds; mean/std (ddof 0)
// of those survivors; flo/fhi = final bounds.  All threads get identical results.
template <typename Get>
__device__ int block_clip_stats(Get get, int n, double sigma, int maxiters, double *scr_d,
                                int *scr_i, double &mean_out, double &std_out, double &flo,
                                double &fhi)
{
    double LO = -INFINITY, HI = INFINITY;
    flo = NAN; fhi = NAN;
    int c = 0;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double x;
        if (get(i, x)) { c++; s += x; }
    }
    int cnt = block_sum(c, scr_i);
    double sum = block_sum(s, scr_d);
    for (int it = 0; it < maxiters && cnt > 0; it++) {
        const double mean = sum / (double)cnt;
        double ss = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double x;
            if (get(i, x) && x >= LO && x <= HI) { const double d = mean - x; ss += d * d; }
        }
        ss = block_sum(ss, scr_d);
        const double sd = sqrt(ss / (double)cnt);
        flo = mean - sigma * sd;
        fhi = mean + sigma * sd;
        LO = fmax(LO, flo);
        HI = fmin(HI, fhi);
        c = 0; s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double x;
            if (get(i, x) && x >= LO && x <= HI) { c++; s += x; }
        }
        const int ncnt = block_sum(c, scr_i);
        const double nsum = block_sum(s, scr_d);
        const bool done = (ncnt == cnt);
        cnt = ncnt;
        sum = nsum;
        if (done) break;
    }
    // final statistics (two-pass) inside the final bounds
    c = 0; s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double x;
        if (get(i, x) && !(x < flo) && !(x > fhi)) { c++; s += x; }
    }
    const int fc = block_sum(c, scr_i);
    const double fs = block_sum(s, scr_d);
    if (fc == 0) { mean_out = NAN; std_out = NAN; return 0; }
    const double mean = fs / (double)fc;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double x;
        if (get(i, x) && !(x < flo) && !(x > fhi)) { const double d = x - mean; ss += d * d; }
    }
    ss = block_sum(ss, scr_d);
    mean_out = mean;
    std_out = sqrt(ss / (double)fc);
    return fc;
}

// Least squares polynomial of degree deg through the points i in [0,n) with use[i] != 0,
// abscissa x_i = x0 + i, ordinate y[i].  Evaluates the fit at every i into fit[i].
// pk / pkm1: smem work arrays of n doubles.  rec[3*(deg+1)] receives (alpha, beta, coef) so
// the caller can derive monomial coefficients.  Returns the number of points used
// (fit is valid only if that exceeds deg).
__device__ int block_polyfit(const double *y, const uint8_t *use, int n, double x0, int deg,
                             double *fit, double *pk, double *pkm1, double *rec,
                             double *scr_d, int *scr_i, double &tc, double &th)
{
    // centre / half-range of the used abscissae keep |t| <= 1
    int c = 0, imin = n, imax = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (use[i]) { c++; imin = min(imin, i); imax = max(imax, i); }
    const int npts = block_sum(c, scr_i);
    // min / max through sums of one-hot is overkill: use integer reductions via shared atomics
    __shared__ int s_min, s_max;
    if (threadIdx.x == 0) { s_min = n; s_max = -1; }
    __syncthreads();
    if (imax >= 0) { atomicMin(&s_min, imin); atomicMax(&s_max, imax); }
    __syncthreads();
    if (npts <= deg) return npts;
    tc = x0 + 0.5 * (double)(s_min + s_max);
    th = fmax(0.5 * (double)(s_max - s_min), 1.0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { pk[i] = 1.0; pkm1[i] = 0.0; fit[i] = 0.0; }
    __syncthreads();
    double norm_prev = 1.0;
    for (int k = 0; k <= deg; k++) {
        double a = 0.0, b = 0.0, d = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            if (use[i]) {
                const double p = pk[i], t = (x0 + (double)i - tc) / th;
                a += p * p;
                b += y[i] * p;
                d += t * p * p;
            }
        const double norm = block_sum(a, scr_d);
        const double sy = block_sum(b, scr_d);
        const double st = block_sum(d, scr_d);
        const double coef = sy / norm, alpha = st / norm, beta = (k == 0) ? 0.0 : norm / norm_prev;
        if (threadIdx.x == 0) { rec[3 * k] = alpha; rec[3 * k + 1] = beta; rec[3 * k + 2] = coef; }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double p = pk[i], t = (x0 + (double)i - tc) / th;
            fit[i] += coef * p;
            const double pn = (t - alpha) * p - beta * pkm1[i];
            pkm1[i] = p;
            pk[i] = pn;
        }
        norm_prev = norm;
        __syncthreads();
    }
    return npts;
}

// ============================================================================================
// F1: polynomial fit to the vertical-overscan row means
// ============================================================================================
__global__ void __launch_bounds__(256)
vos_fit_kernel(const double *__restrict__ mean_vos, bbx_geom g, int deg, double nsigma,
               double *__restrict__ out_fit, double *__restrict__ out_coef,
               double *__restrict__ out_biasm, int32_t *__restrict__ out_ok)
{
    extern __shared__ double smem_d[];
    const int n = g.dy, ch = blockIdx.x;
    double *v = smem_d, *pk = v + n, *pkm1 = pk + n, *fit = pkm1 + n;
    uint8_t *use = (uint8_t *)(fit + n);
    __shared__ double scr_d[33];
    __shared__ int scr_i[33];
    __shared__ double rec[3 * (BBX_MAX_POLY_DEG + 1)];
    __shared__ int s_bad;

    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = mean_vos[(size_t)ch * n + i];
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();

    double mean, sd, flo, fhi;
    block_clip_stats([&](int i, double &x) { x = v[i]; return isfinite(x); }, n, nsigma, 5,
                     scr_d, scr_i, mean, sd, flo, fhi);
    // rows overlapping the horizontal overscan are not fitted (bottom half: the last rows of
    // the tile, top half: the first rows)
    const int overlap = n - g.ysize_chan;
    const bool top = ch >= g.nx;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        bool u = isfinite(v[i]);
        if (u && sd != 0.0) u = fabs(v[i] - mean) / sd <= nsigma;
        if (!(sd == sd)) u = false;
        if (top ? (i < overlap) : (i >= g.ysize_chan)) u = false;
        use[i] = u;
    }
    __syncthreads();
    double tc, th;
    const int npts = block_polyfit(v, use, n, 0.0, deg, fit, pk, pkm1, rec, scr_d, scr_i, tc, th);
    bool ok = npts > deg;
    if (ok) {
        int bad = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) if (!isfinite(fit[i])) bad = 1;
        if (bad) atomicOr(&s_bad, 1);
        __syncthreads();
        ok = !s_bad;
    }
    if (ok) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { s += fit[i]; out_fit[(size_t)ch * n + i] = fit[i]; }
        s = block_sum(s, scr_d);
        if (threadIdx.x == 0) {
            out_biasm[ch] = s / (double)n;
            out_ok[ch] = 1;
            // monomial coefficients: q_k(t) polynomials by the same recurrence, then t -> x
            double q0[BBX_MAX_POLY_DEG + 1] = {0}, q1[BBX_MAX_POLY_DEG + 1] = {0}, acc[BBX_MAX_POLY_DEG + 1] = {0};
            q1[0] = 1.0;                                   // q1 = p_k, q0 = p_{k-1}
            for (int k = 0; k <= deg; k++) {
                const double alpha = rec[3 * k], beta = rec[3 * k + 1], coef = rec[3 * k + 2];
                double qn[BBX_MAX_POLY_DEG + 1];
                for (int j = 0; j <= deg; j++) acc[j] += coef * q1[j];
                for (int j = 0; j <= deg; j++)
                    qn[j] = (j > 0 ? q1[j - 1] : 0.0) - alpha * q1[j] - beta * q0[j];
                for (int j = 0; j <= deg; j++) { q0[j] = q1[j]; q1[j] = qn[j]; }
            }
            // acc = coefficients in t = (x - tc)/th ; expand (x - tc)^j / th^j
            double mono[BBX_MAX_POLY_DEG + 1] = {0};
            for (int j = 0; j <= deg; j++) {
                double binom = 1.0, scale = acc[j] / pow(th, (double)j);
                for (int m = 0; m <= j; m++) {            // C(j,m) x^m (-tc)^(j-m)
                    mono[m] += scale * binom * pow(-tc, (double)(j - m));
                    binom = binom * (double)(j - m) / (double)(m + 1);
                }
            }
            for (int j = 0; j <= BBX_MAX_POLY_DEG; j++) out_coef[ch * (BBX_MAX_POLY_DEG + 1) + j] = (j <= deg) ? mono[j] : 0.0;
        }
    } else {
        // fallback of the reference: subtract the nan-median of the row means.  Rank by
        // counting (O(n^2), failure path only).
        int c = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) if (isfinite(v[i])) c++;
        const int m = block_sum(c, scr_i);
        __shared__ double s_lo, s_hi;
        if (threadIdx.x == 0) { s_lo = NAN; s_hi = NAN; }
        __syncthreads();
        const int k_hi = m / 2, k_lo = (m - 1) / 2;
        for (int i = threadIdx.x; i < n && m > 0; i += blockDim.x) {
            if (!isfinite(v[i])) continue;
            int rank = 0;
            for (int j = 0; j < n; j++)
                if (isfinite(v[j]) && (v[j] < v[i] || (v[j] == v[i] && j < i))) rank++;
            if (rank == k_lo) s_lo = v[i];
            if (rank == k_hi) s_hi = v[i];
        }
        __syncthreads();
        const double med = 0.5 * (s_lo + s_hi);
        for (int i = threadIdx.x; i < n; i += blockDim.x) out_fit[(size_t)ch * n + i] = med;
        if (threadIdx.x == 0) {
            out_biasm[ch] = med;
            out_ok[ch] = 0;
            for (int j = 0; j <= BBX_MAX_POLY_DEG; j++) out_coef[ch * (BBX_MAX_POLY_DEG + 1) + j] = NAN;
        }
    }
}

// ============================================================================================
// K2c: BlackGEM saturated-column counts (blackbox.py:6624-6640)
// ============================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
hos_satcount_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain,
                    const double *__restrict__ vos_fit, ChanF64 sat_e, int lim1, int lim2,
                    int rows_per_block, int32_t *__restrict__ out_cnt)
{
    const int ch = blockIdx.z, r = ch / g.nx, c = ch - r * g.nx;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.xsize_chan) return;
    // distance d = 0 is the data row adjacent to the horizontal overscan
    const int d0 = blockIdx.y * rows_per_block;
    const int d1 = min(d0 + rows_per_block, min(lim2, g.ysize_chan));
    const double thr = 0.9 * sat_e.v[ch];
    const float gn = gain.v[ch];
    int c1 = 0, c2 = 0;
    for (int d = d0; d < d1; d++) {
        const int ly = (r == 0) ? (g.ysize_chan - 1 - d) : d;            // data-section row
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + ly;    // raw row
        const int trow = rr - r * g.dy;
        float v = raw_to_f32<T>(raw[(size_t)rr * g.W + (size_t)c * g.dx + x]) * gn;
        v = sub_f64(v, vos_fit[(size_t)ch * g.dy + trow]);
        const int hit = ((double)v >= thr);
        c2 += hit;
        if (d < lim1) c1 += hit;
    }
    if (c1) atomicAdd(&out_cnt[((size_t)ch * 2 + 0) * g.xsize_chan + x], c1);
    if (c2) atomicAdd(&out_cnt[((size_t)ch * 2 + 1) * g.xsize_chan + x], c2);
}

// uint16 raw frames: for one row the test  f64(f32(f64(f32(n) * gain) - fit[row])) >= thr  is
// monotone in the raw count n (gain > 0), so it is a comparison of n with a per-row integer
// threshold.  A block finds the thresholds of its SATC_ROWS rows by bisection (one thread per
// row, exactly the arithmetic above), then counts with 4 columns per thread: 8-byte loads and
// two packed 16-bit counters per 32-bit register.
#define SATC_ROWS 64
#define SATC_THREADS 128
__global__ void __launch_bounds__(SATC_THREADS)
hos_satcount_u16_kernel(const uint16_t *__restrict__ raw, bbx_geom g, ChanF32 gain,
                        const double *__restrict__ vos_fit, ChanF64 sat_e, int lim1, int lim2,
                        int32_t *__restrict__ out_cnt)
{
    __shared__ unsigned int s_thr[SATC_ROWS];
    const int ch = blockIdx.z, r = ch / g.nx, c = ch - r * g.nx;
    const int d0 = blockIdx.y * SATC_ROWS, d1 = min(d0 + SATC_ROWS, min(lim2, g.ysize_chan));
    const double thr = 0.9 * sat_e.v[ch];
    const float gn = gain.v[ch];
    if ((int)threadIdx.x < d1 - d0) {
        const int d = d0 + threadIdx.x;
        const int ly = (r == 0) ? (g.ysize_chan - 1 - d) : d;
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + ly;
        const double fv = vos_fit[(size_t)ch * g.dy + (rr - r * g.dy)];
        // smallest n in [0, 65536] whose value reaches thr (65536: none; NaN fit: none)
        unsigned int lo = 0, hi = 65536;
        while (lo < hi) {
            const unsigned int mid = (lo + hi) >> 1;
            float v = (float)mid * gn;
            v = sub_f64(v, fv);
            if ((double)v >= thr) hi = mid; else lo = mid + 1;
        }
        s_thr[threadIdx.x] = lo;
    }
    __syncthreads();
    const int x = (blockIdx.x * SATC_THREADS + threadIdx.x) * 4;
    if (x >= g.xsize_chan) return;
    unsigned int a1x = 0, a1y = 0, a2x = 0, a2y = 0;          // packed counters: columns (0,1) and (2,3)
    const uint16_t *col = raw + (size_t)c * g.dx + x;
#pragma unroll 8
    for (int d = d0; d < d1; d++) {
        const int ly = (r == 0) ? (g.ysize_chan - 1 - d) : d;
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + ly;
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(col + (size_t)rr * g.W));
        const unsigned int t = s_thr[d - d0];
        unsigned int hx, hy;
        if (t > 65535u) { hx = 0; hy = 0; }
        else {
            const unsigned int t2 = t | (t << 16);
            hx = __vcmpgeu2(u.x, t2) & 0x00010001u;
            hy = __vcmpgeu2(u.y, t2) & 0x00010001u;
        }
        a2x += hx; a2y += hy;
        if (d < lim1) { a1x += hx; a1y += hy; }
    }
    int32_t *o1 = out_cnt + ((size_t)ch * 2 + 0) * g.xsize_chan + x, *o2 = out_cnt + ((size_t)ch * 2 + 1) * g.xsize_chan + x;
    if (a1x & 0xffffu) atomicAdd(o1 + 0, (int)(a1x & 0xffffu));
    if (a1x >> 16) atomicAdd(o1 + 1, (int)(a1x >> 16));
    if (a1y & 0xffffu) atomicAdd(o1 + 2, (int)(a1y & 0xffffu));
    if (a1y >> 16) atomicAdd(o1 + 3, (int)(a1y >> 16));
    if (a2x & 0xffffu) atomicAdd(o2 + 0, (int)(a2x & 0xffffu));
    if (a2x >> 16) atomicAdd(o2 + 1, (int)(a2x >> 16));
    if (a2y & 0xffffu) atomicAdd(o2 + 2, (int)(a2y & 0xffffu));
    if (a2y >> 16) atomicAdd(o2 + 3, (int)(a2y >> 16));
}

// ============================================================================================
// K2a: horizontal-overscan strip statistics
// ============================================================================================
#define HOS_MAXR 16

template <typename T>
__global__ void __launch_bounds__(256)
hos_stats_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain,
                 const double *__restrict__ vos_fit, int tel_kind, float data_limit,
                 const int32_t *__restrict__ satcnt, double *__restrict__ out_dlevel,
                 float *__restrict__ out_mean, float *__restrict__ out_std,
                 int32_t *__restrict__ out_n, uint8_t *__restrict__ out_satcol)
{
    extern __shared__ float smem_f[];
    const int ch = blockIdx.x, r = ch / g.nx, c = ch - r * g.nx;
    const int R = g.hos_rows, NC = g.xsize_chan;
    float *val = smem_f;                                  // [R][NC]  (data columns only)
    uint8_t *mh = (uint8_t *)(val + (size_t)R * NC);      // [R][NC]  strip mask
    uint8_t *tmp = mh + (size_t)R * NC;                   // [R][NC]  dilation scratch
    uint8_t *mx = tmp + (size_t)R * NC;                   // [NC]     column flags
    __shared__ double scr_d[33];
    __shared__ int scr_i[33];
    const int y0 = (r == 0) ? g.hos_y0_bot : g.hos_y0_top;
    const float gn = gain.v[ch];

    for (int i = threadIdx.x; i < R * NC; i += blockDim.x) {
        const int rr = i / NC, x = i - rr * NC;
        const int row = y0 + rr;
        float v = raw_to_f32<T>(raw[(size_t)row * g.W + (size_t)c * g.dx + x]) * gn;
        val[i] = sub_f64(v, vos_fit[(size_t)ch * g.dy + (row - r * g.dy)]);
    }
    __syncthreads();

    // level offset: clipped mean (sigma 3, mean centre) of the last 300 data columns
    const int w300 = min(300, NC), xa = NC - w300;
    double dlevel, sd, flo, fhi;
    block_clip_stats([&](int i, double &x) {
                         const int rr = i / w300, xx = xa + (i - rr * w300);
                         x = (double)val[rr * NC + xx];
                         return isfinite(x);
                     }, R * w300, 3.0, 5, scr_d, scr_i, dlevel, sd, flo, fhi);
    if (threadIdx.x == 0) out_dlevel[ch] = dlevel;
    for (int i = threadIdx.x; i < R * NC; i += blockDim.x) val[i] = sub_f64(val[i], dlevel);
    __syncthreads();

    // strip mask
    if (tel_kind == BBX_TEL_ML) {
        for (int i = threadIdx.x; i < R * NC; i += blockDim.x) mh[i] = val[i] > data_limit;
        __syncthreads();
        for (int x = threadIdx.x; x < NC; x += blockDim.x) {
            int s = 0;
            for (int rr = 0; rr < R; rr++) s += mh[rr * NC + x];
            mx[x] = (double)s > 0.5 * (double)R;
        }
        __syncthreads();
        // a flagged column with no flagged neighbour column is released (binary_opening with
        // a 2-element structure removes exactly the isolated ones)
        for (int x = threadIdx.x; x < NC; x += blockDim.x) {
            const bool lone = mx[x] && !(x > 0 && mx[x - 1]) && !(x + 1 < NC && mx[x + 1]);
            if (lone) for (int rr = 0; rr < R; rr++) mh[rr * NC + x] = 0;
        }
        __syncthreads();
        // two 3x3 dilations with zero border = one 5x5 dilation
        for (int i = threadIdx.x; i < R * NC; i += blockDim.x) {
            const int rr = i / NC, x = i - rr * NC;
            uint8_t m = 0;
            for (int a = max(rr - 2, 0); a <= min(rr + 2, R - 1); a++)
                for (int b = max(x - 2, 0); b <= min(x + 2, NC - 1); b++) m |= mh[a * NC + b];
            tmp[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < R * NC; i += blockDim.x) mh[i] = tmp[i];
        for (int x = threadIdx.x; x < NC; x += blockDim.x) out_satcol[(size_t)ch * NC + x] = 0;
    } else {
        for (int x = threadIdx.x; x < NC; x += blockDim.x) {
            const int c1 = satcnt[((size_t)ch * 2 + 0) * NC + x], c2 = satcnt[((size_t)ch * 2 + 1) * NC + x];
            const uint8_t sc = (c1 >= 3) || (c2 >= 10);
            out_satcol[(size_t)ch * NC + x] = sc;
            for (int rr = 0; rr < R; rr++) mh[rr * NC + x] = sc;
        }
    }
    __syncthreads();

    // column-wise clip (sigma 2.5, mean centre, <= 5 iterations); one thread per column,
    // sequential float64 sums in row order; then float32 mean / std(ddof=1) of the survivors
    for (int x = threadIdx.x; x < NC; x += blockDim.x) {
        double xv[HOS_MAXR];
        bool ok[HOS_MAXR];
        int cnt = 0;
        for (int rr = 0; rr < R; rr++) {
            xv[rr] = (double)val[rr * NC + x];
            ok[rr] = !mh[rr * NC + x] && isfinite(xv[rr]);
            cnt += ok[rr];
        }
        double flo2 = NAN, fhi2 = NAN;
        bool in[HOS_MAXR];
        for (int rr = 0; rr < R; rr++) in[rr] = ok[rr];
        for (int it = 0; it < 5 && cnt > 0; it++) {
            double m = 0.0, ss = 0.0;
            for (int rr = 0; rr < R; rr++) if (in[rr]) m += xv[rr];
            m /= (double)cnt;
            for (int rr = 0; rr < R; rr++) if (in[rr]) { const double d = m - xv[rr]; ss += d * d; }
            const double s = sqrt(ss / (double)cnt);
            flo2 = m - 2.5 * s;
            fhi2 = m + 2.5 * s;
            int nc = 0;
            for (int rr = 0; rr < R; rr++) { in[rr] = in[rr] && xv[rr] >= flo2 && xv[rr] <= fhi2; nc += in[rr]; }
            const bool done = (nc == cnt);
            cnt = nc;
            if (done) break;
        }
        int n = 0;
        float tot = 0.f;
        for (int rr = 0; rr < R; rr++) {
            in[rr] = ok[rr] && !(xv[rr] < flo2) && !(xv[rr] > fhi2);
            if (in[rr]) { n++; tot = tot + val[rr * NC + x]; }
        }
        float mean = NAN, sdev = NAN;
        if (n > 0) {
            mean = (float)((double)tot / (double)n);
            if (n > 1) {
                float var = 0.f;
                for (int rr = 0; rr < R; rr++)
                    if (in[rr]) { const float d = val[rr * NC + x] - mean; const float q = d * d; var = var + q; }
                var = (float)((double)var / (double)(n - 1));
                sdev = sqrtf(var);
            }
        }
        out_mean[(size_t)ch * NC + x] = mean;
        out_std[(size_t)ch * NC + x] = sdev;
        out_n[(size_t)ch * NC + x] = n;
    }
}

// ============================================================================================
// K2b: clipped std of the overscan-subtracted vertical-overscan strip (RDN{i})
// one cluster of VSTD_CLUSTER CTAs per channel; partial sums are exchanged through
// distributed shared memory and combined in rank order (deterministic)
// ============================================================================================
#define VSTD_CLUSTER 8
#define VSTD_THREADS 512
#define VSTD_MAXV 6          // strip width <= 192 (174 unbinned, 87 binned)
#define VSTD_LIST_PER_WARP 640
#define VSTD_CORE_SIGMA 2.25
#define VSTD_SAMPLE_ROWS 8

struct VstdPartial { double s, q, cs, cq; long long n, cn; int ovf, pad; };

#ifdef VSTD_PROFILE
// development build only (BBX_NVCC_EXTRA=-DVSTD_PROFILE): clock64 at the phase boundaries of CTA 0
__device__ long long g_vstd_clk[32];
__device__ int g_vstd_nclk;
#define VSTD_TICK() do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && g_vstd_nclk < 32) g_vstd_clk[g_vstd_nclk++] = clock64(); } while (0)
extern "C" int bbx_debug_vstd_clocks(long long *out_h, int *n_h)
{
    int n = 0;
    cudaMemcpyFromSymbol(&n, g_vstd_nclk, sizeof(int));
    cudaMemcpyFromSymbol(out_h, g_vstd_clk, sizeof(long long) * 32);
    *n_h = n;
    n = 0;
    cudaMemcpyToSymbol(g_vstd_nclk, &n, sizeof(int));
    return 0;
}
#else
#define VSTD_TICK() do { } while (0)
#endif

// u16 -> f32 through the exponent trick (2^23 + n) - 2^23: exact, and off the quarter-rate
// conversion unit, which the float32 <-> float64 conversions below need
template <typename T> __device__ __forceinline__ float raw_to_f32_alu(T v);
template <> __device__ __forceinline__ float raw_to_f32_alu<uint16_t>(uint16_t v)
{
    return __uint_as_float(0x4b000000u | (unsigned int)v) - 8388608.0f;        // exact for v < 2^23
}
template <> __device__ __forceinline__ float raw_to_f32_alu<float>(float v) { return v; }

// The strip is walked ONCE in the common case.  A warp owns rows rank*rows_per_cta + warp,
// +nwarps, ...; a lane holds the <= 6 strip values of its row in registers (the next row's
// loads are in flight during the arithmetic).  Every clip iteration needs (count, sum, sum of
// squares) inside the current interval; the population variance follows as q/n - mean^2 (the
// values are overscan residuals of a few e-, so there is no cancellation to speak of).
//
// The single walk accumulates the moments of all valid values (the first evaluation), the
// moments of the values inside a core band, and parks every value outside the band (a few per
// cent) in a per-warp shared-memory list, in a deterministic order.  The band is
// mean +- 2.25 sd of a once-clipped sample of VSTD_SAMPLE_ROWS rows spread over the channel
// (float32 bounds, so the band test is a float32 comparison).  As long as the clip bounds of
// an evaluation contain the band (they are 3 sigma bounds of the whole strip, so they do unless
// the sample is unrepresentative) the evaluation is "band moments + the listed values inside
// the bounds"; otherwise, or if a list overflows, it walks the strip again.
template <typename T>
__global__ void __cluster_dims__(VSTD_CLUSTER, 1, 1) __launch_bounds__(VSTD_THREADS, 2)
vos_std_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain,
               const double *__restrict__ vos_fit, const double *__restrict__ dlevel_arr,
               double *__restrict__ out_std)
{
    extern __shared__ float s_list[];                        // [warps][VSTD_LIST_PER_WARP]
    cg::cluster_group cluster = cg::this_cluster();
    const int ch = blockIdx.y, r = ch / g.nx, c = ch - r * g.nx;
    const int rank = (int)cluster.block_rank();
    const int rows_per = (g.dy + VSTD_CLUSTER - 1) / VSTD_CLUSTER;
    const int row0 = rank * rows_per, row1 = min(row0 + rows_per, g.dy);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float gn = gain.v[ch];
    const double *fitrow = vos_fit + (size_t)ch * g.dy;
    const double dlevel = dlevel_arr[ch];
    const int hos_t0 = ((r == 0) ? g.hos_y0_bot : g.hos_y0_top) - r * g.dy;
    const T *base = raw + (size_t)(r * g.dy) * g.W + (size_t)c * g.dx + g.vos_x0;
    float *mylist = s_list + (size_t)warp * VSTD_LIST_PER_WARP;
    int nlist = 0;                                           // entries in this warp's list (warp-uniform)

    __shared__ VstdPartial wpart[VSTD_THREADS / 32];   // per-warp partial sums
    __shared__ VstdPartial part[2];     // this CTA's sums; double-buffered slot read by the other CTAs
    __shared__ VstdPartial total;       // sums of the whole cluster
    __shared__ float s_band[2];
    int phase = 0;

    struct Mom { double S, Q; long long N; };
    double CS = 0.0, CQ = 0.0;          // band moments of the whole channel
    long long CN = 0;
    bool use_list = false;
    float clo = 0.f, chi = 0.f;         // core band (float32 values; empty if clo > chi)

    // all fields through the fixed xor-butterfly: deterministic
    auto warp_reduce = [&](VstdPartial &v) {
        v.s = warp_sum(v.s); v.q = warp_sum(v.q); v.n = warp_sum(v.n);
        v.cs = warp_sum(v.cs); v.cq = warp_sum(v.cq); v.cn = warp_sum(v.cn); v.ovf = warp_sum(v.ovf);
    };
    // publish this CTA's partial sums, combine those of the cluster (two block barriers and
    // one cluster barrier per evaluation)
    auto exchange = [&](double s, double q, long long n, double cs, double cq, long long cn, int ovf,
                        Mom &m, bool with_core) {
        VstdPartial v = {s, q, cs, cq, n, cn, ovf, 0};
        warp_reduce(v);
        if (lane == 0) wpart[warp] = v;
        __syncthreads();
        if (warp == 0) {
            VstdPartial t = {0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0};
            if (lane < nwarps) t = wpart[lane];
            warp_reduce(t);
            if (lane == 0) part[phase] = t;
        }
        cluster.sync();
        if (warp == 0) {
            VstdPartial t = {0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0};
            if (lane < VSTD_CLUSTER) t = cluster.map_shared_rank(part, lane)[phase];
            warp_reduce(t);
            if (lane == 0) total = t;
        }
        __syncthreads();
        m.S = total.s; m.Q = total.q; m.N = total.n;
        if (with_core) { CS = total.cs; CQ = total.cq; CN = total.cn; use_list = (total.ovf == 0); }
        phase ^= 1;      // the next publish uses the other slot: no remote read can be overtaken
    };

    auto load_row = [&](int trow, T *dst) {
        const T *p = base + (size_t)trow * g.W;
#pragma unroll
        for (int k = 0; k < VSTD_MAXV; k++) {
            const int j = lane + 32 * k;
            dst[k] = (j < g.vos_w) ? p[j] : (T)0;
        }
    };
    // the overscan-subtracted value of one strip pixel (float32, as the reference holds it)
    auto value = [&](T rawv, double fv, bool in_hos) {
        float x = raw_to_f32_alu<T>(rawv) * gn;
        x = sub_f64(x, fv);
        if (in_hos) x = sub_f64(x, dlevel);
        return x;
    };

    VSTD_TICK();
    // which of this lane's VSTD_MAXV strip columns exist (bit k: column lane + 32 k < vos_w)
    unsigned int kmask = 0;
#pragma unroll
    for (int k = 0; k < VSTD_MAXV; k++) kmask |= (lane + 32 * k < g.vos_w) ? (1u << k) : 0u;

    // ---- core band from a once-clipped sample: VSTD_SAMPLE_ROWS rows spread over the channel,
    // one warp per row, combined in warp order (every CTA of the cluster computes the same numbers)
    {
        __shared__ double s_smp[VSTD_SAMPLE_ROWS][3];
        __shared__ double s_lohi[2];
        float xs[VSTD_MAXV];
        bool ok[VSTD_MAXV];
        if (warp < VSTD_SAMPLE_ROWS) {
            const int trow = (int)(((long long)g.dy * (2 * warp + 1)) / (2 * VSTD_SAMPLE_ROWS));
            T v[VSTD_MAXV];
            load_row(trow, v);
            const double fv = fitrow[trow];
            const bool in_hos = trow >= hos_t0 && trow < hos_t0 + g.hos_rows;
#pragma unroll
            for (int k = 0; k < VSTD_MAXV; k++) {
                xs[k] = value(v[k], fv, in_hos);
                ok[k] = ((kmask >> k) & 1u) && vos_valid(xs[k]);
            }
        }
        double lo = -INFINITY, hi = INFINITY;
        for (int pass = 0; pass < 2; pass++) {
            if (warp < VSTD_SAMPLE_ROWS) {
                double sm = 0.0, qm = 0.0, nm = 0.0;
#pragma unroll
                for (int k = 0; k < VSTD_MAXV; k++) {
                    const double xd = (double)xs[k];
                    if (ok[k] && xd >= lo && xd <= hi) { nm += 1.0; sm += xd; qm += xd * xd; }
                }
                sm = warp_sum(sm); qm = warp_sum(qm); nm = warp_sum(nm);
                if (lane == 0) { s_smp[warp][0] = nm; s_smp[warp][1] = sm; s_smp[warp][2] = qm; }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double nm = 0.0, sm = 0.0, qm = 0.0;
                for (int t = 0; t < VSTD_SAMPLE_ROWS; t++) { nm += s_smp[t][0]; sm += s_smp[t][1]; qm += s_smp[t][2]; }
                const double mean = nm > 0.0 ? sm / nm : 0.0;
                const double sd = nm > 0.0 ? sqrt(fmax(qm / nm - mean * mean, 0.0)) : -1.0;
                if (pass == 0) { s_lohi[0] = mean - 3.0 * sd; s_lohi[1] = mean + 3.0 * sd; }
                else {
                    // empty band (clo > chi) if the sample is unusable: every value goes to the lists
                    s_band[0] = sd > 0.0 ? (float)(mean - VSTD_CORE_SIGMA * sd) : 1.0f;
                    s_band[1] = sd > 0.0 ? (float)(mean + VSTD_CORE_SIGMA * sd) : 0.0f;
                }
            }
            __syncthreads();
            lo = s_lohi[0]; hi = s_lohi[1];
        }
    }
    clo = s_band[0]; chi = s_band[1];
    VSTD_TICK();

    // The walk over this CTA's rows.  BUILD (the first walk): moments of the values inside the
    // band, every other valid value into the list -- the moments of ALL valid values are then
    // "band + whole list".  Otherwise: moments of the valid values inside [lo, hi].
    auto strip_pass = [&](double lo, double hi, bool closed_nan_ok, bool build, Mom &m) {
        double s = 0.0, q = 0.0;
        int n = 0;
        int ovf = 0;
        T v[VSTD_MAXV], vn[VSTD_MAXV];
        if (row0 + warp < row1) load_row(row0 + warp, v);
        for (int trow = row0 + warp; trow < row1; trow += nwarps) {
            if (trow + nwarps < row1) load_row(trow + nwarps, vn);     // next row in flight during the arithmetic
            const double fv = fitrow[trow];
            const bool in_hos = trow >= hos_t0 && trow < hos_t0 + g.hos_rows;
#pragma unroll
            for (int k = 0; k < VSTD_MAXV; k++) {
                const float x = value(v[k], fv, in_hos);
                const double xd = (double)x;
                const float ax = fabsf(x);
                const bool valid = ((kmask >> k) & 1u) && ax > MASKED_ZERO_TOL && ax <= 3.402823466e+38f;
                if (build) {
                    const bool core = valid && x >= clo && x <= chi;
                    const double xm = core ? xd : 0.0;
                    n += core; s += xm; q = fma(xm, xm, q);
                    const unsigned int ballot = __ballot_sync(0xffffffffu, valid && !core);
                    if (ballot) {
                        const int pos = nlist + __popc(ballot & ((1u << lane) - 1u));
                        if (valid && !core) { if (pos < VSTD_LIST_PER_WARP) mylist[pos] = x; else ovf = 1; }
                        nlist += __popc(ballot);
                    }
                } else {
                    const bool in = valid && (closed_nan_ok ? (!(xd < lo) && !(xd > hi)) : (xd >= lo && xd <= hi));
                    const double xm = in ? xd : 0.0;
                    n += in; s += xm; q = fma(xm, xm, q);
                }
            }
#pragma unroll
            for (int k = 0; k < VSTD_MAXV; k++) v[k] = vn[k];
        }
        VSTD_TICK();
        if (build) exchange(0.0, 0.0, 0, s, q, (long long)n, ovf, m, true);
        else exchange(s, q, (long long)n, 0.0, 0.0, 0, 0, m, false);
        VSTD_TICK();
    };

    // band moments + the listed values inside [lo, hi]
    auto list_pass = [&](double lo, double hi, bool closed_nan_ok, Mom &m) {
        double s = 0.0, q = 0.0;
        long long n = 0;
        for (int j = lane; j < nlist; j += 32) {
            const double xd = (double)mylist[j];
            const bool in = closed_nan_ok ? (!(xd < lo) && !(xd > hi)) : (xd >= lo && xd <= hi);
            if (in) { n++; s += xd; q = fma(xd, xd, q); }
        }
        exchange(s, q, n, 0.0, 0.0, 0, 0, m, false);
        m.S += CS; m.Q += CQ; m.N += CN;
        VSTD_TICK();
    };

    auto eval = [&](double lo, double hi, bool closed_nan_ok, Mom &m) {
        if (use_list && clo <= chi && lo <= (double)clo && hi >= (double)chi) list_pass(lo, hi, closed_nan_ok, m);
        else if (use_list && clo > chi) list_pass(lo, hi, closed_nan_ok, m);      // empty band: the lists hold everything
        else strip_pass(lo, hi, closed_nan_ok, false, m);
    };

    double LO = -INFINITY, HI = INFINITY, flo = NAN, fhi = NAN;
    Mom m;
    strip_pass(LO, HI, false, true, m);       // band moments + lists
    eval(LO, HI, false, m);                   // all valid values: band + whole lists (or a second walk)
    for (int it = 0; it < 5 && m.N > 0; it++) {
        const double mean = m.S / (double)m.N;
        const double var = fmax(m.Q / (double)m.N - mean * mean, 0.0);
        const double sd = sqrt(var);
        flo = mean - 3.0 * sd;
        fhi = mean + 3.0 * sd;
        LO = fmax(LO, flo);
        HI = fmin(HI, fhi);
        Mom m2;
        eval(LO, HI, false, m2);
        const bool done = (m2.N == m.N);
        m = m2;
        if (done) break;
    }
    // final: population std of everything inside the final bounds (not the intersection)
    eval(flo, fhi, true, m);
    double result = NAN;
    if (m.N > 0) {
        const double mean = m.S / (double)m.N;
        result = sqrt(fmax(m.Q / (double)m.N - mean * mean, 0.0));
    }
    if (rank == 0 && threadIdx.x == 0) out_std[ch] = result;
    cluster.sync();     // keep every CTA's shared memory alive until all remote reads are done
    VSTD_TICK();
}

// ============================================================================================
// F2: overscan vector from the horizontal-overscan column statistics
// ============================================================================================
#define HOS_IDX_SWITCH 150
#define HOS_OVERLAP 30

__global__ void __launch_bounds__(256)
hos_fit_kernel(const float *__restrict__ hos_mean, const float *__restrict__ hos_std,
               const int32_t *__restrict__ hos_n, const uint8_t *__restrict__ satcol,
               bbx_geom g, int tel_kind, int split_chan, int split_col,
               double *__restrict__ out_oscan, uint8_t *__restrict__ out_need,
               int32_t *__restrict__ out_status)
{
    extern __shared__ double smem_d[];
    const int ch = blockIdx.x, NC = g.xsize_chan;
    double *mean = smem_d, *thr = mean + NC, *pk = thr + NC, *pkm1 = pk + NC, *fit = pkm1 + NC,
           *fit2 = fit + NC;
    uint8_t *valid = (uint8_t *)(fit2 + NC), *vp = valid + NC, *mf = vp + NC;
    __shared__ double scr_d[33];
    __shared__ int scr_i[33];
    __shared__ double rec[3 * (BBX_MAX_POLY_DEG + 1)];
    int status = 0;

    for (int x = threadIdx.x; x < NC; x += blockDim.x) {
        const size_t o = (size_t)ch * NC + x;
        const int n = hos_n[o];
        const bool v = n > 1;
        float err = 0.f;
        if (v) err = (float)((double)hos_std[o] / sqrt((double)n));
        mean[x] = (double)hos_mean[o];
        thr[x] = (double)(3.0f * err);
        valid[x] = v;
        vp[x] = v && x >= HOS_IDX_SWITCH - HOS_OVERLAP;
    }
    __syncthreads();

    // 5-sigma pre-clean of the columns entering the polynomial fit
    double m0, sd0, flo, fhi;
    block_clip_stats([&](int i, double &x) { x = mean[i]; return vp[i] && isfinite(x); }, NC, 5.0, 5,
                     scr_d, scr_i, m0, sd0, flo, fhi);
    for (int x = threadIdx.x; x < NC; x += blockDim.x)
        if (vp[x] && sd0 != 0.0) vp[x] = fabs(mean[x] - m0) / sd0 <= 5.0;
    __syncthreads();

    double tc, th;
    auto fit3 = [&](uint8_t *mask, int deg, double *dst) {
        for (int it = 0; it < 3; it++) {
            const int npts = block_polyfit(mean, mask, NC, 1.0, deg, dst, pk, pkm1, rec, scr_d, scr_i, tc, th);
            if (npts <= deg) { status = 1; return; }
            for (int x = threadIdx.x; x < NC; x += blockDim.x)
                if (mask[x]) mask[x] = fabs(dst[x] - mean[x]) <= thr[x];
            __syncthreads();
        }
    };
    if (ch != split_chan) {
        fit3(vp, 7, fit);
    } else {
        for (int x = threadIdx.x; x < NC; x += blockDim.x) mf[x] = vp[x] && x < split_col;
        __syncthreads();
        fit3(mf, 5, fit);
        for (int x = threadIdx.x; x < NC; x += blockDim.x) mf[x] = vp[x] && x >= split_col;
        __syncthreads();
        fit3(mf, 5, fit2);
        for (int x = threadIdx.x; x < NC; x += blockDim.x) if (x >= split_col) fit[x] = fit2[x];
        __syncthreads();
    }

    for (int x = threadIdx.x; x < NC; x += blockDim.x) {
        const size_t o = (size_t)ch * NC + x;
        double v = status ? NAN : fit[x];
        uint8_t need = 0;
        if (x < HOS_IDX_SWITCH) {
            bool use_mean = valid[x];
            if (tel_kind == BBX_TEL_BG) use_mean = use_mean && !satcol[o];
            if (use_mean || (x < 3 && valid[x])) v = mean[x];
            else { v = NAN; need = 1; }
        }
        out_oscan[o] = v;
        out_need[o] = need;
    }
    if (threadIdx.x == 0) out_status[ch] = status;
}

// ============================================================================================
// host entry points
// ============================================================================================
static void fill_chan_f32(ChanF32 &d, const float *s) { for (int i = 0; i < BBX_NCHAN; i++) d.v[i] = s ? s[i] : 1.0f; }
static void fill_chan_f64(ChanF64 &d, const double *s) { for (int i = 0; i < BBX_NCHAN; i++) d.v[i] = s ? s[i] : 0.0; }

static int check_geom(const bbx_geom *g, const char *who)
{
    BBX_REQUIRE(g != nullptr, "%s: geometry is null", who);
    BBX_REQUIRE(g->ny == 2 && g->ny * g->nx == BBX_NCHAN, "%s: expected 2 x 8 channels, got %d x %d", who, g->ny, g->nx);
    BBX_REQUIRE(g->dy > 0 && g->dx > 0 && g->H == g->ny * g->dy && g->W == g->nx * g->dx,
                "%s: frame %d x %d is not %d x %d tiles of %d x %d", who, g->H, g->W, g->ny, g->nx, g->dy, g->dx);
    BBX_REQUIRE(g->ysize_chan > 0 && g->ysize_chan <= g->dy && g->xsize_chan > 0 && g->xsize_chan <= g->dx,
                "%s: data section %d x %d does not fit the tile", who, g->ysize_chan, g->xsize_chan);
    return 0;
}

extern "C" int bbx_vos_rowstats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                                double sigma, int maxiters, double *out_mean, void *stream)
{
    if (check_geom(g, "bbx_vos_rowstats")) return -1;
    BBX_REQUIRE(g->vos_w > 0 && g->vos_w <= 32 * VOS_MAXV, "bbx_vos_rowstats: strip width %d not in 1..%d", g->vos_w, 32 * VOS_MAXV);
    BBX_REQUIRE(g->vos_x0 + g->vos_w <= g->dx, "bbx_vos_rowstats: strip exceeds the channel tile");
    ChanF32 gn; fill_chan_f32(gn, gain_h);
    const int warps = BBX_NCHAN * g->dy, blocks = ceil_div((long long)warps * 32, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (raw_type == BBX_RAW_U16)
        vos_rowstats_kernel<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t *)raw, *g, gn, sigma, maxiters, out_mean);
    else
        vos_rowstats_kernel<float><<<blocks, 256, 0, s>>>((const float *)raw, *g, gn, sigma, maxiters, out_mean);
    BBX_CHECK_LAUNCH("bbx_vos_rowstats");
    return 0;
}

extern "C" int bbx_vos_fit(const double *mean_vos, const bbx_geom *g, int deg, double nsigma,
                           double *out_fit, double *out_coef, double *out_biasm, int32_t *out_ok, void *stream)
{
    if (check_geom(g, "bbx_vos_fit")) return -1;
    BBX_REQUIRE(deg >= 0 && deg <= BBX_MAX_POLY_DEG, "bbx_vos_fit: degree %d not in 0..%d", deg, BBX_MAX_POLY_DEG);
    const size_t smem = (size_t)g->dy * (4 * sizeof(double) + 1) + 16;
    BBX_REQUIRE(smem <= 225 * 1024, "bbx_vos_fit: %d rows need %zu bytes of shared memory", g->dy, smem);
    BBX_CUDA(cudaFuncSetAttribute(vos_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vos_fit_kernel<<<BBX_NCHAN, 256, smem, (cudaStream_t)stream>>>(mean_vos, *g, deg, nsigma, out_fit, out_coef, out_biasm, out_ok);
    BBX_CHECK_LAUNCH("bbx_vos_fit");
    return 0;
}

extern "C" int bbx_hos_satcount(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                                const double *vos_fit, const double *sat_e_h, int lim1, int lim2,
                                int32_t *out_cnt, void *stream)
{
    if (check_geom(g, "bbx_hos_satcount")) return -1;
    BBX_REQUIRE(lim1 >= 0 && lim2 >= lim1, "bbx_hos_satcount: row limits %d, %d", lim1, lim2);
    BBX_REQUIRE(lim2 <= g->ysize_chan, "bbx_hos_satcount: row window %d exceeds the %d data rows of a channel", lim2, g->ysize_chan);
    ChanF32 gn; fill_chan_f32(gn, gain_h);
    ChanF64 se; fill_chan_f64(se, sat_e_h);
    cudaStream_t s = (cudaStream_t)stream;
    BBX_CUDA(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t) * BBX_NCHAN * 2 * (size_t)g->xsize_chan, s));
    if (lim2 == 0) return 0;
    bool gain_pos = true;
    for (int i = 0; i < BBX_NCHAN; i++) gain_pos = gain_pos && (gn.v[i] > 0.0f) && (gn.v[i] < INFINITY);
    if (raw_type == BBX_RAW_U16 && gain_pos && g->xsize_chan % 4 == 0 && g->dx % 4 == 0 && g->W % 4 == 0 &&
        ((uintptr_t)raw & 7) == 0) {
        dim3 grid4(ceil_div(g->xsize_chan / 4, SATC_THREADS), ceil_div(lim2, SATC_ROWS), BBX_NCHAN);
        hos_satcount_u16_kernel<<<grid4, SATC_THREADS, 0, s>>>((const uint16_t *)raw, *g, gn, vos_fit, se, lim1, lim2, out_cnt);
        BBX_CHECK_LAUNCH("bbx_hos_satcount");
        return 0;
    }
    const int rows_per_block = 64;
    dim3 grid(ceil_div(g->xsize_chan, 256), ceil_div(lim2, rows_per_block), BBX_NCHAN);
    if (raw_type == BBX_RAW_U16)
        hos_satcount_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t *)raw, *g, gn, vos_fit, se, lim1, lim2, rows_per_block, out_cnt);
    else
        hos_satcount_kernel<float><<<grid, 256, 0, s>>>((const float *)raw, *g, gn, vos_fit, se, lim1, lim2, rows_per_block, out_cnt);
    BBX_CHECK_LAUNCH("bbx_hos_satcount");
    return 0;
}

extern "C" int bbx_hos_stats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                             const double *vos_fit, int tel_kind, float data_limit, const int32_t *satcnt,
                             double *out_dlevel, float *out_mean, float *out_std, int32_t *out_n,
                             uint8_t *out_satcol, void *stream)
{
    if (check_geom(g, "bbx_hos_stats")) return -1;
    BBX_REQUIRE(g->hos_rows > 0 && g->hos_rows <= HOS_MAXR, "bbx_hos_stats: %d overscan rows not in 1..%d", g->hos_rows, HOS_MAXR);
    BBX_REQUIRE(tel_kind == BBX_TEL_ML || satcnt != nullptr, "bbx_hos_stats: BlackGEM masking needs the saturated-column counts");
    const size_t cells = (size_t)g->hos_rows * g->xsize_chan;
    const size_t smem = cells * (sizeof(float) + 2) + g->xsize_chan + 16;
    BBX_REQUIRE(smem <= 225 * 1024, "bbx_hos_stats: strip of %d x %d needs %zu bytes of shared memory", g->hos_rows, g->xsize_chan, smem);
    ChanF32 gn; fill_chan_f32(gn, gain_h);
    cudaStream_t s = (cudaStream_t)stream;
    if (raw_type == BBX_RAW_U16) {
        BBX_CUDA(cudaFuncSetAttribute(hos_stats_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hos_stats_kernel<uint16_t><<<BBX_NCHAN, 256, smem, s>>>((const uint16_t *)raw, *g, gn, vos_fit, tel_kind, data_limit, satcnt,
                                                              out_dlevel, out_mean, out_std, out_n, out_satcol);
    } else {
        BBX_CUDA(cudaFuncSetAttribute(hos_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hos_stats_kernel<float><<<BBX_NCHAN, 256, smem, s>>>((const float *)raw, *g, gn, vos_fit, tel_kind, data_limit, satcnt,
                                                           out_dlevel, out_mean, out_std, out_n, out_satcol);
    }
    BBX_CHECK_LAUNCH("bbx_hos_stats");
    return 0;
}

extern "C" int bbx_vos_std(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                           const double *vos_fit, const double *dlevel, double *out_std, void *stream)
{
    if (check_geom(g, "bbx_vos_std")) return -1;
    BBX_REQUIRE(g->vos_w > 0 && g->vos_w <= 32 * VSTD_MAXV, "bbx_vos_std: strip width %d not in 1..%d", g->vos_w, 32 * VSTD_MAXV);
    ChanF32 gn; fill_chan_f32(gn, gain_h);
    dim3 grid(VSTD_CLUSTER, BBX_NCHAN);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (VSTD_THREADS / 32) * VSTD_LIST_PER_WARP;
    if (raw_type == BBX_RAW_U16) {
        BBX_CUDA(cudaFuncSetAttribute(vos_std_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vos_std_kernel<uint16_t><<<grid, VSTD_THREADS, smem, s>>>((const uint16_t *)raw, *g, gn, vos_fit, dlevel, out_std);
    } else {
        BBX_CUDA(cudaFuncSetAttribute(vos_std_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vos_std_kernel<float><<<grid, VSTD_THREADS, smem, s>>>((const float *)raw, *g, gn, vos_fit, dlevel, out_std);
    }
    BBX_CHECK_LAUNCH("bbx_vos_std");
    return 0;
}

extern "C" int bbx_hos_fit(const float *hos_mean, const float *hos_std, const int32_t *hos_n,
                           const uint8_t *satcol, const bbx_geom *g, int tel_kind, int split_chan, int split_col,
                           double *out_oscan, uint8_t *out_need_spline, int32_t *out_status, void *stream)
{
    if (check_geom(g, "bbx_hos_fit")) return -1;
    const size_t smem = (size_t)g->xsize_chan * (6 * sizeof(double) + 3) + 16;
    BBX_REQUIRE(smem <= 225 * 1024, "bbx_hos_fit: %d columns need %zu bytes of shared memory", g->xsize_chan, smem);
    BBX_CUDA(cudaFuncSetAttribute(hos_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hos_fit_kernel<<<BBX_NCHAN, 256, smem, (cudaStream_t)stream>>>(hos_mean, hos_std, hos_n, satcol, *g, tel_kind, split_chan, split_col,
                                                                  out_oscan, out_need_spline, out_status);
    BBX_CHECK_LAUNCH("bbx_hos_fit");
    return 0;
}
