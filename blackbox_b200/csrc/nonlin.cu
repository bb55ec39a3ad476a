// nonlin.cu -- non-linearity correction (reference: nonlin_corr, blackbox.py:7392-7437; switched
// off in the reference's settings).  Per channel a spline s(counts) gives the
// fractional deviation from linearity:
//     counts = data / gain                      (float32; counts, not electrons)
//     frac   = counts <= 50000 ? s(counts) : 1  (float64; the reference initialises frac_corr with
//                                                ones, so pixels above the limit -- and NaNs -- are
//                                                divided by 2: kept as is)
//     data   = f32( f64(data) / (frac + 1) )
// s is a FITPACK B-spline (scipy UnivariateSpline pickled by the reference): evaluated exactly as
// splev / fpbspl do (interval search with extrapolation from the end intervals, de Boor recurrence
// in this operation order, no FMA), so the float64 values equal scipy's bit for bit.
#include "bbx_common.cuh"

#define NL_MAXK 5
#define NL_MAXT 64            // knots per channel (scipy smoothing splines of such curves have ~10)

struct NonlinSplines {
    double t[BBX_NCHAN][NL_MAXT];
    double c[BBX_NCHAN][NL_MAXT];
    int n[BBX_NCHAN];         // number of knots
    int k[BBX_NCHAN];         // degree
};

__device__ __forceinline__ double nl_splev(const double *__restrict__ t, const double *__restrict__ c, int n, int k,
                                           double x)
{
    const int nk1 = n - k - 1;
    int l = k;                                           // t[l] <= x < t[l+1], clamped to [k, nk1 - 1]
    while (l < nk1 - 1 && x >= t[l + 1]) l++;
    double h[NL_MAXK + 1], hh[NL_MAXK + 1];
    h[0] = 1.0;
    for (int j = 1; j <= k; j++) {
        for (int i = 0; i < j; i++) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; i++) {
            const int li = l + i, lj = li - j;
            if (t[li] == t[lj]) { h[i] = 0.0; continue; }
            const double f = hh[i - 1] / (t[li] - t[lj]);
            h[i - 1] = h[i - 1] + f * (t[li] - x);
            h[i] = f * (x - t[lj]);
        }
    }
    double sp = 0.0;
    const int ll = l - k;
    for (int j = 0; j <= k; j++) sp = sp + c[ll + j] * h[j];
    return sp;
}

__global__ void __launch_bounds__(256)
nonlin_kernel(float *img, int W, int ysc, int xsc, ChanF32 gain, const __grid_constant__ NonlinSplines spl,
              float max_counts)
{
    __shared__ double s_t[NL_MAXT], s_c[NL_MAXT];
    const int ch = blockIdx.z, r = ch >> 3, cc = ch & 7;
    const int n = spl.n[ch], k = spl.k[ch];
    for (int i = threadIdx.x; i < n; i += blockDim.x) { s_t[i] = spl.t[ch][i]; s_c[i] = spl.c[ch][i]; }
    __syncthreads();
    const float g = gain.v[ch];
    const int y = blockIdx.y;
    float *row = img + (size_t)(r * ysc + y) * W + (size_t)cc * xsc;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < xsc; x += gridDim.x * blockDim.x) {
        const float d = row[x];
        const float counts = d / g;
        double frac = 1.0;
        if (counts <= max_counts) frac = nl_splev(s_t, s_c, n, k, (double)counts);
        row[x] = (float)((double)d / (frac + 1.0));
    }
}

extern "C" int bbx_nonlin_corr(float *img, int H, int W, int ysize_chan, int xsize_chan, const float *gain_h,
                               const double *knots_h, const double *coefs_h, const int *nknots_h,
                               const int *degree_h, int max_knots, float max_counts, void *stream)
{
    BBX_REQUIRE(img && gain_h && knots_h && coefs_h && nknots_h && degree_h, "bbx_nonlin_corr: null argument");
    BBX_REQUIRE(ysize_chan > 0 && xsize_chan > 0 && H == 2 * ysize_chan && W == 8 * xsize_chan,
                "bbx_nonlin_corr: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    NonlinSplines host;                                  // 16.5 KB, passed by value (kernel parameter space)
    for (int ch = 0; ch < BBX_NCHAN; ch++) {
        const int n = nknots_h[ch], k = degree_h[ch];
        BBX_REQUIRE(k >= 1 && k <= NL_MAXK, "bbx_nonlin_corr: channel %d: spline degree %d not in 1..%d", ch + 1, k, NL_MAXK);
        BBX_REQUIRE(n >= 2 * (k + 1) && n <= NL_MAXT && n <= max_knots,
                    "bbx_nonlin_corr: channel %d: %d knots (need %d..%d)", ch + 1, n, 2 * (k + 1), NL_MAXT);
        host.n[ch] = n; host.k[ch] = k;
        for (int i = 0; i < NL_MAXT; i++) {
            host.t[ch][i] = i < n ? knots_h[(size_t)ch * max_knots + i] : 0.0;
            host.c[ch][i] = i < n ? coefs_h[(size_t)ch * max_knots + i] : 0.0;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    ChanF32 gn;
    for (int i = 0; i < BBX_NCHAN; i++) gn.v[i] = gain_h[i];
    dim3 grid((unsigned int)ceil_div(xsize_chan, 256 * 2), (unsigned int)ysize_chan, BBX_NCHAN);
    nonlin_kernel<<<grid, 256, 0, st>>>(img, W, ysize_chan, xsize_chan, gn, host, max_counts);
    BBX_CHECK_LAUNCH("bbx_nonlin_corr");
    return 0;
}
