// xtalk.cu -- 16x16-channel crosstalk correction (reference: xtalk_corr, blackbox.py:7138-7258)
//
//   S_s      = data_s * ((data_s > 0) & !bad & !cosmic)                    float32
//   corr_v   = sum_{s same CCD half} c[s][v] S_s(y, x)  +  sum_{s other half} c[s][v] S_s(H-1-y, x)
//              (two 8-term float64 dot products, added; np.matmul of f32 x f64 -> f64)
//   data_v   = f32( f64(data_v) - corr_v * !edge )
// All sources are the UNCORRECTED pixels (the reference builds the correction cube first).
//
// A tile position (ly, lx) in all 16 channels -- the 8 bottom channels at row ly and the 8 top
// channels at the mirrored row -- is exactly a set of mutual sources, so every pixel of the
// frame is read once and written once (4+1 B read, 4 B written per pixel; 16 float64 FMAs per
// pixel: the FP64 pipe needs ~130 us per 10560^2 frame, HBM ~155 us).
//
// xtalk_tile_kernel: a CTA stages 128 tile positions x 16 channels in shared memory with
// 128-bit coalesced loads, each thread then corrects one position in all 16 channels out of
// shared memory (16 + 4 live doubles: ~64 registers, 6+ CTAs per SM hide the loads of the other
// CTAs behind the FP64 work), and the tile goes back with 128-bit stores.
// xtalk_kernel<PX>: generic fallback (any width / alignment), registers only.
#include "bbx_common.cuh"

struct XtalkCoef { double c[16][16]; };   // [victim][source] (the order the dot products walk), kernel parameter (constant bank)

// PX = 4 (float4 + 32-bit mask words), 2 or 1 pixels per thread and channel.  Only the pixel
// values and two 64-bit flag words are kept in registers; the masked source value
// S = v * (ok ? 1 : 0) is formed on the fly.
template <int PX>
__global__ void __launch_bounds__(128)
xtalk_kernel(float *img, const uint8_t *__restrict__ mask, int W, int ysc, int xsc, XtalkCoef k,
             uint32_t bits_src_bad, uint32_t bit_edge)
{
    const int groups = xsc / PX;
    const long long total = (long long)ysc * groups;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ly = (int)(t / groups), lx = (int)(t - (long long)ly * groups) * PX;
        float v[16][PX];
        unsigned long long src_ok = 0, vic_ok = 0;           // bit (c * PX + p)
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
            const size_t off = (size_t)row * W + (size_t)(c & 7) * xsc + lx;
            uint32_t mm = 0;
            if (PX == 4) {
                const float4 f = *reinterpret_cast<const float4 *>(img + off);
                v[c][0] = f.x; v[c][PX > 1 ? 1 : 0] = f.y; v[c][PX > 2 ? 2 : 0] = f.z; v[c][PX > 3 ? 3 : 0] = f.w;
                if (mask) mm = *reinterpret_cast<const uint32_t *>(mask + off);
            } else if (PX == 2) {
                const float2 f = *reinterpret_cast<const float2 *>(img + off);
                v[c][0] = f.x; v[c][PX > 1 ? 1 : 0] = f.y;
                if (mask) mm = *reinterpret_cast<const uint16_t *>(mask + off);
            } else {
                v[c][0] = img[off];
                if (mask) mm = mask[off];
            }
#pragma unroll
            for (int p = 0; p < PX; p++) {
                const uint32_t m = (mm >> (8 * p)) & 0xffu;
                if ((v[c][p] > 0.0f) && !(m & bits_src_bad)) src_ok |= 1ull << (c * PX + p);
                if (!(m & bit_edge)) vic_ok |= 1ull << (c * PX + p);
            }
        }
        // pixel by pixel: the 16 masked sources as float64, then the 16 victims; the outputs
        // overwrite v[.][p] once its sources have been captured (keeps the register count low)
#pragma unroll
        for (int p = 0; p < PX; p++) {
            double S[16];
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const float sf = v[c][p] * (((src_ok >> (c * PX + p)) & 1ull) ? 1.0f : 0.0f);
                S[c] = (double)sf;
            }
#pragma unroll
            for (int vch = 0; vch < 16; vch++) {
                const int same0 = (vch < 8) ? 0 : 8, other0 = 8 - same0;
                // same-half sources first (quadrant q=0 / q=3), then the mirrored half
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int s = 0; s < 8; s++) a = fma(S[same0 + s], k.c[vch][same0 + s], a);
#pragma unroll
                for (int s = 0; s < 8; s++) b = fma(S[other0 + s], k.c[vch][other0 + s], b);
                double corr = 0.0 + a;
                corr = corr + b;
                corr = corr * (((vic_ok >> (vch * PX + p)) & 1ull) ? 1.0 : 0.0);
                v[vch][p] = (float)((double)v[vch][p] - corr);
            }
        }
#pragma unroll
        for (int vch = 0; vch < 16; vch++) {
            const int row = (vch < 8) ? ly : (ysc + (ysc - 1 - ly));
            const size_t off = (size_t)row * W + (size_t)(vch & 7) * xsc + lx;
            if (PX == 4) *reinterpret_cast<float4 *>(img + off) = make_float4(v[vch][0], v[vch][PX > 1 ? 1 : 0], v[vch][PX > 2 ? 2 : 0], v[vch][PX > 3 ? 3 : 0]);
            else if (PX == 2) *reinterpret_cast<float2 *>(img + off) = make_float2(v[vch][0], v[vch][PX > 1 ? 1 : 0]);
            else img[off] = v[vch][0];
        }
    }
}

#define XT_TILE 128          // tile positions per CTA pass (= threads per CTA)

template <int MINB>
__global__ void __launch_bounds__(XT_TILE, MINB)
xtalk_tile_kernel(float *img, const uint8_t *__restrict__ mask, int W, int ysc, int xsc, XtalkCoef k,
                  uint32_t bits_src_bad, uint32_t bit_edge)
{
    __shared__ __align__(16) float tile[16][XT_TILE];
    __shared__ __align__(16) uint8_t mt[16][XT_TILE];
    const int gpr = xsc / 4;                                   // float4 groups per channel row
    const long long ngroups = (long long)ysc * gpr;
    const long long ntiles = (ngroups + XT_TILE / 4 - 1) / (XT_TILE / 4);
    const int g4 = threadIdx.x & 31, c0 = threadIdx.x >> 5;    // load role: group in tile, first channel
    // one tile per CTA (no tile loop: the compiler would hoist the 256 loop-invariant coefficient
    // loads out of it into registers; this way they stay LDCU.128 next to their DFMAs)
    {
        const long long t = blockIdx.x;
        if (t >= ntiles) return;
        // ---- stage: thread (c0, g4) moves group g4 of channels c0, c0+4, c0+8, c0+12
        const long long G = t * (XT_TILE / 4) + g4;
        const bool live = G < ngroups;
        const int ly = live ? (int)(G / gpr) : 0, lx = live ? (int)(G - (long long)ly * gpr) * 4 : 0;
        size_t off[4];
        float4 f[4];
        uint32_t mm[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = c0 + 4 * i;
            const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
            off[i] = (size_t)row * W + (size_t)(c & 7) * xsc + lx;
            f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            mm[i] = 0;
            if (live) {
                f[i] = *reinterpret_cast<const float4 *>(img + off[i]);
                if (mask) mm[i] = *reinterpret_cast<const uint32_t *>(mask + off[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = c0 + 4 * i;
            *reinterpret_cast<float4 *>(&tile[c][g4 * 4]) = f[i];
            *reinterpret_cast<uint32_t *>(&mt[c][g4 * 4]) = mm[i];
        }
        __syncthreads();
        // ---- correct: thread p owns tile position p in all 16 channels
        {
            const int p = threadIdx.x;
            double S[16];
            uint32_t vic_ok = 0;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const float v = tile[c][p];
                const uint32_t m = mt[c][p];
                const bool ok = (v > 0.0f) && !(m & bits_src_bad);
                const float sf = v * (ok ? 1.0f : 0.0f);
                S[c] = (double)sf;
                if (!(m & bit_edge)) vic_ok |= 1u << c;
            }
#pragma unroll
            for (int vch = 0; vch < 16; vch++) {
                const int same0 = (vch < 8) ? 0 : 8, other0 = 8 - same0;
                // same-half sources first (quadrant q=0 / q=3), then the mirrored half
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int s = 0; s < 8; s++) a = fma(S[same0 + s], k.c[vch][same0 + s], a);
#pragma unroll
                for (int s = 0; s < 8; s++) b = fma(S[other0 + s], k.c[vch][other0 + s], b);
                double corr = 0.0 + a;
                corr = corr + b;
                corr = corr * (((vic_ok >> vch) & 1u) ? 1.0 : 0.0);
                tile[vch][p] = (float)((double)tile[vch][p] - corr);
            }
        }
        __syncthreads();
        // ---- write back
        if (live) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = c0 + 4 * i;
                *reinterpret_cast<float4 *>(img + off[i]) = *reinterpret_cast<const float4 *>(&tile[c][g4 * 4]);
            }
        }
    }
}

extern "C" int bbx_xtalk(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                         const double *coeffs_h, const bbx_maskbits *bits, void *stream)
{
    return bbx_xtalk_variant(img, mask, H, W, ysize_chan, xsize_chan, coeffs_h, bits, 0, stream);
}

// variant 0: the kernel bbx_xtalk picks (tiled when the layout allows); 1 / 2 / 4: the generic
// register-only kernel with that many pixels per thread and channel (parity tests, tools/xt_bench.py)
extern "C" int bbx_xtalk_variant(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                 const double *coeffs_h, const bbx_maskbits *bits, int variant, void *stream)
{
    BBX_REQUIRE(img && coeffs_h && bits, "bbx_xtalk: null argument");
    BBX_REQUIRE(variant == 0 || variant == 1 || variant == 2 || variant == 4, "bbx_xtalk: variant %d", variant);
    BBX_REQUIRE(H == 2 * ysize_chan && W == 8 * xsize_chan, "bbx_xtalk: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    XtalkCoef k;
    for (int s = 0; s < 16; s++) for (int v = 0; v < 16; v++) k.c[v][s] = coeffs_h[s * 16 + v];
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t src_bad = (uint32_t)(bits->bad | bits->cosmic);
    const bool px4 = (xsize_chan % 4 == 0) && ((uintptr_t)img % 16) == 0 && ((uintptr_t)mask % 4) == 0;
    if (px4 && variant == 0) {
        const long long ngroups = (long long)ysize_chan * (xsize_chan / 4);
        const long long ntiles = (ngroups + XT_TILE / 4 - 1) / (XT_TILE / 4);
        BBX_REQUIRE(ntiles < 2147483647LL, "bbx_xtalk: frame too large");
        const int blocks = (int)ntiles;
        // register allocation capped for 6 resident CTAs per SM (80 registers, no spills; measured
        // on B200: 5 -> 0.266 ms, 6 -> 0.248 ms, 7 -> 0.249 ms, 8 -> 0.252 ms with spills)
        xtalk_tile_kernel<6><<<blocks, XT_TILE, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad, (uint32_t)bits->edge);
        BBX_CHECK_LAUNCH("xtalk_tile_kernel");
        return 0;
    }
    const bool px2 = (xsize_chan % 2 == 0) && ((uintptr_t)img % 8) == 0 && ((uintptr_t)mask % 2) == 0;
    // measured on B200 (10560^2, tools/xt_bench.py): 2 px/thread 0.395 ms, 4 px/thread 0.435 ms
    // (255 registers), 1 px/thread 0.76 ms
    int px = px2 ? 2 : 1;
    if (variant == 4 && px4) px = 4;
    if (variant == 1 || (variant == 2 && !px2)) px = 1;
    const long long total = (long long)ysize_chan * (xsize_chan / px);
    long long want = (total + 127) / 128;
    const int blocks = (int)(want < BBX_SM_COUNT * 16 ? want : BBX_SM_COUNT * 16);
    if (px == 4) xtalk_kernel<4><<<blocks, 128, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad, (uint32_t)bits->edge);
    else if (px == 2) xtalk_kernel<2><<<blocks, 128, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad, (uint32_t)bits->edge);
    else xtalk_kernel<1><<<blocks, 128, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad, (uint32_t)bits->edge);
    BBX_CHECK_LAUNCH("xtalk_kernel");
    return 0;
}
