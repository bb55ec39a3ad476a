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
//
// xtalk_tma_kernel (variant 5; NOT the default -- measured on B200, 10560^2: 0.32 ms against the
// tile kernel's 0.227 ms, see below): the same tile walk as a PERSISTENT kernel with asynchronous
// staging.  The image is described to the TMA unit as a 3-D
// tensor (x within channel, channel column, row), so ONE cp.async.bulk.tensor box {BX, 8, 1} brings
// the 8 bottom channels of a tile row into shared memory and a second one the 8 mirrored top
// channels (8 KB per tile, no thread touches an address); a 4-stage mbarrier ring keeps the loads
// of tile k+2 in flight while the 256 DFMAs per thread of tile k run; the results leave through
// registers as full-line stores.  Measured (tools/xt_bench.py, profiles/r02_xtalk_variants.txt):
// 0.333 ms with the results stored back through the stage by cp.async.bulk.tensor, 0.317 ms with the
// mask through cp.async, 0.320 ms with register stores -- against 0.227 ms for the synchronous
// tile kernel.  The kernel is bound by instruction issue, not by load latency: ~900 instructions per
// tile position (256 DFMA + 144 LDCU for the coefficients + conversions) are 0.17 ms of issue time
// on 148 SMs, the FP64 pipe needs 0.10 ms, HBM 0.155 ms; the one-tile-per-CTA kernel with 6 CTAs
// per SM already overlaps its loads with other CTAs' DFMAs, and the persistent loop makes the
// compiler keep part of the coefficient matrix in registers (96 registers, 5 CTAs per SM).  Kept
// as a tested variant and as the record of the experiment.  The kernel also counts the pixels per mask bit on its way (the
// mask is final here: mask_header's M-*NUM, blackbox.py:4601-4620, cost no extra pass).
#include <cuda.h>
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

// COUNT: also count the pixels per mask bit (mask_header's M-*NUM, blackbox.py:4601-4620; the mask
// is final when the crosstalk correction runs, and this kernel sees every byte of it exactly once):
// per warp and only where a warp holds any mask bit at all, per CTA into shared memory, per CTA
// one atomic per non-zero bit into one of XT_SLOTS spread copies of the eight counters
#define XT_SLOTS 16
template <int MINB, bool COUNT>
__global__ void __launch_bounds__(XT_TILE, MINB)
xtalk_tile_kernel(float *img, const uint8_t *__restrict__ mask, int W, int ysc, int xsc, XtalkCoef k,
                  uint32_t bits_src_bad, uint32_t bit_edge, unsigned long long *__restrict__ slots)
{
    __shared__ __align__(16) float tile[16][XT_TILE];
    __shared__ __align__(16) uint8_t mt[16][XT_TILE];
    __shared__ unsigned int s_cnt[8];
    if (COUNT && threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
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
        if (COUNT) {
            const uint32_t any = mm[0] | mm[1] | mm[2] | mm[3];
            if (__any_sync(0xffffffffu, any != 0)) {
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const uint32_t pl = 0x01010101u << b;
                    unsigned int v = (unsigned int)(__popc(mm[0] & pl) + __popc(mm[1] & pl) + __popc(mm[2] & pl) + __popc(mm[3] & pl));
                    v = __reduce_add_sync(0xffffffffu, v);
                    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[b], v);
                }
            }
        }
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
        if (COUNT && threadIdx.x < 8 && s_cnt[threadIdx.x])
            atomicAdd(&slots[(blockIdx.x % XT_SLOTS) * 8 + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
}

__global__ void xtalk_count_sum_kernel(const unsigned long long *__restrict__ slots, unsigned long long *__restrict__ out)
{
    const int b = threadIdx.x;
    if (b >= 8) return;
    unsigned long long t = 0;
    for (int s = 0; s < XT_SLOTS; s++) t += slots[s * 8 + b];
    out[b] = t;
}

// --------------------------------------------------------------------------------------------
// TMA path
// --------------------------------------------------------------------------------------------
#define XT_STAGES 4
#define XT_THREADS 128

__device__ __forceinline__ uint32_t xt_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xt_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(xt_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void xt_mbar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(xt_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xt_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "XT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra XT_DONE;\n\t"
        "bra XT_WAIT;\n\t"
        "XT_DONE:\n\t}"
        :: "r"(xt_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void xt_tma_load(void *dst, const CUtensorMap *map, int x, int ch, int row, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(xt_smem(dst)), "l"(map), "r"(x), "r"(ch), "r"(row), "r"(xt_smem(bar)) : "memory");
}
__device__ __forceinline__ void xt_tma_store(const CUtensorMap *map, const void *src, int x, int ch, int row)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(map), "r"(xt_smem(src)), "r"(x), "r"(ch), "r"(row) : "memory");
}

// The kernel is bound by instruction issue (256 DFMAs per position plus everything around them:
// the synchronous tile kernel spends ~900 instructions per position, 0.17 ms of issue time on 148
// SMs), so the point of the TMA staging is as much the instructions it removes -- no address
// arithmetic, no loads, no stores in the threads -- as the latency it hides.
//   image: 3-D tensor map (x in channel, channel column, row), box {128, 8, 1}: two loads per tile
//   mask:  a channel is 1320 bytes wide -- neither a legal TMA stride nor (for odd channels) a
//          16-byte aligned box origin, which the TMA unit insists on (an unaligned origin is an
//          illegal-instruction fault, found the hard way) -- so the 2 KB of mask bytes of a tile
//          come with four 4-byte cp.async per thread, completing on the SAME mbarrier as the
//          image boxes (cp.async.mbarrier.arrive.noinc)
// A box that sticks out of its channel (the 11th of a 1320-wide channel) is clipped by the TMA
// unit; the threads beyond the channel edge sit the tile out.
#define XT_BX 128
struct XtStage { float img[16][XT_BX]; uint8_t msk[16][XT_BX]; };

__global__ void __launch_bounds__(XT_THREADS, 5)
xtalk_tma_kernel(const __grid_constant__ CUtensorMap map, float *__restrict__ img, const uint8_t *__restrict__ mask, int W,
                 int ysc, int xsc, XtalkCoef k, uint32_t bits_src_bad, uint32_t bit_edge,
                 unsigned long long *__restrict__ counts)
{
    extern __shared__ __align__(128) unsigned char xt_raw[];
    XtStage *stage = reinterpret_cast<XtStage *>(xt_raw);
    __shared__ __align__(8) uint64_t full[XT_STAGES];
    __shared__ unsigned int s_cnt[8];
    const int p = threadIdx.x;
    const int nxb = (xsc + XT_BX - 1) / XT_BX;
    const int ntiles = ysc * nxb;
    const bool have_mask = mask != nullptr;
    if (p == 0) {
        // arrivals per phase: thread 0's expect_tx for the two image boxes, plus (with a mask) one
        // per thread when its cp.async of the mask bytes have landed
#pragma unroll
        for (int s = 0; s < XT_STAGES; s++) xt_mbar_init(&full[s], have_mask ? 1 + XT_THREADS : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p < 8) s_cnt[p] = 0;
    __syncthreads();

    auto issue = [&](int ly, int xb, int s) {                    // tile (ly, xb) into stage s
        const int x0 = xb * XT_BX, top = ysc + (ysc - 1 - ly);
        if (p == 0) {
            xt_mbar_expect(&full[s], 16u * XT_BX * 4u);
            xt_tma_load(&stage[s].img[0][0], &map, x0, 0, ly, &full[s]);
            xt_tma_load(&stage[s].img[8][0], &map, x0, 0, top, &full[s]);
        }
        if (have_mask) {
            // 16 channels x 32 words: thread p moves word (p % 32) of channels p / 32, + 4, + 8, + 12
            const int wx = x0 + 4 * (p & 31);
            if (wx < xsc) {                                      // xsc % 4 == 0: the word lies inside the channel
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int c = (p >> 5) + 4 * i;
                    const uint8_t *src = mask + (size_t)((c < 8) ? ly : top) * W + (size_t)(c & 7) * xsc + wx;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                                 :: "r"(xt_smem(&stage[s].msk[c][4 * (p & 31)])), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(xt_smem(&full[s])) : "memory");
        }
    };
    // tile walk: t = blockIdx.x + i gridDim.x, kept as (row, box) without divisions in the loop
    const int step = gridDim.x, dly = step / nxb, dxb = step - dly * nxb;
    auto advance = [&](int &ly, int &xb) {
        ly += dly; xb += dxb;
        if (xb >= nxb) { xb -= nxb; ly++; }
    };
    int ly = (int)blockIdx.x / nxb, xb = (int)blockIdx.x - ly * nxb;       // current tile
    int ly2 = ly, xb2 = xb;                                               // the tile two ahead (next to issue)
    if (ly2 < ysc) issue(ly2, xb2, 0);
    advance(ly2, xb2);
    if (ly2 < ysc) issue(ly2, xb2, 1);
    advance(ly2, xb2);
    (void)ntiles;
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; ly < ysc; it++) {
        const int s = it % XT_STAGES;
        // the stage tile it+2 goes into was last read by tile it-2: one block barrier per pass
        // keeps the refill behind every thread's reads (the results leave through registers,
        // nothing reads a stage after its pass)
        __syncthreads();
        if (ly2 < ysc) issue(ly2, xb2, (it + 2) % XT_STAGES);
        advance(ly2, xb2);
        xt_mbar_wait(&full[s], (uint32_t)(it / XT_STAGES) & 1u);
        const int x0 = xb * XT_BX;
        if (x0 + p < xsc) {
            XtStage &st = stage[s];
            double S[16];
            uint32_t vic_ok = 0, mw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const float v = st.img[c][p];
                const uint32_t m = have_mask ? (uint32_t)st.msk[c][p] : 0u;
                mw[c >> 2] |= m << (8 * (c & 3));
                const bool ok = (v > 0.0f) && !(m & bits_src_bad);
                const float sf = v * (ok ? 1.0f : 0.0f);        // the reference's data * mask (NaN stays NaN)
                S[c] = (double)sf;
                if (!(m & bit_edge)) vic_ok |= 1u << c;
            }
#pragma unroll
            for (int vch = 0; vch < 16; vch++) {
                const int same0 = (vch < 8) ? 0 : 8, other0 = 8 - same0;
                // same-half sources first (quadrant q=0 / q=3), then the mirrored half
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int sidx = 0; sidx < 8; sidx++) a = fma(S[same0 + sidx], k.c[vch][same0 + sidx], a);
#pragma unroll
                for (int sidx = 0; sidx < 8; sidx++) b = fma(S[other0 + sidx], k.c[vch][other0 + sidx], b);
                double corr = 0.0 + a;
                corr = corr + b;
                corr = corr * (((vic_ok >> vch) & 1u) ? 1.0 : 0.0);
                const float res = (float)((double)st.img[vch][p] - corr);
                // 32 consecutive floats per warp and channel: a full 128-byte line per store
                img[(size_t)((vch < 8) ? ly : ysc + (ysc - 1 - ly)) * W + (size_t)(vch & 7) * xsc + x0 + p] = res;
            }
            if (counts && (mw[0] | mw[1] | mw[2] | mw[3])) {     // most positions carry no mask bit at all
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const uint32_t pl = 0x01010101u << b;
                    cnt[b] += __popc(mw[0] & pl) + __popc(mw[1] & pl) + __popc(mw[2] & pl) + __popc(mw[3] & pl);
                }
            }
        }
        advance(ly, xb);
    }
    if (counts) {
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int v = warp_sum(cnt[b]);
            if ((p & 31) == 0 && v) atomicAdd(&s_cnt[b], (unsigned int)v);
        }
        __syncthreads();
        if (p < 8 && s_cnt[p]) atomicAdd(&counts[p], (unsigned long long)s_cnt[p]);
    }
}

typedef CUresult (*xt_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (libbbx.so does not link libcuda)
static xt_encode_fn xt_encoder(void)
{
    static xt_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (xt_encode_fn)sym;
    }
    return fn;
}

// 0 = launched; 1 = this layout / driver cannot take the TMA path (caller falls back); < 0 = error
static int xtalk_tma_launch(float *img, const uint8_t *mask, int H, int W, int ysc, int xsc, const XtalkCoef &k,
                            uint32_t src_bad, uint32_t edge, unsigned long long *counts, cudaStream_t st)
{
    if (xsc % 4 != 0 || ((uintptr_t)img % 16) != 0 || ((uintptr_t)mask % 4) != 0 || W != 8 * xsc || H != 2 * ysc ||
        xsc < XT_BX)
        return 1;
    xt_encode_fn enc = xt_encoder();
    if (!enc) return 1;
    CUtensorMap map;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)xsc, 8, (cuuint64_t)H};
        const cuuint64_t strides[2] = {(cuuint64_t)xsc * 4, (cuuint64_t)W * 4};
        const cuuint32_t box[3] = {XT_BX, 8, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 1;
    }
    static bool attr_set = false;
    const int smem = (int)(XT_STAGES * sizeof(XtStage));
    if (!attr_set) {
        BBX_CUDA(cudaFuncSetAttribute(xtalk_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    if (counts) BBX_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(unsigned long long), st));
    const long long ntiles = (long long)ysc * ((xsc + XT_BX - 1) / XT_BX);
    long long want = (long long)BBX_SM_COUNT * 5;
    const int blocks = (int)(ntiles < want ? ntiles : want);
    xtalk_tma_kernel<<<blocks, XT_THREADS, smem, st>>>(map, img, mask, W, ysc, xsc, k, src_bad, edge, counts);
    BBX_CHECK_LAUNCH("xtalk_tma_kernel");
    return 0;
}

extern "C" int bbx_xtalk(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                         const double *coeffs_h, const bbx_maskbits *bits, void *stream)
{
    return bbx_xtalk_counts(img, mask, H, W, ysize_chan, xsize_chan, coeffs_h, bits, 0, nullptr, stream);
}

extern "C" int bbx_xtalk_variant(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                 const double *coeffs_h, const bbx_maskbits *bits, int variant, void *stream)
{
    return bbx_xtalk_counts(img, mask, H, W, ysize_chan, xsize_chan, coeffs_h, bits, variant, nullptr, stream);
}

// variant 0: the kernel bbx_xtalk picks (the tile kernel when the layout allows, else the generic
// one); 3: the tile kernel; 5: the TMA-staged persistent kernel (falls back to 0 where the layout
// does not allow it); 1 / 2 / 4: the generic register-only kernel with that many pixels per thread
// and channel (parity tests, tools/xt_bench.py).  out_counts (device uint64 [BBX_XTALK_COUNTS_LEN],
// may be null): [0:8] pixels per mask bit of `mask`, filled by the call -- by the crosstalk kernel
// itself on the tile and TMA paths, by bbx_mask_counts' kernel otherwise; the rest is scratch.
extern "C" int bbx_xtalk_counts(float *img, const uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                const double *coeffs_h, const bbx_maskbits *bits, int variant,
                                unsigned long long *out_counts, void *stream)
{
    BBX_REQUIRE(img && coeffs_h && bits, "bbx_xtalk: null argument");
    BBX_REQUIRE(variant >= 0 && variant <= 5, "bbx_xtalk: variant %d", variant);
    BBX_REQUIRE(H == 2 * ysize_chan && W == 8 * xsize_chan, "bbx_xtalk: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    BBX_REQUIRE(out_counts == nullptr || mask != nullptr, "bbx_xtalk: mask counts asked for without a mask");
    XtalkCoef k;
    for (int s = 0; s < 16; s++) for (int v = 0; v < 16; v++) k.c[v][s] = coeffs_h[s * 16 + v];
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t src_bad = (uint32_t)(bits->bad | bits->cosmic);
    if (variant == 5) {
        const int rc = xtalk_tma_launch(img, mask, H, W, ysize_chan, xsize_chan, k, src_bad, (uint32_t)bits->edge, out_counts, st);
        if (rc <= 0) return rc;
    }
    const bool px4 = (xsize_chan % 4 == 0) && ((uintptr_t)img % 16) == 0 && ((uintptr_t)mask % 4) == 0;
    if (px4 && (variant == 0 || variant == 3)) {
        const long long ngroups = (long long)ysize_chan * (xsize_chan / 4);
        const long long ntiles = (ngroups + XT_TILE / 4 - 1) / (XT_TILE / 4);
        BBX_REQUIRE(ntiles < 2147483647LL, "bbx_xtalk: frame too large");
        const int blocks = (int)ntiles;
        // register allocation capped for 6 resident CTAs per SM (80 registers, no spills; measured
        // on B200: 5 -> 0.266 ms, 6 -> 0.248 ms, 7 -> 0.249 ms, 8 -> 0.252 ms with spills)
        if (out_counts) {
            // the counts ride along: XT_SLOTS spread copies behind the eight results, summed by a 8-thread kernel
            BBX_CUDA(cudaMemsetAsync(out_counts, 0, (8 + 8 * XT_SLOTS) * sizeof(unsigned long long), st));
            xtalk_tile_kernel<6, true><<<blocks, XT_TILE, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad,
                                                                  (uint32_t)bits->edge, out_counts + 8);
            xtalk_count_sum_kernel<<<1, 32, 0, st>>>(out_counts + 8, out_counts);
        } else {
            xtalk_tile_kernel<6, false><<<blocks, XT_TILE, 0, st>>>(img, mask, W, ysize_chan, xsize_chan, k, src_bad,
                                                                   (uint32_t)bits->edge, nullptr);
        }
        BBX_CHECK_LAUNCH("xtalk_tile_kernel");
        return 0;
    }
    if (out_counts && bbx_mask_counts(mask, (size_t)H * W, out_counts, stream)) return -2;
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
