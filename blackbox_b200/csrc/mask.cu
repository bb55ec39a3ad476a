// mask.cu -- mask_init morphology (reference: blackbox.py:4473-4566, fill_sat_holes 4584-4596),
// connected-component counts (ndimage.label, blackbox.py:4354, 4544) and mask_header counts.
//
// All integer / byte work, HBM-bound.  "Computed saturation" (mask_sat of the reference, i.e.
// data >= satlevel, as opposed to a 'saturated' bit that a bad-pixel mask might carry) travels
// in the spare bit BBX_TMP_SAT (0x80) of the mask; bbx_fill_sat_holes clears it at the end.
#include <cooperative_groups.h>
#include "bbx_common.cuh"
#include "bg_track.cuh"

namespace cg = cooperative_groups;


// --------------------------------------------------------------------------------------------
// crosstalk-victim bit: a pixel is flagged if the pixel at the same tile position (y mirrored
// between the CCD halves) is saturated in any OTHER channel.  One thread owns the 16 mirrored
// words (4 pixels each) of one tile position, so every byte is read and written exactly once.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
xtalk_victim_kernel(uint8_t *__restrict__ mask, int W, int ysc, int xsc, uint32_t bit_xtalk)
{
    const int words = xsc / 4;
    const long long total = (long long)ysc * words;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ly = (int)(t / words), lx = (int)(t - (long long)ly * words) * 4;
        uint32_t w[16], tot = 0;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
            const size_t off = (size_t)row * W + (size_t)(c & 7) * xsc + lx;
            w[c] = *reinterpret_cast<const uint32_t *>(mask + off);
            tot += (w[c] >> 7) & 0x01010101u;           // per-byte count of saturated sources
        }
        if (tot == 0) continue;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t others = tot - ((w[c] >> 7) & 0x01010101u);
            uint32_t nz = others | (others >> 1) | (others >> 2) | (others >> 3) | (others >> 4);
            nz &= 0x01010101u;
            if (nz) {
                const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
                const size_t off = (size_t)row * W + (size_t)(c & 7) * xsc + lx;
                *reinterpret_cast<uint32_t *>(mask + off) = w[c] | (nz * bit_xtalk);
            }
        }
    }
}

// generic (any xsize_chan): one thread per tile position and pixel
__global__ void __launch_bounds__(256)
xtalk_victim_scalar_kernel(uint8_t *__restrict__ mask, int W, int ysc, int xsc, uint32_t bit_xtalk)
{
    const long long total = (long long)ysc * xsc;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ly = (int)(t / xsc), lx = (int)(t - (long long)ly * xsc);
        uint8_t w[16];
        int tot = 0;
        for (int c = 0; c < 16; c++) {
            const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
            w[c] = mask[(size_t)row * W + (size_t)(c & 7) * xsc + lx];
            tot += (w[c] >> 7) & 1;
        }
        if (tot == 0) continue;
        for (int c = 0; c < 16; c++)
            if (tot - ((w[c] >> 7) & 1) > 0) {
                const int row = (c < 8) ? ly : (ysc + (ysc - 1 - ly));
                mask[(size_t)row * W + (size_t)(c & 7) * xsc + lx] = w[c] | (uint8_t)bit_xtalk;
            }
    }
}

// saturated-connected: 8-neighbour of a saturated pixel, not itself saturated (zero border)
__global__ void __launch_bounds__(256)
satcon_kernel(uint8_t *__restrict__ mask, int H, int W, uint32_t bit_satcon)
{
    const long long total = (long long)H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (long long)y * W);
        const uint8_t m = mask[p];
        if (m & BBX_TMP_SAT) continue;
        bool near = false;
        for (int dy = -1; dy <= 1; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -1; dx <= 1; dx++) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                near |= (mask[(size_t)yy * W + xx] & BBX_TMP_SAT) != 0;
            }
        }
        if (near) mask[p] = m | (uint8_t)bit_satcon;
    }
}

extern "C" int bbx_mask_sat_neighbours(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                       const bbx_maskbits *bits, void *stream)
{
    BBX_REQUIRE(mask && bits, "bbx_mask_sat_neighbours: null argument");
    BBX_REQUIRE(H == 2 * ysize_chan && W == 8 * xsize_chan, "bbx_mask_sat_neighbours: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    BBX_REQUIRE(((bits->bad | bits->cosmic | bits->saturated | bits->satcon | bits->sattrail | bits->edge | bits->crosstalk) & 0x80) == 0,
                "bbx_mask_sat_neighbours: mask value 128 is reserved for internal use");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = BBX_SM_COUNT * 8;
    if (xsize_chan % 4 == 0 && ((uintptr_t)mask % 4) == 0)
        xtalk_victim_kernel<<<blocks, 256, 0, s>>>(mask, W, ysize_chan, xsize_chan, (uint32_t)bits->crosstalk);
    else
        xtalk_victim_scalar_kernel<<<blocks, 256, 0, s>>>(mask, W, ysize_chan, xsize_chan, (uint32_t)bits->crosstalk);
    BBX_CHECK_LAUNCH("xtalk_victim_kernel");
    satcon_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, H, W, (uint32_t)bits->satcon);
    BBX_CHECK_LAUNCH("satcon_kernel");
    return 0;
}

// --------------------------------------------------------------------------------------------
// fill_sat_holes
//   m      = saturated | saturated-connected bits
//   closed = erode3(dilate3(m)), outside the image counts as 0 in both steps
//   filled = closed plus every background region not 8-connected to the image border
//   mask[filled & mask == 0] = saturated-connected
// State image S (work): 0 = closed foreground, 1 = background, 2 = background known to reach
// the border.  Background pixels with a free straight line to the border in +-x or +-y are
// resolved implicitly from the per-row / per-column foreground extents; only the remaining
// "enclosed in all four directions" pixels are propagated (tile-local fixed point inside a
// persistent cooperative kernel, grid-wide rounds until nothing changes).
// --------------------------------------------------------------------------------------------
struct HoleWork {
    uint8_t *S;          // [H*W]
    int *rowmin, *rowmax, *colmin, *colmax;
    int *tiles;          // active tile list
    int *tile_flag;      // per tile: already on the list
    int *counters;       // [0] number of active tiles, [1] changed flag, [2] any foreground
    int max_tiles;
};
#define HOLE_TILE 64

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static HoleWork carve_hole_work(void *work, int H, int W)
{
    HoleWork hw;
    uint8_t *p = (uint8_t *)work;
    hw.S = p; p += align_up((size_t)H * W, 256);
    hw.rowmin = (int *)p; p += align_up(sizeof(int) * H, 256);
    hw.rowmax = (int *)p; p += align_up(sizeof(int) * H, 256);
    hw.colmin = (int *)p; p += align_up(sizeof(int) * W, 256);
    hw.colmax = (int *)p; p += align_up(sizeof(int) * W, 256);
    hw.max_tiles = ceil_div(H, HOLE_TILE) * ceil_div(W, HOLE_TILE);
    hw.tiles = (int *)p; p += align_up(sizeof(int) * hw.max_tiles, 256);
    hw.tile_flag = (int *)p; p += align_up(sizeof(int) * hw.max_tiles, 256);
    hw.counters = (int *)p;
    return hw;
}

extern "C" size_t bbx_fill_holes_work_bytes(int H, int W)
{
    const size_t tiles = (size_t)ceil_div(H, HOLE_TILE) * ceil_div(W, HOLE_TILE);
    return align_up((size_t)H * W, 256) + 2 * align_up(sizeof(int) * H, 256) + 2 * align_up(sizeof(int) * W, 256) +
           2 * align_up(sizeof(int) * tiles, 256) + 256;
}

__global__ void hole_init_kernel(HoleWork hw, int H, int W)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) { hw.rowmin[i] = W; hw.rowmax[i] = -1; }
    if (i < W) { hw.colmin[i] = H; hw.colmax[i] = -1; }
    if (i < 4) hw.counters[i] = 0;
}

// closing + state image + foreground extents.  32x32 outputs per block, halo 2.
__global__ void __launch_bounds__(1024)
hole_close_kernel(const uint8_t *__restrict__ mask, HoleWork hw, int H, int W, uint32_t mbits)
{
    __shared__ uint8_t m[36][36];
    __shared__ uint8_t d[34][34];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int i = ty * 32 + tx; i < 36 * 36; i += 1024) {
        const int yy = i / 36, xx = i - yy * 36;
        const int gy = y0 + yy - 2, gx = x0 + xx - 2;
        uint8_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = (mask[(size_t)gy * W + gx] & mbits) != 0;
        m[yy][xx] = v;
    }
    __syncthreads();
    for (int i = ty * 32 + tx; i < 34 * 34; i += 1024) {
        const int yy = i / 34, xx = i - yy * 34;            // dilated value at (y0+yy-1, x0+xx-1)
        const int gy = y0 + yy - 1, gx = x0 + xx - 1;
        uint8_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) v |= m[yy + a][xx + b];
        }
        d[yy][xx] = v;      // positions outside the image stay 0: erosion fails next to the border
    }
    __syncthreads();
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy < H && gx < W) {
        uint8_t e = 1;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) e &= d[ty + a][tx + b];
        hw.S[(size_t)gy * W + gx] = e ? 0 : 1;
        if (e) {
            atomicMin(&hw.rowmin[gy], gx); atomicMax(&hw.rowmax[gy], gx);
            atomicMin(&hw.colmin[gx], gy); atomicMax(&hw.colmax[gx], gy);
            hw.counters[2] = 1;
        }
    }
}

__device__ __forceinline__ bool hole_free_line(const HoleWork &hw, int y, int x)
{
    return x < hw.rowmin[y] || x > hw.rowmax[y] || y < hw.colmin[x] || y > hw.colmax[x];
}

// one block per image row: scan only the span between the first and last foreground pixel and
// register tiles that hold unresolved background
__global__ void __launch_bounds__(256)
hole_candidates_kernel(HoleWork hw, int H, int W, int *__restrict__ tile_flag)
{
    const int y = blockIdx.x;
    const int a = hw.rowmin[y], b = hw.rowmax[y];
    if (a > b) return;
    const int tiles_x = (W + HOLE_TILE - 1) / HOLE_TILE;
    for (int x = a + threadIdx.x; x <= b; x += blockDim.x) {
        if (hw.S[(size_t)y * W + x] == 1 && !hole_free_line(hw, y, x)) {
            const int t = (y / HOLE_TILE) * tiles_x + x / HOLE_TILE;
            if (atomicExch(&tile_flag[t], 1) == 0) {
                const int slot = atomicAdd(&hw.counters[0], 1);
                hw.tiles[slot] = t;
            }
        }
    }
}

// persistent cooperative kernel: propagate "reaches the border" through unresolved background
__global__ void __launch_bounds__(256)
hole_propagate_kernel(HoleWork hw, int H, int W, int max_rounds, int32_t *unconverged)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ uint8_t t[HOLE_TILE + 2][HOLE_TILE + 2];
    __shared__ int s_changed, s_any;
    const int ntiles = hw.counters[0];
    const int tiles_x = (W + HOLE_TILE - 1) / HOLE_TILE;
    int round = 0;
    if (ntiles == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) *unconverged = 0; return; }
    for (;;) {
        if (blockIdx.x == 0 && threadIdx.x == 0) hw.counters[1] = 0;
        grid.sync();
        for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
            const int tile = hw.tiles[ti];
            const int y0 = (tile / tiles_x) * HOLE_TILE, x0 = (tile % tiles_x) * HOLE_TILE;
            // load tile + 1-pixel halo: 0 fg, 1 unknown, 2 outside (explicit or implicit)
            for (int i = threadIdx.x; i < (HOLE_TILE + 2) * (HOLE_TILE + 2); i += blockDim.x) {
                const int yy = i / (HOLE_TILE + 2), xx = i - yy * (HOLE_TILE + 2);
                const int gy = y0 + yy - 1, gx = x0 + xx - 1;
                uint8_t v = 2;                                  // outside the image = border value 1
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    v = hw.S[(size_t)gy * W + gx];
                    if (v == 1 && hole_free_line(hw, gy, gx)) v = 2;
                }
                t[yy][xx] = v;
            }
            if (threadIdx.x == 0) s_any = 0;
            __syncthreads();
            for (;;) {
                if (threadIdx.x == 0) s_changed = 0;
                __syncthreads();
                for (int i = threadIdx.x; i < HOLE_TILE * HOLE_TILE; i += blockDim.x) {
                    const int yy = i / HOLE_TILE + 1, xx = i % HOLE_TILE + 1;
                    if (t[yy][xx] != 1) continue;
                    const bool reach = t[yy - 1][xx - 1] == 2 || t[yy - 1][xx] == 2 || t[yy - 1][xx + 1] == 2 ||
                                       t[yy][xx - 1] == 2 || t[yy][xx + 1] == 2 ||
                                       t[yy + 1][xx - 1] == 2 || t[yy + 1][xx] == 2 || t[yy + 1][xx + 1] == 2;
                    if (reach) { t[yy][xx] = 2; s_changed = 1; }
                }
                __syncthreads();
                const int ch = s_changed;
                __syncthreads();
                if (!ch) break;
                if (threadIdx.x == 0) s_any = 1;
            }
            __syncthreads();
            if (s_any) {
                for (int i = threadIdx.x; i < HOLE_TILE * HOLE_TILE; i += blockDim.x) {
                    const int yy = i / HOLE_TILE, xx = i % HOLE_TILE;
                    const int gy = y0 + yy, gx = x0 + xx;
                    if (gy < H && gx < W && t[yy + 1][xx + 1] == 2) hw.S[(size_t)gy * W + gx] = 2;
                }
                if (threadIdx.x == 0) hw.counters[1] = 1;
            }
            __syncthreads();
        }
        grid.sync();
        round++;
        const int changed = hw.counters[1];
        if (!changed || round >= max_rounds) {
            if (blockIdx.x == 0 && threadIdx.x == 0) *unconverged = changed ? 1 : 0;
            break;
        }
        grid.sync();            // everybody has read the flag before it is reset
    }
}

__global__ void __launch_bounds__(256)
hole_commit_kernel(uint8_t *__restrict__ mask, HoleWork hw, int H, int W, uint32_t bit_satcon,
                   const int32_t *__restrict__ unconverged)
{
    if (*unconverged) return;         // bbx_fill_holes_more has to finish the propagation first
    const long long total = (long long)H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        uint8_t m = mask[p];
        const uint8_t m0 = m;
        m &= (uint8_t)~BBX_TMP_SAT;
        if (m == 0) {
            const uint8_t s = hw.S[p];
            bool filled = (s == 0);
            if (s == 1) {
                const int y = (int)(p / W), x = (int)(p - (long long)y * W);
                filled = !hole_free_line(hw, y, x);
            }
            if (filled) m = (uint8_t)bit_satcon;
        }
        if (m != m0) mask[p] = m;
    }
}

static int launch_propagate(HoleWork hw, int H, int W, int rounds, int32_t *unconverged, cudaStream_t s)
{
    int per_sm = 0;
    BBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hole_propagate_kernel, 256, 0));
    BBX_REQUIRE(per_sm > 0, "hole_propagate_kernel cannot be made resident");
    int dev = 0, sms = 0;
    BBX_CUDA(cudaGetDevice(&dev));
    BBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * (per_sm > 2 ? 2 : per_sm);
    void *args[] = {&hw, &H, &W, &rounds, &unconverged};
    BBX_CUDA(cudaLaunchCooperativeKernel((const void *)hole_propagate_kernel, dim3(blocks), dim3(256), args, 0, s));
    return 0;
}

extern "C" int bbx_fill_sat_holes(uint8_t *mask, int H, int W, const bbx_maskbits *bits, void *work,
                                  int rounds, int32_t *unconverged, void *stream)
{
    BBX_REQUIRE(mask && bits && work && unconverged, "bbx_fill_sat_holes: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    HoleWork hw = carve_hole_work(work, H, W);
    const int n = H > W ? H : W;
    hole_init_kernel<<<ceil_div(n, 256), 256, 0, s>>>(hw, H, W);
    BBX_CHECK_LAUNCH("hole_init_kernel");
    const uint32_t mbits = (uint32_t)(bits->saturated | bits->satcon);
    hole_close_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 32)), dim3(32, 32), 0, s>>>(mask, hw, H, W, mbits);
    BBX_CHECK_LAUNCH("hole_close_kernel");
    BBX_CUDA(cudaMemsetAsync(hw.tile_flag, 0, sizeof(int) * hw.max_tiles, s));
    hole_candidates_kernel<<<H, 256, 0, s>>>(hw, H, W, hw.tile_flag);
    BBX_CHECK_LAUNCH("hole_candidates_kernel");
    if (launch_propagate(hw, H, W, rounds, unconverged, s)) return -2;
    hole_commit_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, hw, H, W, (uint32_t)bits->satcon, unconverged);
    BBX_CHECK_LAUNCH("hole_commit_kernel");
    return 0;
}

extern "C" int bbx_fill_holes_more(uint8_t *mask, int H, int W, const bbx_maskbits *bits, void *work,
                                   int rounds, int32_t *unconverged, void *stream)
{
    BBX_REQUIRE(mask && bits && work && unconverged, "bbx_fill_holes_more: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    HoleWork hw = carve_hole_work(work, H, W);
    if (launch_propagate(hw, H, W, rounds, unconverged, s)) return -2;
    hole_commit_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, hw, H, W, (uint32_t)bits->satcon, unconverged);
    BBX_CHECK_LAUNCH("hole_commit_kernel");
    return 0;
}

// --------------------------------------------------------------------------------------------
// 8-connected component count (union-find with atomicMin, label = linear pixel index)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int i)
{
    int p = L[i];
    while (p != i) { i = p; p = L[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int *L, int a, int b)
{
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t *__restrict__ mask, uint32_t bit, long long total, int *__restrict__ L)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x)
        if (mask[p] & bit) L[p] = (int)p;
}

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t *__restrict__ mask, uint32_t bit, int H, int W, int *__restrict__ L)
{
    const long long total = (long long)H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        if (!(mask[p] & bit)) continue;
        const int y = (int)(p / W), x = (int)(p - (long long)y * W);
        if (x > 0 && (mask[p - 1] & bit)) uf_union(L, (int)p, (int)p - 1);
        if (y > 0) {
            const long long q = p - W;
            if (mask[q] & bit) uf_union(L, (int)p, (int)q);
            if (x > 0 && (mask[q - 1] & bit)) uf_union(L, (int)p, (int)q - 1);
            if (x + 1 < W && (mask[q + 1] & bit)) uf_union(L, (int)p, (int)q + 1);
        }
    }
}

__global__ void __launch_bounds__(256)
ccl_count_kernel(const uint8_t *__restrict__ mask, uint32_t bit, long long total, const int *__restrict__ L,
                 int32_t *__restrict__ out)
{
    int c = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x)
        if ((mask[p] & bit) && L[p] == (int)p) c++;
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

extern "C" int bbx_count_objects(const uint8_t *mask, int bit, int H, int W, int32_t *labels,
                                 int32_t *out_count, void *stream)
{
    BBX_REQUIRE(mask && labels && out_count, "bbx_count_objects: null argument");
    BBX_REQUIRE((long long)H * W < 2147483647LL, "bbx_count_objects: image too large for 32-bit labels");
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)H * W;
    BBX_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
    ccl_init_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, (uint32_t)bit, total, labels);
    BBX_CHECK_LAUNCH("ccl_init_kernel");
    ccl_merge_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, (uint32_t)bit, H, W, labels);
    BBX_CHECK_LAUNCH("ccl_merge_kernel");
    ccl_count_kernel<<<BBX_SM_COUNT * 16, 256, 0, s>>>(mask, (uint32_t)bit, total, labels, out_count);
    BBX_CHECK_LAUNCH("ccl_count_kernel");
    return 0;
}

// --------------------------------------------------------------------------------------------
// per-bit pixel counts
// --------------------------------------------------------------------------------------------
// 16 mask bytes per load; bit plane k of a 32-bit word is (w >> k) & 0x01010101, its population
// count the number of pixels carrying bit k.  Counts stay in registers (int: < 2^31 bytes per
// thread), one warp reduction and one atomic per warp and bit at the end.
__global__ void __launch_bounds__(256)
mask_counts_kernel(const uint8_t *__restrict__ mask, size_t n, unsigned long long *__restrict__ out)
{
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)mask & 15)) & 15));     // bytes before 16-byte alignment
    const size_t nvec = (n - head) / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(mask + head);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < nvec; i += nth) {
        const uint4 q = ld_stream_u4(v + i);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t pl = 0x01010101u << k;
            c[k] += __popc(q.x & pl) + __popc(q.y & pl) + __popc(q.z & pl) + __popc(q.w & pl);
        }
    }
    // the unaligned head and the tail, byte by byte
    for (size_t p = tid; p < head; p += nth) {
        const uint32_t m = mask[p];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] += (m >> k) & 1;
    }
    for (size_t p = head + nvec * 16 + tid; p < n; p += nth) {
        const uint32_t m = mask[p];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] += (m >> k) & 1;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int t = warp_sum(c[k]);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&out[k], (unsigned long long)t);
    }
}

extern "C" int bbx_mask_counts(const uint8_t *mask, size_t n, unsigned long long *out_counts, void *stream)
{
    BBX_REQUIRE(mask && out_counts, "bbx_mask_counts: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    BBX_CUDA(cudaMemsetAsync(out_counts, 0, 8 * sizeof(unsigned long long), s));
    if (n == 0) return 0;
    mask_counts_kernel<<<BBX_SM_COUNT * 4, 256, 0, s>>>(mask, n, out_counts);
    BBX_CHECK_LAUNCH("mask_counts_kernel");
    return 0;
}

// ============================================================================================
// Sparse mask morphology: the same results as the dense kernels above, driven by the list of
// saturated pixels ("seeds") that bbx_reduce_apply appends while it writes the mask.  Saturated
// pixels are a few 1e-5 of a frame, so apart from one memset of the state image nothing here
// touches more than small neighbourhoods of the seeds.
// ============================================================================================
#define SEED_BPM 0x80000000u

__device__ __forceinline__ void byte_or(uint8_t *base, size_t p, unsigned int bits)
{
    unsigned int *word = reinterpret_cast<unsigned int *>(base + (p & ~(size_t)3));
    atomicOr(word, bits << ((unsigned int)(p & 3) * 8));
}
// the same on the frame mask, telling LACosmic's background statistics when this was the first
// bit of the byte (bg_track.cuh: the pixel was counted as unmasked by the fused pass)
__device__ __forceinline__ void mask_or(uint8_t *mask, size_t p, unsigned int bits, const BgTrack &trk)
{
    unsigned int *word = reinterpret_cast<unsigned int *>(mask + (p & ~(size_t)3));
    const unsigned int sh = (unsigned int)(p & 3) * 8;
    const unsigned int old = atomicOr(word, bits << sh);
    if (((old >> sh) & 0xffu) == 0 && bits != 0) bg_untrack(trk, p);
}
__device__ __forceinline__ void byte_and(uint8_t *base, size_t p, unsigned int bits)
{
    unsigned int *word = reinterpret_cast<unsigned int *>(base + (p & ~(size_t)3));
    const unsigned int sh = (unsigned int)(p & 3) * 8;
    atomicAnd(word, (bits << sh) | ~(0xffu << sh));
}

__device__ __forceinline__ unsigned int seed_len(const unsigned int *count, unsigned int cap, int32_t *status)
{
    const unsigned int n = *count;
    if (n > cap) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, 1); return cap; }
    return n;
}

// crosstalk victims (15 mirrored positions) and saturated-connected neighbours of every
// saturated pixel; also resets the object counter and the status word
__global__ void __launch_bounds__(128)
ms_neigh_kernel(uint8_t *mask, int H, int W, int ysc, int xsc, unsigned int bit_xtalk, unsigned int bit_satcon,
                const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count, unsigned int cap,
                int32_t *status, BgTrack trk)
{
    const unsigned int n = seed_len(count, cap, status);
    // 24 work items per seed: 15 victims + 8 neighbours (+1 idle)
    const unsigned long long total = (unsigned long long)n * 24ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[t / 24];
        if (sd & SEED_BPM) continue;
        const int k = (int)(t % 24);
        const int y = (int)(sd / (unsigned int)W), x = (int)(sd - (unsigned int)y * (unsigned int)W);
        if (k < 16) {
            // victim channel k (skip the source channel itself)
            const int r = y / ysc, c = x / xsc, ly = y - r * ysc, lx = x - c * xsc;
            const int src = r * 8 + c;
            if (k == src) continue;
            const int vr = k / 8, vc = k % 8;
            const int vy = (vr == r) ? ly : (ysc - 1 - ly);
            mask_or(mask, (size_t)(vr * ysc + vy) * W + (size_t)vc * xsc + lx, bit_xtalk, trk);
        } else {
            const int j = k - 16;                     // 8 neighbours, skipping the centre
            const int o = j < 4 ? j : j + 1;
            const int qy = y + o / 3 - 1, qx = x + o % 3 - 1;
            if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
            const size_t q = (size_t)qy * W + qx;
            if (!(mask[q] & BBX_TMP_SAT)) mask_or(mask, q, bit_satcon, trk);
        }
    }
}

__global__ void __launch_bounds__(128)
ms_ccl_init_kernel(const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count, unsigned int cap,
                   int *__restrict__ L, int32_t *out_nobj)
{
    const unsigned int n = min(*count, cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_nobj = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[k];
        if (!(sd & SEED_BPM)) L[sd] = (int)sd;
    }
}

__global__ void __launch_bounds__(128)
ms_ccl_merge_kernel(const uint8_t *__restrict__ mask, int H, int W, const unsigned int *__restrict__ seeds,
                    const unsigned int *__restrict__ count, unsigned int cap, int *__restrict__ L)
{
    if (*count > cap) return;                   // incomplete list: labels of neighbours may be missing
    const unsigned int n = *count;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[k];
        if (sd & SEED_BPM) continue;
        const int p = (int)sd;
        const int y = p / W, x = p - y * W;
        if (x > 0 && (mask[p - 1] & BBX_TMP_SAT)) uf_union(L, p, p - 1);
        if (y > 0) {
            const int q = p - W;
            if (mask[q] & BBX_TMP_SAT) uf_union(L, p, q);
            if (x > 0 && (mask[q - 1] & BBX_TMP_SAT)) uf_union(L, p, q - 1);
            if (x + 1 < W && (mask[q + 1] & BBX_TMP_SAT)) uf_union(L, p, q + 1);
        }
    }
}

__global__ void __launch_bounds__(128)
ms_ccl_count_kernel(const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count, unsigned int cap,
                    const int *__restrict__ L, int32_t *out_nobj)
{
    const unsigned int n = min(*count, cap);
    int c = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[k];
        if (!(sd & SEED_BPM) && L[sd] == (int)sd) c++;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out_nobj, c);
}

// closing evaluated only within 2 pixels of a seed: closed(q) = AND over N3(q) of
// [OR over N3(r) of m], everything outside the image = 0
__device__ __forceinline__ bool closed_at(const uint8_t *__restrict__ mask, int H, int W, int qy, int qx, unsigned int mbits)
{
    if (qy < 1 || qy >= H - 1 || qx < 1 || qx >= W - 1) return false;     // erosion fails at the frame
    // m on the 5x5 around q, as 5 row bit-fields
    unsigned int rows[5];
#pragma unroll
    for (int a = 0; a < 5; a++) {
        const int yy = qy + a - 2;
        unsigned int bitsr = 0;
        if (yy >= 0 && yy < H) {
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const int xx = qx + b - 2;
                if (xx >= 0 && xx < W && (mask[(size_t)yy * W + xx] & mbits)) bitsr |= 1u << b;
            }
        }
        rows[a] = bitsr;
    }
    // dilated value at (qy + a - 1, qx + b - 1), a,b in 0..2: any m in rows a..a+2, cols b..b+2
    bool all = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const unsigned int v = rows[a] | rows[a + 1] | rows[a + 2];
#pragma unroll
        for (int b = 0; b < 3; b++) all = all && ((v >> b) & 7u) != 0;
    }
    return all;
}

__global__ void __launch_bounds__(128)
ms_close_kernel(const uint8_t *__restrict__ mask, HoleWork hw, int H, int W, unsigned int mbits,
                const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count, unsigned int cap)
{
    const unsigned int n = min(*count, cap);
    const unsigned long long total = (unsigned long long)n * 25ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[t / 25] & ~SEED_BPM;
        const int k = (int)(t % 25);
        const int y = (int)(sd / (unsigned int)W), x = (int)(sd - (unsigned int)y * (unsigned int)W);
        const int qy = y + k / 5 - 2, qx = x + k % 5 - 2;
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
        if (!closed_at(mask, H, W, qy, qx, mbits)) continue;
        hw.S[(size_t)qy * W + qx] = 0;
        atomicMin(&hw.rowmin[qy], qx); atomicMax(&hw.rowmax[qy], qx);
        atomicMin(&hw.colmin[qx], qy); atomicMax(&hw.colmax[qx], qy);
    }
}

// commit: closed pixels around the seeds, enclosed background in the candidate tiles, and the
// removal of the internal saturation marker
__global__ void __launch_bounds__(128)
ms_commit_seeds_kernel(uint8_t *mask, HoleWork hw, int H, int W, unsigned int bit_satcon,
                       const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count, unsigned int cap,
                       const int32_t *__restrict__ unconverged, int32_t *status, BgTrack trk)
{
    if (*unconverged) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, 2); return; }
    const unsigned int n = min(*count, cap);
    const unsigned long long total = (unsigned long long)n * 25ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[t / 25] & ~SEED_BPM;
        const int k = (int)(t % 25);
        const int y = (int)(sd / (unsigned int)W), x = (int)(sd - (unsigned int)y * (unsigned int)W);
        const int qy = y + k / 5 - 2, qx = x + k % 5 - 2;
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
        const size_t q = (size_t)qy * W + qx;
        if (hw.S[q] == 0 && (mask[q] & 0x7fu) == 0) mask_or(mask, q, bit_satcon, trk);
    }
}

__global__ void __launch_bounds__(256)
ms_commit_tiles_kernel(uint8_t *mask, HoleWork hw, int H, int W, unsigned int bit_satcon,
                       const int32_t *__restrict__ unconverged, BgTrack trk)
{
    if (*unconverged) return;
    const int ntiles = hw.counters[0];
    const int tiles_x = (W + HOLE_TILE - 1) / HOLE_TILE;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const int tile = hw.tiles[ti];
        const int y0 = (tile / tiles_x) * HOLE_TILE, x0 = (tile % tiles_x) * HOLE_TILE;
        for (int i = threadIdx.x; i < HOLE_TILE * HOLE_TILE; i += blockDim.x) {
            const int gy = y0 + i / HOLE_TILE, gx = x0 + i % HOLE_TILE;
            if (gy >= H || gx >= W) continue;
            const size_t q = (size_t)gy * W + gx;
            if (hw.S[q] == 1 && !hole_free_line(hw, gy, gx) && (mask[q] & 0x7fu) == 0) mask_or(mask, q, bit_satcon, trk);
        }
    }
}

// The state image goes back to "background everywhere" (1) where this frame touched it: the 5x5
// boxes around the seeds (ms_close_kernel) and the candidate tiles (hole_propagate_kernel), so the
// next frame needs no 111 MB memset.
__global__ void __launch_bounds__(256)
ms_restore_state_kernel(HoleWork hw, int H, int W, const unsigned int *__restrict__ seeds,
                        const unsigned int *__restrict__ count, unsigned int cap)
{
    const unsigned int n = min(*count, cap);
    const unsigned long long total = (unsigned long long)n * 25ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[t / 25] & ~SEED_BPM;
        const int k = (int)(t % 25);
        const int y = (int)(sd / (unsigned int)W), x = (int)(sd - (unsigned int)y * (unsigned int)W);
        const int qy = y + k / 5 - 2, qx = x + k % 5 - 2;
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
        hw.S[(size_t)qy * W + qx] = 1;
    }
    const int ntiles = hw.counters[0];
    const int tiles_x = (W + HOLE_TILE - 1) / HOLE_TILE;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const int tile = hw.tiles[ti];
        const int y0 = (tile / tiles_x) * HOLE_TILE, x0 = (tile % tiles_x) * HOLE_TILE;
        for (int i = threadIdx.x; i < HOLE_TILE * HOLE_TILE; i += blockDim.x) {
            const int gy = y0 + i / HOLE_TILE, gx = x0 + i % HOLE_TILE;
            if (gy < H && gx < W) hw.S[(size_t)gy * W + gx] = 1;
        }
    }
}

__global__ void __launch_bounds__(128)
ms_clear_marker_kernel(uint8_t *mask, const unsigned int *__restrict__ seeds, const unsigned int *__restrict__ count,
                       unsigned int cap, const int32_t *__restrict__ unconverged)
{
    if (*unconverged) return;
    const unsigned int n = min(*count, cap);
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int sd = seeds[k];
        if (!(sd & SEED_BPM)) byte_and(mask, sd, 0x7fu);
    }
}

extern "C" int bbx_mask_morph_sparse(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                     const bbx_maskbits *bits, const unsigned int *seeds, const unsigned int *seed_count,
                                     unsigned int seed_cap, void *work, int32_t *labels, int32_t *out_nobj, int rounds,
                                     int32_t *status, void *stream)
{
    return bbx_mask_morph_sparse_track(mask, H, W, ysize_chan, xsize_chan, bits, seeds, seed_count, seed_cap, work,
                                       labels, out_nobj, rounds, status, nullptr, nullptr, 0, stream);
}

// The same with two optional extras.  img / lac_work (after bbx_reduce_apply_scan: the reduced image
// and LACosmic's work buffer of that call) let the morphology take every pixel it masks for the first
// time out of the background statistics the fused pass has collected against the seed mask.
// state_clean != 0: the first H * W bytes of `work` are all ones -- as a fresh caller sets them once
// and as every call of this function leaves them -- so the 111 MB memset per frame is skipped
// (the dense entry points bbx_fill_sat_holes / _more do not keep that promise).
extern "C" int bbx_mask_morph_sparse_track(uint8_t *mask, int H, int W, int ysize_chan, int xsize_chan,
                                           const bbx_maskbits *bits, const unsigned int *seeds,
                                           const unsigned int *seed_count, unsigned int seed_cap, void *work,
                                           int32_t *labels, int32_t *out_nobj, int rounds, int32_t *status,
                                           const float *img, void *lac_work, int state_clean, void *stream)
{
    const BgTrack trk = lac_sparse_bg_track(img, lac_work, H, W);
    BBX_REQUIRE(mask && bits && seeds && seed_count && work && labels && out_nobj && status, "bbx_mask_morph_sparse: null argument");
    BBX_REQUIRE(H == 2 * ysize_chan && W == 8 * xsize_chan, "bbx_mask_morph_sparse: %d x %d is not 2 x 8 channels of %d x %d", H, W, ysize_chan, xsize_chan);
    BBX_REQUIRE(((uintptr_t)mask & 3) == 0, "bbx_mask_morph_sparse: mask must be 4-byte aligned");
    BBX_REQUIRE(((bits->bad | bits->cosmic | bits->saturated | bits->satcon | bits->sattrail | bits->edge | bits->crosstalk) & 0x80) == 0,
                "bbx_mask_morph_sparse: mask value 128 is reserved for internal use");
    cudaStream_t s = (cudaStream_t)stream;
    HoleWork hw = carve_hole_work(work, H, W);
    const int lb = BBX_SM_COUNT * 4;
    const unsigned int mbits = (unsigned int)(bits->saturated | bits->satcon);
    BBX_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    ms_neigh_kernel<<<lb, 128, 0, s>>>(mask, H, W, ysize_chan, xsize_chan, (unsigned int)bits->crosstalk,
                                       (unsigned int)bits->satcon, seeds, seed_count, seed_cap, status, trk);
    ms_ccl_init_kernel<<<lb, 128, 0, s>>>(seeds, seed_count, seed_cap, labels, out_nobj);
    ms_ccl_merge_kernel<<<lb, 128, 0, s>>>(mask, H, W, seeds, seed_count, seed_cap, labels);
    ms_ccl_count_kernel<<<lb, 128, 0, s>>>(seeds, seed_count, seed_cap, labels, out_nobj);
    // hole filling: state image = background everywhere, closed foreground near the seeds.  A
    // caller that keeps the work buffer between frames says so (state_clean): this function leaves the
    // state image as it wants to find it, all ones
    if (!state_clean) BBX_CUDA(cudaMemsetAsync(hw.S, 1, (size_t)H * W, s));
    const int n = H > W ? H : W;
    hole_init_kernel<<<ceil_div(n, 256), 256, 0, s>>>(hw, H, W);
    ms_close_kernel<<<lb, 128, 0, s>>>(mask, hw, H, W, mbits, seeds, seed_count, seed_cap);
    BBX_CUDA(cudaMemsetAsync(hw.tile_flag, 0, sizeof(int) * hw.max_tiles, s));
    hole_candidates_kernel<<<H, 256, 0, s>>>(hw, H, W, hw.tile_flag);
    BBX_CHECK_LAUNCH("bbx_mask_morph_sparse");
    int32_t *unconverged = status + 1;           // status[1]: scratch word of the propagation
    if (launch_propagate(hw, H, W, rounds, unconverged, s)) return -2;
    ms_commit_seeds_kernel<<<lb, 128, 0, s>>>(mask, hw, H, W, (unsigned int)bits->satcon, seeds, seed_count, seed_cap,
                                              unconverged, status, trk);
    ms_commit_tiles_kernel<<<BBX_SM_COUNT * 2, 256, 0, s>>>(mask, hw, H, W, (unsigned int)bits->satcon, unconverged, trk);
    ms_clear_marker_kernel<<<lb, 128, 0, s>>>(mask, seeds, seed_count, seed_cap, unconverged);
    ms_restore_state_kernel<<<lb, 256, 0, s>>>(hw, H, W, seeds, seed_count, seed_cap);
    BBX_CHECK_LAUNCH("bbx_mask_morph_sparse");
    return 0;
}
