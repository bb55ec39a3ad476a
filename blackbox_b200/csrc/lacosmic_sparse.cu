// lacosmic_sparse.cu -- LACosmic with lazy evaluation: bit-identical results to the dense
// implementation (lacosmic.cu) at a tiny fraction of the arithmetic.
//
// Observation (all inequalities hold exactly in float32 because +, /, sqrt and the median are
// monotone under round-to-nearest):
//   * s  = L+ / (2 noise) >= 0 everywhere, hence med5(s) >= 0 and  s' = s - med5(s) <= s;
//   * noise = sqrt(max(med5(img), 1e-5) + rn^2) >= noise_min = sqrt(1e-5 + rn^2), hence
//     s <= s_ub = L+ / (2 noise_min);
//   * a pixel can only become a cosmic-ray pixel if it lies within 2 pixels of a candidate
//     c0 = (s' > sigclip) & good & (s'/f > objlim), and c0 needs s' > sigclip.
// So one dense, HBM-bound pass computes only the 5-point Laplacian L+ and keeps the pixels with
// s_ub > sigclip (a few 1e-4 of a typical frame).  Everything else -- the 5x5 / 3x3 / 7x7
// medians, the fine-structure image, the two growth steps and the cleaning -- is evaluated
// only where its value can influence the result, from short device-side work lists:
//
//   scan   (dense)   L+ -> list A  = { s_ub > sigclip }
//   cand1  (list A)  good, med5(img), s > sigclip            -> list B
//   cand2  (list B)  warp per pixel: s on 5x5, s' > sigclip -> S1 flag; f on 7x7 -> c0 flag, list C0
//   grow1  (C0 x 9)  thread per neighbour q: c1 = S1(q)  (good & s' > sigclip implies q in list B,
//                    and cand2 has tested every pixel of B: no arithmetic left)   -> list C1
//   grow2a (C1 x 9)  thread per neighbour q: members of C1 pass (s' > sigclip >= sigcliplow);
//                    every other unmasked q is queued ONCE (T2 flag)            -> list Q
//   grow2b (list Q)  warp per pixel: s' > sigcliplow                  -> c2: count, crmask, CR list
//   clean  (CR list) lower median of the unflagged 5x5 neighbours
//
// Iterations after the first do not rescan the image: L+ of a pixel changes only if the pixel or
// one of its 4 neighbours was rewritten by the cleaning step, so the next list A is a subset of
// (previous list A) u (cross neighbourhood of the cosmic-ray list) -- re-evaluated sparsely.
//
// The reference's background level (lower median of all unmasked pixels of the INPUT image,
// used for CR pixels without any usable neighbour -- the interior of a fat cosmic-ray blob, which
// sigfrac = 0.01 makes common) costs no extra pass: a strided sample of 32768 pixels brackets the
// median rank (+-5 sigma of the sampling error); the dense scan, which touches every pixel
// anyway, counts the unmasked pixels below the bracket and histograms those inside it with ONE
// BIN PER FLOAT32 KEY (the bracket is an e- or two wide: a few 1e4 distinct keys, a 4 MB table of global
// atomics with next to no contention); the median is the key whose cumulative count reaches the
// wanted rank.  Exact, no list of values, no selection passes.  If the bracket misses (or is
// wider than the table) and the level is then actually needed, LAC_STATUS_NEED_BG is raised and
// the caller repeats the call with mode 2, which runs the radix select of the dense path (three
// passes over image + mask) up front.  If a work list overflows, LAC_STATUS_OVERFLOW is raised
// and the caller repeats the call with the dense implementation.
#include "lacosmic_common.cuh"
#include "apply_common.cuh"

#define FLAG_C0 1u
#define FLAG_C1 2u
#define FLAG_C2 4u
#define FLAG_A 8u
#define FLAG_S1 16u          // good & s' > sigclip (set by cand2 for the pixels of list B that pass)
#define FLAG_T2 32u          // queued for the s' > sigcliplow test of growth step 2 (list Q)
#define FLAG_BITS 6          // flag byte = (iteration stamp << FLAG_BITS) | FLAG_*: stamps 1..3
#define FLAG_MASK 0x3fu
#define STAMP_PERIOD 3

struct SparseCounters {
    unsigned int nA[2];                  // list A of the current / previous iteration (ping-pong)
    unsigned int nB, nC0, nC1, nCR;
    unsigned int bg_valid;               // background level known
    unsigned int clean_lo, clean_hi;     // CR-list range that is new in the current iteration
    unsigned int nQ;                     // list Q (shares the storage of list B, which is spent by then)
};

#define BG_SAMPLES 32768u
#define BG_BINS (1u << 20)


struct SparseWork {
    uint8_t *flags;              // [N] per pixel: (iteration stamp << FLAG_BITS) | FLAG_*
    BgState *bg;
    unsigned int *bghist;        // [BG_BINS] pixels per float32 key inside the bracket
    unsigned int *bgsample;      // [BG_SAMPLES] float32 keys of the strided sample (0xffffffff: not usable)
    unsigned int *listA[2], *listB, *listC0, *listC1, *listCR;
    unsigned int capA, capB, capC, capCR;
    SparseCounters *cnt;
    SelState *sel;
    float *background;
};

static size_t sp_align(size_t v) { return (v + 255) / 256 * 256; }

static void sparse_caps(size_t n, unsigned int &capA, unsigned int &capB, unsigned int &capC, unsigned int &capCR)
{
    // generous: a typical 10560^2 frame has ~1e4..1e5 entries in A and ~1e4 cosmic-ray pixels
    capA = (unsigned int)(n / 8 + 1024);
    capB = (unsigned int)(n / 32 + 1024);
    capC = (unsigned int)(n / 32 + 1024);
    capCR = (unsigned int)(n / 16 + 1024);
}

size_t lac_sparse_work_bytes(int H, int W)
{
    const size_t n = (size_t)H * W;
    unsigned int a, b, c, cr;
    sparse_caps(n, a, b, c, cr);
    return sp_align(n) + 2 * sp_align(4ull * a) + sp_align(4ull * b) + 2 * sp_align(4ull * c) + sp_align(4ull * cr) +
           sp_align(sizeof(SparseCounters)) + sp_align(sizeof(SelState)) + sp_align(sizeof(BgState)) +
           sp_align(4ull * BG_BINS) + sp_align(4ull * BG_SAMPLES) + 512;
}

static SparseWork carve_sparse(void *work, size_t n)
{
    SparseWork w;
    sparse_caps(n, w.capA, w.capB, w.capC, w.capCR);
    uint8_t *p = (uint8_t *)work;
    w.flags = p; p += sp_align(n);
    w.listA[0] = (unsigned int *)p; p += sp_align(4ull * w.capA);
    w.listA[1] = (unsigned int *)p; p += sp_align(4ull * w.capA);
    w.listB = (unsigned int *)p; p += sp_align(4ull * w.capB);
    w.listC0 = (unsigned int *)p; p += sp_align(4ull * w.capC);
    w.listC1 = (unsigned int *)p; p += sp_align(4ull * w.capC);
    w.listCR = (unsigned int *)p; p += sp_align(4ull * w.capCR);
    w.cnt = (SparseCounters *)p; p += sp_align(sizeof(SparseCounters));
    w.sel = (SelState *)p; p += sp_align(sizeof(SelState));
    w.bg = (BgState *)p; p += sp_align(sizeof(BgState));
    w.bghist = (unsigned int *)p; p += sp_align(4ull * BG_BINS);
    w.bgsample = (unsigned int *)p; p += sp_align(4ull * BG_SAMPLES);
    w.background = (float *)p;
    return w;
}

// --------------------------------------------------------------------------------------------
// appends `value`; returns the slot (>= cap: the list overflowed and the entry was dropped)
__device__ __forceinline__ unsigned int list_push(unsigned int *list, unsigned int *count, unsigned int cap,
                                                  unsigned int value, long long *info)
{
    const unsigned int slot = atomicAdd(count, 1u);
    if (slot < cap) list[slot] = value;
    else atomicOr((unsigned long long *)&info[INFO_STATUS], (unsigned long long)LAC_STATUS_OVERFLOW);
    return slot;
}

// set `bit` in the flag byte of pixel p for iteration stamp `stamp`; returns true if the bit
// was not set before (for this stamp).  32-bit CAS on the word that holds the byte.
__device__ __forceinline__ bool flag_set(uint8_t *flags, size_t p, unsigned int stamp, unsigned int bit)
{
    unsigned int *word = reinterpret_cast<unsigned int *>(flags + (p & ~(size_t)3));
    const unsigned int shift = (unsigned int)(p & 3) * 8;
    unsigned int old = *word;
    for (;;) {
        const unsigned int b = (old >> shift) & 0xffu;
        const unsigned int cur = ((b >> FLAG_BITS) == stamp) ? (b & FLAG_MASK) : 0u;
        if (cur & bit) return false;
        const unsigned int nb = (stamp << FLAG_BITS) | cur | bit;
        const unsigned int nw = (old & ~(0xffu << shift)) | (nb << shift);
        const unsigned int seen = atomicCAS(word, old, nw);
        if (seen == old) return true;
        old = seen;
    }
}

// is `bit` set in the flag byte of pixel p for this stamp (plain read: a stale answer only costs
// a redundant evaluation)
__device__ __forceinline__ bool flag_has(const uint8_t *flags, size_t p, unsigned int stamp, unsigned int bit)
{
    const unsigned int b = flags[p];
    return (b >> FLAG_BITS) == stamp && (b & bit) != 0;
}

__device__ __forceinline__ unsigned int list_len(const unsigned int *count, unsigned int cap)
{
    const unsigned int n = *count;
    return n < cap ? n : cap;
}

// list-A threshold: `L+ / den_min > sigclip` (den_min = 2 sqrt(1e-5 + rn^2), the smallest
// possible 2*noise) is replaced by the division-free superset `L+ > thr_lo`,
// thr_lo = sigclip * den_min * (1 - 2^-20): pixels it lets through in excess are rejected by
// the exact tests of the candidate kernels.
__device__ __forceinline__ float lac_thr_lo(const LacParams &prm)
{
    float nmin = 0.00001f + lac_rn2(prm);
    nmin = sqrtf(nmin);
    const float den_min = 2.0f * nmin;
    return (float)((double)prm.sigclip * (double)den_min * (1.0 - 9.5367431640625e-07));
}

// --------------------------------------------------------------------------------------------
// dense scan (first iteration): 4 pixels per thread, 128-bit loads of the three rows
// --------------------------------------------------------------------------------------------
// A block owns a 512-pixel wide, SCAN_ROWS tall strip: 128 threads x 4 pixels walk down the
// rows keeping the previous / current / next row in registers, so every pixel is loaded once
// per strip (plus one halo row at either end).
//
// The kernel is instruction-issue bound (ncu: 90 % issue slots for the plain 25-operation
// Laplacian), so almost every pixel is dismissed by a cheaper, rigorous bound first.  With
// m1 = min(l, r), m2 = min(u, d) the largest of the four sub-pixel Laplacians is, in exact
// arithmetic, t = 2c - m1 - m2; the float32 evaluation exceeds it by at most 4 roundings of
// intermediates <= 8 mag (mag = largest magnitude involved), i.e. < 2^-18 mag, and the final
// average of the clipped values by a factor < 1 + 2^-22.  t is evaluated with round-up
// operations, so  t_ru <= thr_lo (1 - 2^-21) - 2^-18 mag  implies L+ <= thr_lo.  NaNs fail the
// comparison and take the full evaluation.
//
// COLLECT (mode 0): also the statistics for the background level, see the file header.
#define SCAN_THREADS 128
#define SCAN_ROWS 32
// One pixel the slow way (frame rows / columns, unaligned images)
template <bool COLLECT>
__device__ __forceinline__ void sp_scan_pixel(const float *__restrict__ img, const uint8_t *__restrict__ inmask,
                                              uint8_t *__restrict__ crmask,
                                              int H, int W, int y, int x, float thr_lo, unsigned int key_a,
                                              unsigned int width, unsigned int &n_valid, unsigned int &n_below,
                                              const SparseWork &w, long long *info)
{
    const size_t i = (size_t)y * W + x;
    crmask[i] = 0;                       // the scan also clears the two per-pixel byte maps
    w.flags[i] = 0;
    if (COLLECT && !(inmask && inmask[i])) {
        const unsigned int key = f32_key(img[i]);
        n_valid++;
        if (key < key_a) n_below++;
        else if (key - key_a < width) atomicAdd(&w.bghist[key - key_a], 1u);
    }
    const float lp = laplace_plus_at(img, H, W, y, x);
    if (lp > thr_lo) list_push(w.listA[0], &w.cnt->nA[0], w.capA, (unsigned int)i, info);
}

template <bool COLLECT>
__global__ void __launch_bounds__(SCAN_THREADS, 8)
sp_scan_kernel(const float *__restrict__ img, const uint8_t *__restrict__ inmask, uint8_t *__restrict__ crmask,
               int H, int W, LacParams prm, SparseWork w, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const float thr_lo = lac_thr_lo(prm);
    const float thr_s = __fmul_rd(thr_lo, thr_lo >= 0.f ? 0.99999952316284f : 1.00000047683716f);   // thr_lo (1 -+ 2^-21)
    const int x0 = (blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const int ya = blockIdx.y * SCAN_ROWS, yb = min(ya + SCAN_ROWS, H);
    unsigned int n_valid = 0, n_below = 0;
    unsigned int key_a = 0, width = 0;
    if (COLLECT) { key_a = w.bg->key_a; width = w.bg->width; }
    // vector path: aligned rows, the 4 pixels and their left / right neighbours inside the row
    const bool fast = (W % 4 == 0) && (((uintptr_t)img & 15) == 0) && x0 > 0 && x0 + 4 < W &&
                      (((uintptr_t)crmask | (uintptr_t)w.flags) & 3) == 0 &&
                      (!COLLECT || inmask == nullptr || ((uintptr_t)inmask & 3) == 0);
    if (x0 < W && !fast) {
        for (int y = ya; y < yb; y++)
            for (int k = 0; k < 4 && x0 + k < W; k++)
                sp_scan_pixel<COLLECT>(img, inmask, crmask, H, W, y, x0 + k, thr_lo, key_a, width, n_valid, n_below, w, info);
    } else if (x0 < W) {
        int y0 = ya, y1 = yb;
        if (y0 == 0) {                                     // first / last image row: no vector path
            for (int k = 0; k < 4; k++)
                sp_scan_pixel<COLLECT>(img, inmask, crmask, H, W, 0, x0 + k, thr_lo, key_a, width, n_valid, n_below, w, info);
            y0 = 1;
        }
        if (y1 == H) {
            y1 = H - 1;
            if (y1 >= y0)
                for (int k = 0; k < 4; k++)
                    sp_scan_pixel<COLLECT>(img, inmask, crmask, H, W, H - 1, x0 + k, thr_lo, key_a, width, n_valid, n_below, w, info);
        }
        if (y0 < y1) {
            unsigned int pix = (unsigned int)((size_t)y0 * W + x0);
            // background statistics on the raw float bits when the bracket starts at a
            // non-negative value (key_a >= 0x80000000): key < key_a  <=>  (int)bits < (int)lo_bits,
            // key - key_a = bits - lo_bits for every value at or above the bracket's lower end,
            // and a negative value's difference wraps far beyond any bracket width (< 2^20)
            const bool raw_bits = COLLECT && key_a >= 0x80000000u;
            const int lo_bits = (int)(key_a & 0x7fffffffu);
            auto mag4 = [](const float4 &v) { return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))); };
            // A row bundle: the 4 pixels, their left / right neighbours and the mask word, all
            // loaded when the row first comes into reach (two rows before it is the centre row,
            // so every load has a full row of arithmetic to hide behind); four bundles rotate
            // through the roles next -> down -> centre -> up without register moves.
            struct Row { float4 v; float l, r; unsigned int mm; };
            auto load_row = [&](Row &q, int yy) {
                if (yy >= H) return;                             // beyond the frame: never read
                const float *src = img + (size_t)yy * W + x0;
                q.v = *reinterpret_cast<const float4 *>(src);
                q.l = src[-1]; q.r = src[4];
                q.mm = 0;
                if (COLLECT && inmask) q.mm = *reinterpret_cast<const unsigned int *>(inmask + (size_t)yy * W + x0);
            };
            auto proc = [&](const Row &U, const Row &Cn, const Row &D) {
                const float4 &up = U.v, &cur = Cn.v, &dn = D.v;
                const float lft = Cn.l, rgt = Cn.r;
                const unsigned int mm = Cn.mm;
                const float mag_up = mag4(up), mag_cur = mag4(cur), mag_dn = mag4(dn);
                *reinterpret_cast<unsigned int *>(crmask + pix) = 0u;
                *reinterpret_cast<unsigned int *>(w.flags + pix) = 0u;
                const float cc[6] = {lft, cur.x, cur.y, cur.z, cur.w, rgt};
                const float uu[4] = {up.x, up.y, up.z, up.w}, dd[4] = {dn.x, dn.y, dn.z, dn.w};
                if (COLLECT) {
                    if (raw_bits) {
                        n_valid += (unsigned int)__popc(__vcmpeq4(mm, 0u)) >> 3;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            // a masked pixel is given the bits of +NaN: neither below nor inside
                            const int bits = ((mm >> (8 * k)) & 0xffu) ? 0x7fffffff : __float_as_int(cc[k + 1]);
                            n_below += bits < lo_bits;
                            const unsigned int d = (unsigned int)bits - (unsigned int)lo_bits;
                            if (d < width) atomicAdd(&w.bghist[d], 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const bool ok = ((mm >> (8 * k)) & 0xffu) == 0;
                            const unsigned int key = f32_key(cc[k + 1]);
                            const unsigned int d = key - key_a;          // wraps to a huge value below the bracket
                            n_valid += ok;
                            n_below += ok && key < key_a;
                            if (ok && d < width) atomicAdd(&w.bghist[d], 1u);
                        }
                    }
                }
                const float mag = fmaxf(fmaxf(mag_up, mag_dn), fmaxf(mag_cur, fmaxf(fabsf(lft), fabsf(rgt))));
                const float thr_row = __fmaf_rd(mag, -3.814697265625e-06f, thr_s);          // - 2^-18 mag
                float t[4];
                bool any = false;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    t[k] = __fsub_ru(cc[k + 1], fminf(cc[k], cc[k + 2]));
                    t[k] = __fadd_ru(t[k], cc[k + 1]);
                    t[k] = __fsub_ru(t[k], fminf(uu[k], dd[k]));
                    any |= !(t[k] <= thr_row);
                }
                if (any) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (t[k] <= thr_row) continue;
                        const float cv = cc[k + 1], l = cc[k], r = cc[k + 2], c4 = 4.0f * cv;
                        float s00 = c4 - cv; s00 = s00 - l; s00 = s00 - cv; s00 = s00 - uu[k];
                        float s01 = c4 - r; s01 = s01 - cv; s01 = s01 - cv; s01 = s01 - uu[k];
                        float s10 = c4 - cv; s10 = s10 - l; s10 = s10 - dd[k]; s10 = s10 - cv;
                        float s11 = c4 - r; s11 = s11 - cv; s11 = s11 - dd[k]; s11 = s11 - cv;
                        s00 = fmaxf(s00, 0.f); s01 = fmaxf(s01, 0.f); s10 = fmaxf(s10, 0.f); s11 = fmaxf(s11, 0.f);
                        float p = s00 + s01; p = p + s10; p = p + s11;
                        const float lp = p * 0.25f;
                        if (lp > thr_lo) list_push(w.listA[0], &w.cnt->nA[0], w.capA, pix + k, info);
                    }
                }
                pix += (unsigned int)W;
            };
            Row ra, rb, rc, rd;
            load_row(ra, y0 - 1); load_row(rb, y0); load_row(rc, y0 + 1);
            int y = y0;
            for (; y + 4 <= y1; y += 4) {
                load_row(rd, y + 2); proc(ra, rb, rc);
                load_row(ra, y + 3); proc(rb, rc, rd);
                load_row(rb, y + 4); proc(rc, rd, ra);
                load_row(rc, y + 5); proc(rd, ra, rb);
            }
            if (y < y1) { load_row(rd, y + 2); proc(ra, rb, rc); y++; }
            if (y < y1) { load_row(ra, y + 2); proc(rb, rc, rd); y++; }
            if (y < y1) { proc(rc, rd, ra); y++; }
        }
    }
    if (COLLECT) {
        const unsigned int tv = (unsigned int)warp_sum((int)n_valid), tb = (unsigned int)warp_sum((int)n_below);
        if ((threadIdx.x & 31) == 0) {
            if (tv) atomicAdd(&w.bg->n_valid, (unsigned long long)tv);
            if (tb) atomicAdd(&w.bg->n_below, (unsigned long long)tb);
        }
    }
}

// --------------------------------------------------------------------------------------------
// The fused per-pixel pass (apply.cu) and the dense scan above in ONE kernel: the reduced image is
// scanned for Laplacian candidates while it is being made, so LACosmic's first iteration does not
// read the 446 MB it has just seen written (the scan alone: 0.25 ms per frame, the largest item
// of the chain after the fused pass itself).
//   * a thread owns 4 columns and walks down FUSE_ROWS rows plus one halo row at either end,
//     holding the reduced values of three consecutive rows in registers; left / right neighbours
//     come from the neighbouring lanes by shuffle.  The first and the last lane of a warp are
//     halo lanes: they compute their 4 columns like everybody else but own nothing (a warp covers
//     128 columns and owns 120), so no thread ever needs a pixel outside its warp.  (The first
//     version had the edge lanes evaluate the missing neighbour pixel by pixel: two dependent
//     chains of scattered loads per row and warp, 2.2 ms per frame.);
//   * the statistics of the background level are taken against the SEED mask (bad-pixel mask,
//     non-finite, saturated: all the fused pass knows); the mask morphology that follows takes every
//     pixel it masks for the first time out again (bg_untrack, called where mask.cu sets a bit in
//     a zero byte) -- the final mask is what detect_cosmics' inmask is;
//   * the bracket of the background level comes from a sample evaluated through apply_value_at
//     BEFORE the pass (sp_bg_gather_raw_kernel).
// --------------------------------------------------------------------------------------------
#define FUSE_THREADS 128
#define FUSE_ROWS 64

struct FuseArgs {
    uint8_t *crmask;
    LacParams prm;
    SparseWork w;
    long long *info;
};

// background statistics of one unmasked pixel value (the accounting of sp_scan_kernel<true>)
__device__ __forceinline__ void bg_count(float v, unsigned int key_a, unsigned int width, unsigned int &n_valid,
                                         unsigned int &n_below, unsigned int *__restrict__ hist)
{
    const unsigned int key = f32_key(v);
    n_valid++;
    if (key < key_a) n_below++;
    else if (key - key_a < width) atomicAdd(&hist[key - key_a], 1u);
}

template <typename T>
__global__ void __launch_bounds__(FUSE_THREADS, 6)
reduce_apply_scan_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain, ApplyArgs a, FuseArgs f)
{
    const int RW = g.nx * g.xsize_chan, RH = g.ny * g.ysize_chan;
    const int lane = threadIdx.x & 31;
    // warp w of the grid row covers the 4-column groups 30 w - 1 .. 30 w + 30 and owns 30 w .. 30 w + 29
    const int gwarp = blockIdx.x * (FUSE_THREADS / 32) + (threadIdx.x >> 5);
    const int x = (gwarp * 30 + lane - 1) * 4;
    const bool live = x >= 0 && x < RW;                        // whole warps stay: the lanes shuffle
    const bool own_col = live && lane >= 1 && lane <= 30;
    const int xs = live ? x : 0;
    const int c = xs / g.xsize_chan, lx = xs - c * g.xsize_chan;
    const int ya = blockIdx.y * FUSE_ROWS, yb = min(ya + FUSE_ROWS, RH);
    const bool has_bias = a.mbias != nullptr, has_flat = a.mflat != nullptr;
    const unsigned int key_a = f.w.bg->key_a, width = f.w.bg->width;
    const bool raw_bits = key_a >= 0x80000000u;
    const int lo_bits = (int)(key_a & 0x7fffffffu);
    const float thr_lo = lac_thr_lo(f.prm);
    const float thr_s = __fmul_rd(thr_lo, thr_lo >= 0.f ? 0.99999952316284f : 1.00000047683716f);   // thr_lo (1 -+ 2^-21)
    unsigned int n_valid = 0, n_below = 0;
    int r_cur = -1;
    double osc[4] = {0.0, 0.0, 0.0, 0.0};
    float gn = 1.0f, satl = 0.0f;
    bool has_sat = false;

    struct RowIn { float v[4]; float4 mb, mf; uint32_t mm; double fitv; };
    struct RowV { float v[4]; float l, r; };
    auto load_row = [&](int y, RowIn &in) {
        in.mb = make_float4(0.f, 0.f, 0.f, 0.f);
        in.mf = make_float4(1.f, 1.f, 1.f, 1.f);
        in.mm = 0;
        in.fitv = 0.0;
        in.v[0] = in.v[1] = in.v[2] = in.v[3] = 0.f;
        if (!live) return;
        const int r = (y >= g.ysize_chan) ? 1 : 0;             // ny == 2
        const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + (y - r * g.ysize_chan);
        const size_t ro = (size_t)rr * g.W + (size_t)c * g.dx + lx;
        const size_t oo = (size_t)y * RW + x;
        RawVec4<T>::load(raw + ro, in.v);
        if (has_bias) in.mb = __ldcs(reinterpret_cast<const float4 *>(a.mbias + oo));
        if (has_flat) in.mf = __ldcs(reinterpret_cast<const float4 *>(a.mflat + oo));
        if (a.bpm) in.mm = __ldcs(reinterpret_cast<const unsigned int *>(a.bpm + oo));
        in.fitv = a.vos_fit ? a.vos_fit[(size_t)(r * g.nx + c) * g.dy + (rr - r * g.dy)] : 0.0;
    };
    // the reduced values of row y (and, for the strip's own rows, everything the fused pass
    // writes); `own`: the row belongs to this strip (halo rows are only computed)
    auto finish_row = [&](int y, const RowIn &in, bool own, RowV &out) {
        uint32_t mout = 0;
        if (live) {
            const int r = (y >= g.ysize_chan) ? 1 : 0;
            if (r != r_cur) {                                  // once per strip (twice if it straddles the CCD halves)
                r_cur = r;
                const int ch = r * g.nx + c;
                gn = gain.v[ch];
                if (a.oscan) {
#pragma unroll
                    for (int k = 0; k < 4; k++) osc[k] = a.oscan[(size_t)ch * g.xsize_chan + lx + k];
                }
                has_sat = a.satlevel != nullptr;
                if (has_sat) {
                    const double lv = a.satlevel[ch];
                    if (lv != lv) has_sat = false;             // NaN level: the comparison is never true
                    else satl = f32_ceil_of(lv);
                }
            }
            const float mb[4] = {in.mb.x, in.mb.y, in.mb.z, in.mb.w}, mf[4] = {in.mf.x, in.mf.y, in.mf.z, in.mf.w};
            bool any_seed = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float w = in.v[k] * gn;
                w = sub_f64(w, in.fitv);
                w = sub_f64(w, osc[k]);
                if (has_bias) w = w - mb[k];
                uint32_t m = (in.mm >> (8 * k)) & 0xffu;
                if (!isfinite(w)) { w = 0.f; if (m == 0) m |= (uint32_t)a.bit_bad; }
                if (has_sat && w >= satl) m |= (uint32_t)(a.bit_sat | BBX_TMP_SAT);
                any_seed |= (m & (BBX_TMP_SAT | a.seed_bits)) != 0;
                if (has_flat) w = w / mf[k];
                out.v[k] = w;
                mout |= m << (8 * k);
            }
            if (own && own_col) {
                const size_t oo = (size_t)y * RW + x;
                if (a.seeds && any_seed) {
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
                __stcs(reinterpret_cast<float4 *>(a.out_img + oo), make_float4(out.v[0], out.v[1], out.v[2], out.v[3]));
                __stcs(reinterpret_cast<unsigned int *>(a.out_mask + oo), mout);
                // what the dense scan cleared: the cosmic-ray mask and LACosmic's flag bytes
                __stcs(reinterpret_cast<unsigned int *>(f.crmask + oo), 0u);
                __stcs(reinterpret_cast<unsigned int *>(f.w.flags + oo), 0u);
                // background statistics against the seed mask.  Nearly every group is unmasked and
                // the bracket starts at a non-negative value: then the order-preserving keys are the
                // raw float bits (see sp_scan_kernel) and the four pixels cost four compares
                if (mout == 0 && raw_bits) {
                    n_valid += 4;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int bits = __float_as_int(out.v[k]);
                        n_below += bits < lo_bits;
                        const unsigned int d = (unsigned int)bits - (unsigned int)lo_bits;
                        if (d < width) atomicAdd(&f.w.bghist[d], 1u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (((mout >> (8 * k)) & 0xffu) == 0) bg_count(out.v[k], key_a, width, n_valid, n_below, f.w.bghist);
                }
            }
        } else {
            out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.f;
        }
        // horizontal neighbours: the lanes either side (the owning lanes 1 .. 30 always have both)
        out.l = __shfl_up_sync(0xffffffffu, out.v[3], 1);
        out.r = __shfl_down_sync(0xffffffffu, out.v[0], 1);
    };
    // L+ of the 4 pixels of row y: U / D are the rows above / below (hu / hd: they exist)
    auto laplace_row = [&](int y, const RowV &U, const RowV &Cn, const RowV &D, bool hu, bool hd) {
        if (!own_col) return;
        const unsigned int pix = (unsigned int)((size_t)y * RW + x);
        const float cc[6] = {Cn.l, Cn.v[0], Cn.v[1], Cn.v[2], Cn.v[3], Cn.r};
        if (hu && hd && x > 0 && x + 4 < RW) {
            // interior: the cheap rigorous bound first (see sp_scan_kernel), the full Laplacian for the few it lets through
            auto mag4 = [](const float (&v)[4]) { return fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))); };
            const float mag = fmaxf(fmaxf(mag4(U.v), mag4(D.v)), fmaxf(mag4(Cn.v), fmaxf(fabsf(Cn.l), fabsf(Cn.r))));
            const float thr_row = __fmaf_rd(mag, -3.814697265625e-06f, thr_s);            // - 2^-18 mag
            float t[4];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                t[k] = __fsub_ru(cc[k + 1], fminf(cc[k], cc[k + 2]));
                t[k] = __fadd_ru(t[k], cc[k + 1]);
                t[k] = __fsub_ru(t[k], fminf(U.v[k], D.v[k]));
                any |= !(t[k] <= thr_row);
            }
            if (!any) return;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (t[k] <= thr_row) continue;
                const float cv = cc[k + 1], l = cc[k], r = cc[k + 2], c4 = 4.0f * cv;
                float s00 = c4 - cv; s00 = s00 - l; s00 = s00 - cv; s00 = s00 - U.v[k];
                float s01 = c4 - r; s01 = s01 - cv; s01 = s01 - cv; s01 = s01 - U.v[k];
                float s10 = c4 - cv; s10 = s10 - l; s10 = s10 - D.v[k]; s10 = s10 - cv;
                float s11 = c4 - r; s11 = s11 - cv; s11 = s11 - D.v[k]; s11 = s11 - cv;
                s00 = fmaxf(s00, 0.f); s01 = fmaxf(s01, 0.f); s10 = fmaxf(s10, 0.f); s11 = fmaxf(s11, 0.f);
                float p = s00 + s01; p = p + s10; p = p + s11;
                const float lp = p * 0.25f;
                if (lp > thr_lo) list_push(f.w.listA[0], &f.w.cnt->nA[0], f.w.capA, pix + k, f.info);
            }
            return;
        }
        // the frame of the image: missing neighbours drop out of the sub-pixel Laplacians (laplace_plus_at)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool hl = x + k > 0, hr = x + k + 1 < RW;
            const float cv = cc[k + 1], l = cc[k], r = cc[k + 2], u = U.v[k], d = D.v[k], c4 = 4.0f * cv;
            float s00 = c4 - cv; if (hl) s00 = s00 - l; s00 = s00 - cv; if (hu) s00 = s00 - u;
            float s01 = c4; if (hr) s01 = s01 - r; s01 = s01 - cv; s01 = s01 - cv; if (hu) s01 = s01 - u;
            float s10 = c4 - cv; if (hl) s10 = s10 - l; if (hd) s10 = s10 - d; s10 = s10 - cv;
            float s11 = c4; if (hr) s11 = s11 - r; s11 = s11 - cv; if (hd) s11 = s11 - d; s11 = s11 - cv;
            s00 = s00 < 0.f ? 0.f : s00; s01 = s01 < 0.f ? 0.f : s01;
            s10 = s10 < 0.f ? 0.f : s10; s11 = s11 < 0.f ? 0.f : s11;
            float p = s00 + s01;
            p = p + s10;
            p = p + s11;
            const float lp = p / 4.0f;
            if (lp > thr_lo) list_push(f.w.listA[0], &f.w.cnt->nA[0], f.w.capA, pix + k, f.info);
        }
    };

    // rows ya-1 .. yb: row y is finished into R, then the Laplacian of row y-1 (Q) is taken between
    // P (row y-2) and R.  ONE copy of the loop body (the rows move through P, Q, R by register
    // copies): unrolled over the rotation it was 11 copies of ~1000 instructions and ran out of the
    // instruction cache.  Row RH (below the image) is a virtual row: nothing to finish, it only
    // triggers the Laplacian of the last image row, which has no row below it.
    const int y0 = max(ya - 1, 0), y1 = yb;
    RowIn cur, nxt;
    RowV P, Q, R;
    P.v[0] = P.v[1] = P.v[2] = P.v[3] = P.l = P.r = 0.f;
    Q = P; R = P;
    load_row(y0, cur);
#pragma unroll 1
    for (int y = y0; y <= y1; y++) {
        if (y + 1 <= y1 && y + 1 < RH) load_row(y + 1, nxt);
        if (y < RH) finish_row(y, cur, y >= ya && y < yb, R);
        const int ym = y - 1;
        if (ym >= ya && ym >= y0) laplace_row(ym, P, Q, R, ym > 0, y < RH);
        P = Q; Q = R; cur = nxt;
    }
    const unsigned int tv = (unsigned int)warp_sum((int)n_valid), tb = (unsigned int)warp_sum((int)n_below);
    if (lane == 0) {
        if (tv) atomicAdd(&f.w.bg->n_valid, (unsigned long long)tv);
        if (tb) atomicAdd(&f.w.bg->n_below, (unsigned long long)tb);
    }
}

// the strided sample for the bracket of the background level, evaluated on the RAW frame through
// the fused pass's own arithmetic (the reduced image does not exist yet); seed mask
template <typename T>
__global__ void __launch_bounds__(1024)
sp_bg_gather_raw_kernel(const T *__restrict__ raw, bbx_geom g, ChanF32 gain, ApplyArgs a, SparseWork w)
{
    const int RW = g.nx * g.xsize_chan, RH = g.ny * g.ysize_chan;
    const size_t n = (size_t)RW * RH;
    const size_t stride = n / BG_SAMPLES > 0 ? n / BG_SAMPLES : 1;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= BG_SAMPLES) return;
    const size_t i = j * stride;
    unsigned int key = 0xffffffffu;
    if (i < n) {
        uint32_t m;
        const float v = apply_value_at<T>(raw, g, gain, a, (int)(i / RW), (int)(i % RW), m);
        if (m == 0 && v == v) key = f32_key(v);
    }
    w.bgsample[j] = key;
}

// ---- background level ----------------------------------------------------------------------
// One block: a strided sample of BG_SAMPLES unmasked pixels brackets the median rank by
// +-(2.625 sqrt(ns) + 8) sample ranks (5.25 sigma of the binomial sampling error of the median's
// rank).  The two bracketing sample order statistics are located with a two-pass radix select
// (11 + 11 key bits) on the sample held in shared memory; the remaining 10 bits are rounded
// outwards, which widens the bracket by at most 2 x 1024 of its ~1e5 keys.
#define BG_SEL_BINS 2048

__device__ __forceinline__ void bg_find_bin(const unsigned int *hist, unsigned int rank, unsigned int *part,
                                            unsigned int *out_bin, unsigned int *out_rank)
{
    // 1024 threads: warp q sums bins [64 q, 64 q + 64); thread 0 walks the 32 warp sums, then
    // the 64 bins of the warp that holds the rank
    const int t = threadIdx.x, lane = t & 31, wq = t >> 5;
    const unsigned int mine = (unsigned int)warp_sum((int)(hist[2 * t] + hist[2 * t + 1]));
    if (lane == 0) part[wq] = mine;
    __syncthreads();
    if (t == 0) {
        unsigned int acc = 0;
        int seg = 31;
        for (int q = 0; q < 32; q++) {
            if (acc + part[q] > rank) { seg = q; break; }
            acc += part[q];
        }
        unsigned int bin = 64 * seg + 63;
        for (int j = 0; j < 64; j++) {
            const unsigned int hv = hist[64 * seg + j];
            if (acc + hv > rank) { bin = 64 * seg + j; break; }
            acc += hv;
        }
        *out_bin = bin;
        *out_rank = rank - acc;
    }
    __syncthreads();
}

// The sample is gathered by a kernel of its own, one pixel per thread over 32 CTAs: 65536 scattered
// sector reads are a few microseconds spread over the chip and 50+ when one SM issues them all
// (round 1: 79 us on one SM in front of every frame).
__global__ void __launch_bounds__(1024)
sp_bg_gather_kernel(const float *__restrict__ img, const uint8_t *__restrict__ inmask, size_t n, SparseWork w)
{
    const size_t stride = n / BG_SAMPLES > 0 ? n / BG_SAMPLES : 1;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= BG_SAMPLES) return;
    const size_t i = j * stride;
    unsigned int key = 0xffffffffu;
    if (i < n && !(inmask && inmask[i])) {
        const float v = img[i];
        if (v == v) key = f32_key(v);                   // NaNs are left to the full-frame statistics
    }
    w.bgsample[j] = key;                                // (no finite float has the key 0xffffffff)
}

__global__ void __launch_bounds__(1024)
sp_bg_sample_kernel(SparseWork w)
{
    extern __shared__ unsigned int smp[];               // [BG_SAMPLES] float32 keys of the sample
    __shared__ unsigned int hist[2][BG_SEL_BINS];
    __shared__ unsigned int part[1024];
    __shared__ unsigned int s_ns, s_bin[2], s_rank[2], s_sub[2], s_dummy;
    const int t = threadIdx.x;
    if (t == 0) s_ns = 0;
    for (int i = t; i < 2 * BG_SEL_BINS; i += 1024) (&hist[0][0])[i] = 0;
    __syncthreads();
    {
        // 32 independent coalesced loads per thread, then: the top 11 key bits are the same for
        // practically the whole sample of a sky-dominated frame -- count runs per thread instead
        // of hammering one shared-memory word
        unsigned int keys[BG_SAMPLES / 1024];
#pragma unroll
        for (int q = 0; q < (int)(BG_SAMPLES / 1024); q++) keys[q] = w.bgsample[q * 1024 + t];
        unsigned int run_bin = 0xffffffffu, run = 0;
#pragma unroll
        for (int q = 0; q < (int)(BG_SAMPLES / 1024); q++) {
            const unsigned int key = keys[q];
            // warp-aggregated append: one shared-memory atomic per warp instead of one per sample
            const bool usable = key != 0xffffffffu;
            const unsigned int vote = __ballot_sync(0xffffffffu, usable);
            unsigned int base = 0;
            if ((t & 31) == 0 && vote) base = atomicAdd(&s_ns, (unsigned int)__popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!usable) continue;
            smp[base + __popc(vote & ((1u << (t & 31)) - 1u))] = key;
            if ((key >> 21) == run_bin) { run++; continue; }
            if (run) atomicAdd(&hist[0][run_bin], run);
            run_bin = key >> 21; run = 1;
        }
        if (run) atomicAdd(&hist[0][run_bin], run);
    }
    __syncthreads();
    const unsigned int ns = s_ns;
    unsigned int key_a = 0, width = 0;
    bool ok = false;
    unsigned int mid = 0, d = 0;
    if (ns > 0) {
        mid = (ns - 1) / 2;
        d = (unsigned int)(2.625f * sqrtf((float)ns)) + 8u;
        ok = mid > d && mid + d < ns - 1;
    }
    if (ok) {                                           // block-uniform
        // pass 1: the top 11 bits of both order statistics
        bg_find_bin(hist[0], mid - d, part, &s_bin[0], &s_rank[0]);
        bg_find_bin(hist[0], mid + d, part, &s_bin[1], &s_rank[1]);
        const unsigned int b0 = s_bin[0], b1 = s_bin[1];
        for (int i = t; i < 2 * BG_SEL_BINS; i += 1024) (&hist[0][0])[i] = 0;
        __syncthreads();
        // pass 2: the next 11 bits inside those two bins
        for (unsigned int j = t; j < ns; j += 1024) {
            const unsigned int key = smp[j], top = key >> 21, sub = (key >> 10) & 2047u;
            if (top == b0) atomicAdd(&hist[0][sub], 1u);
            if (top == b1) atomicAdd(&hist[1][sub], 1u);
        }
        __syncthreads();
        bg_find_bin(hist[0], s_rank[0], part, &s_sub[0], &s_dummy);
        bg_find_bin(hist[1], s_rank[1], part, &s_sub[1], &s_dummy);
        if (t == 0) {
            const unsigned int ka = (b0 << 21) | (s_sub[0] << 10);                    // rounded down
            const unsigned int kb = (b1 << 21) | (s_sub[1] << 10) | 1023u;            // rounded up
            if (kb >= ka && kb - ka < BG_BINS) { key_a = ka; width = kb - ka + 1u; }
        }
    }
    if (t == 0) {
        w.bg->key_a = key_a; w.bg->width = width;
        w.bg->n_valid = 0; w.bg->n_below = 0;
    }
}

// after the scan: the key whose cumulative count reaches the median rank
__global__ void __launch_bounds__(1024)
sp_bg_rank_kernel(SparseWork w)
{
    __shared__ unsigned long long s_tot[32];
    __shared__ unsigned long long s_k;
    __shared__ int s_warp;
    const unsigned int width = w.bg->width;
    const unsigned long long nv = w.bg->n_valid, nb = w.bg->n_below;
    if (width == 0 || nv == 0) return;                         // bg_valid stays 0
    const unsigned long long want = (nv - 1) / 2;               // lower median
    if (want < nb) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int chunk = (width + 31) / 32;               // bins per warp
    const unsigned int b0 = warp * chunk, b1 = min(b0 + chunk, width);
    unsigned long long mine = 0;
    for (unsigned int b = b0 + lane; b < b1; b += 32) mine += w.bghist[b];
    mine = warp_sum(mine);
    if (lane == 0) s_tot[warp] = mine;
    if (threadIdx.x == 0) s_warp = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = want - nb, acc = 0;
        for (int q = 0; q < 32; q++) {
            if (acc + s_tot[q] > k) { s_warp = q; s_k = k - acc; break; }
            acc += s_tot[q];
        }
    }
    __syncthreads();
    if (warp != s_warp) return;                                 // rank outside the bracket: bg_valid stays 0
    unsigned long long k = s_k;
    for (unsigned int base = b0; base < b1; base += 32) {
        const unsigned int b = base + lane;
        const unsigned long long c = b < b1 ? w.bghist[b] : 0ull;
        // inclusive prefix over the lanes
        unsigned long long pre = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const unsigned int hit = __ballot_sync(0xffffffffu, pre > k);
        if (hit) {
            const int first = __ffs(hit) - 1;
            if (lane == first) {
                *w.background = key_f32(w.bg->key_a + b);
                w.cnt->bg_valid = 1;
            }
            return;
        }
        k -= __shfl_sync(0xffffffffu, pre, 31);
    }
}

// iterations >= 1: list A from the previous list A and the cross neighbourhood of every pixel
// the cleaning step has rewritten (the cumulative CR list); FLAG_A removes duplicates
__global__ void __launch_bounds__(128)
sp_rescan_kernel(const float *__restrict__ img, int H, int W, LacParams prm, SparseWork w, int it, unsigned int stamp,
                 long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const int cur = it & 1, prev = cur ^ 1;
    const unsigned int nprev = list_len(&w.cnt->nA[prev], w.capA), ncr = list_len(&w.cnt->nCR, w.capCR);
    const float thr_lo = lac_thr_lo(prm);
    const unsigned long long total = (unsigned long long)nprev + 5ull * ncr;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        int y, x;
        if (t < nprev) {
            const unsigned int q = w.listA[prev][t];
            y = (int)(q / (unsigned int)W); x = (int)(q - (unsigned int)y * (unsigned int)W);
        } else {
            const unsigned long long u = t - nprev;
            const unsigned int p = w.listCR[u / 5];
            const int k = (int)(u % 5);
            y = (int)(p / (unsigned int)W); x = (int)(p - (unsigned int)y * (unsigned int)W);
            if (k == 1) x--; else if (k == 2) x++; else if (k == 3) y--; else if (k == 4) y++;
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
        }
        const float lp = laplace_plus_at(img, H, W, y, x);
        if (!(lp > thr_lo)) continue;
        const size_t q = (size_t)y * W + x;
        if (flag_set(w.flags, q, stamp, FLAG_A)) list_push(w.listA[cur], &w.cnt->nA[cur], w.capA, (unsigned int)q, info);
    }
}

// list A -> list B: unmasked pixels whose true Laplacian S/N exceeds sigclip
__global__ void __launch_bounds__(128)
sp_cand1_kernel(const float *__restrict__ img, const uint8_t *__restrict__ inmask, int H, int W, LacParams prm,
                SparseWork w, int it, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int n = list_len(&w.cnt->nA[it & 1], w.capA);
    const unsigned int *__restrict__ listA = w.listA[it & 1];
    const float rn2 = lac_rn2(prm);
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int p = listA[k];
        if (inmask && inmask[p]) continue;
        const int y = (int)(p / (unsigned int)W), x = (int)(p - (unsigned int)y * (unsigned int)W);
        // s' = s - med5(s) is 0 in the 2-pixel frame (the median filter copies its input there)
        if (y < 2 || y >= H - 2 || x < 2 || x >= W - 2) continue;
        float nz;
        const float s = lac_s_at(img, H, W, y, x, rn2, nz);
        if (s > prm.sigclip) list_push(w.listB, &w.cnt->nB, w.capB, p, info);
    }
}

// s' of pixel (y,x) by one warp: lanes 0..24 evaluate s on the 5x5 neighbourhood.
// Returns s' in every lane; s_c / noise_c = s and noise of the centre pixel.
__device__ __forceinline__ float warp_sprime(const float *__restrict__ img, int H, int W, int y, int x, float rn2,
                                             int lane, float &s_c, float &noise_c)
{
    float sv = 0.f, nz = 0.f;
    if (lane < 25) {
        const int qy = y + lane / 5 - 2, qx = x + lane % 5 - 2;
        if (qy >= 0 && qy < H && qx >= 0 && qx < W) sv = lac_s_at(img, H, W, qy, qx, rn2, nz);
    }
    s_c = __shfl_sync(0xffffffffu, sv, 12);
    noise_c = __shfl_sync(0xffffffffu, nz, 12);
    if (y < 2 || y >= H - 2 || x < 2 || x >= W - 2) return s_c - s_c;      // frame: med5 copies its input
    float v[25];
#pragma unroll
    for (int k = 0; k < 25; k++) v[k] = __shfl_sync(0xffffffffu, sv, k);
    const float med = bbx_med25(v);
    return s_c - med;
}

// fine-structure value f of pixel (y,x) by one warp (med3 on 7x7, then their median)
__device__ __forceinline__ float warp_fine(const float *__restrict__ img, int H, int W, int y, int x, float noise_c,
                                           int lane)
{
    float a = 0.f, b = 0.f;
    {
        const int k0 = lane, k1 = lane + 32;
        const int qy0 = y + k0 / 7 - 3, qx0 = x + k0 % 7 - 3;
        if (qy0 >= 0 && qy0 < H && qx0 >= 0 && qx0 < W) a = median_at<3>(img, H, W, qy0, qx0);
        if (k1 < 49) {
            const int qy1 = y + k1 / 7 - 3, qx1 = x + k1 % 7 - 3;
            if (qy1 >= 0 && qy1 < H && qx1 >= 0 && qx1 < W) b = median_at<3>(img, H, W, qy1, qx1);
        }
    }
    const float f3c = __shfl_sync(0xffffffffu, a, 24);           // centre = index 3*7+3
    float m7;
    if (y < 3 || y >= H - 3 || x < 3 || x >= W - 3) {
        m7 = f3c;                                                 // frame: med7 copies its input
    } else {
        float v[49];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = __shfl_sync(0xffffffffu, a, k);
#pragma unroll
        for (int k = 32; k < 49; k++) v[k] = __shfl_sync(0xffffffffu, b, k - 32);
        m7 = bbx_med49(v);
    }
    float f = f3c - m7;
    f = f / noise_c;
    if (f < 0.01f) f = 0.01f;
    return f;
}

// list B -> c0: warp per pixel
__global__ void __launch_bounds__(128)
sp_cand2_kernel(const float *__restrict__ img, int H, int W, LacParams prm, SparseWork w, unsigned int stamp,
                long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int n = list_len(&w.cnt->nB, w.capB);
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float rn2 = lac_rn2(prm);
    for (unsigned int k = warp; k < n; k += nwarps) {
        const unsigned int p = w.listB[k];
        const int y = (int)(p / (unsigned int)W), x = (int)(p - (unsigned int)y * (unsigned int)W);
        float s_c, nz_c;
        const float sp = warp_sprime(img, H, W, y, x, rn2, lane, s_c, nz_c);
        if (!(sp > prm.sigclip)) continue;                        // warp-uniform
        if (lane == 0) flag_set(w.flags, p, stamp, FLAG_S1);
        const float f = warp_fine(img, H, W, y, x, nz_c, lane);
        const float ratio = sp / f;
        if (ratio > prm.objlim && lane == 0) {
            if (flag_set(w.flags, p, stamp, FLAG_C0)) list_push(w.listC0, &w.cnt->nC0, w.capC, p, info);
        }
    }
}

// The two growth steps.  For every listed pixel r and each of its 9 neighbours q (dilate3 copies
// its input in the 1-pixel frame, so a frame pixel q only counts for q == r) the reference tests
// good(q) & s'(q) > thr.  No kernel below lets a racy flag read steer a warp collective: the
// per-neighbour kernels are thread-per-item, the warp-per-pixel kernel reads no flags.
__device__ __forceinline__ bool grow_neighbour(unsigned int r, int k, int H, int W, size_t &q, bool &centre)
{
    const int ry = (int)(r / (unsigned int)W), rx = (int)(r - (unsigned int)ry * (unsigned int)W);
    const int qy = ry + k / 3 - 1, qx = rx + k % 3 - 1;
    if (qy < 0 || qy >= H || qx < 0 || qx >= W) return false;
    centre = (k == 4);
    if (!centre && (qy == 0 || qy == H - 1 || qx == 0 || qx == W - 1)) return false;
    q = (size_t)qy * W + qx;
    return true;
}

// a pixel has passed growth step 2: count it, mark it, list it (once per LACosmic call)
__device__ __forceinline__ void grow_accept(uint8_t *__restrict__ crmask, size_t q, const SparseWork &w, int iter,
                                            long long *info)
{
    atomicAdd((unsigned long long *)&info[INFO_NCR + iter], 1ull);
    if (crmask[q] == 0) {
        crmask[q] = 1;
        list_push(w.listCR, &w.cnt->nCR, w.capCR, (unsigned int)q, info);
    }
}

__global__ void __launch_bounds__(128)
sp_grow1_kernel(int H, int W, SparseWork w, unsigned int stamp, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int total = list_len(&w.cnt->nC0, w.capC) * 9u;          // <= 9 capC: no overflow
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        size_t q;
        bool centre;
        if (!grow_neighbour(w.listC0[t / 9u], (int)(t % 9u), H, W, q, centre)) continue;
        // S1 flags are complete (previous kernel); a member of C0 carries S1 as well
        if (!flag_has(w.flags, q, stamp, FLAG_S1)) continue;
        if (flag_set(w.flags, q, stamp, FLAG_C1)) list_push(w.listC1, &w.cnt->nC1, w.capC, (unsigned int)q, info);
    }
}

__global__ void __launch_bounds__(128)
sp_grow2a_kernel(const uint8_t *__restrict__ inmask, uint8_t *__restrict__ crmask, int H, int W, LacParams prm,
                 SparseWork w, unsigned int stamp, int iter, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int total = list_len(&w.cnt->nC1, w.capC) * 9u;
    // a member of C1 passed good & s' > sigclip; that implies this step's test if sigcliplow is not higher
    const bool implied = prm.sigcliplow <= prm.sigclip;
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        size_t q;
        bool centre;
        if (!grow_neighbour(w.listC1[t / 9u], (int)(t % 9u), H, W, q, centre)) continue;
        if (inmask && inmask[q]) continue;
        if (implied && (centre || flag_has(w.flags, q, stamp, FLAG_C1))) {      // C1 flags are complete
            if (flag_set(w.flags, q, stamp, FLAG_C2)) grow_accept(crmask, q, w, iter, info);
        } else if (flag_set(w.flags, q, stamp, FLAG_T2)) {
            list_push(w.listB, &w.cnt->nQ, w.capB, (unsigned int)q, info);      // list Q
        }
    }
}

__global__ void __launch_bounds__(128)
sp_grow2b_kernel(const float *__restrict__ img, uint8_t *__restrict__ crmask, int H, int W, LacParams prm,
                 SparseWork w, unsigned int stamp, int iter, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int n = list_len(&w.cnt->nQ, w.capB);
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float rn2 = lac_rn2(prm);
    for (unsigned int k = warp; k < n; k += nwarps) {
        const unsigned int q = w.listB[k];
        const int qy = (int)(q / (unsigned int)W), qx = (int)(q - (unsigned int)qy * (unsigned int)W);
        float s_c, nz_c;
        const float sp = warp_sprime(img, H, W, qy, qx, rn2, lane, s_c, nz_c);
        if (sp > prm.sigcliplow && lane == 0) {
            if (flag_set(w.flags, q, stamp, FLAG_C2)) grow_accept(crmask, q, w, iter, info);
        }
    }
}

__global__ void sp_init_kernel(long long *info, int n, SparseCounters *cnt, unsigned int bg_valid)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) info[i] = (i == INFO_ACTIVE) ? 1 : 0;
    if (threadIdx.x == 0) {
        cnt->nA[0] = cnt->nA[1] = cnt->nB = cnt->nC0 = cnt->nC1 = cnt->nCR = 0;
        cnt->bg_valid = bg_valid; cnt->nQ = 0;
        cnt->clean_lo = cnt->clean_hi = 0;
    }
}

// The flag bytes of this iteration go back to zero, list by list (every flag sits on a pixel of
// list A -- which holds B and C0 --, of list C1 or of list Q), so that the 2-bit iteration stamp can
// wrap without the 111 MB memset it used to cost every third iteration.  Runs before sp_control_kernel
// resets the counters.
__global__ void __launch_bounds__(256)
sp_flags_cleanup_kernel(SparseWork w, int it, const long long *__restrict__ info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int nA = list_len(&w.cnt->nA[it & 1], w.capA), nC1 = list_len(&w.cnt->nC1, w.capC);
    const unsigned int nQ = list_len(&w.cnt->nQ, w.capB);
    const unsigned int total = nA + nC1 + nQ;
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned int p = t < nA ? w.listA[it & 1][t] : t < nA + nC1 ? w.listC1[t - nA] : w.listB[t - nA - nC1];
        w.flags[p] = 0;
    }
}

// after grow2: stop flag for the iterations that follow, reset of the per-iteration lists
__global__ void sp_control_kernel(long long *info, int iter, SparseCounters *cnt)
{
    if (!info[INFO_ACTIVE]) return;
    info[INFO_ITERS] = iter + 1;
    if (info[INFO_NCR + iter] == 0) info[INFO_ACTIVE] = 0;
    cnt->nA[(iter + 1) & 1] = 0;           // destination of the next iteration's rescan
    cnt->clean_lo = cnt->clean_hi;         // CR-list entries [clean_lo, clean_hi) are new in this iteration
    cnt->clean_hi = cnt->nCR < 0xffffffffu ? cnt->nCR : 0xffffffffu;
    cnt->nB = cnt->nC0 = cnt->nC1 = cnt->nQ = 0;
}

// medmask cleaning: a flagged pixel becomes the lower median of the unflagged, unmasked pixels of
// its 5x5 box (read from the image: unflagged pixels are never rewritten, so the result only
// depends on WHICH neighbours are flagged).  ALL: every pixel of the cumulative CR list (first
// iteration: all of them are new).  !ALL: only flagged pixels with a newly flagged pixel in
// their box can change -- walk the 5x5 boxes of the new CR-list entries (duplicates recompute
// the same value).
// One warp per pixel: lanes 0..24 fetch the 5x5 box in one round trip, every usable value finds its
// position in the stable sorted order with 25 shuffles, and the lane holding the lower median
// writes it (the first version walked the box and insertion-sorted it in one thread: 42 us for
// the 21 000 cosmic-ray pixels of a frame, 8 % of the warps active).
__device__ __forceinline__ void sp_clean_pixel(float *img, const uint8_t *__restrict__ crmask,
                                               const uint8_t *__restrict__ inmask, int H, int W, int y, int x,
                                               const SparseWork &w, long long *info, int lane)
{
    if (x < 2 || x >= W - 2 || y < 2 || y >= H - 2) return;         // warp-uniform
    const size_t p = (size_t)y * W + x;
    float v = 0.f;
    bool ok = false;
    if (lane < 25) {
        const size_t j = (size_t)(y + lane / 5 - 2) * W + (x + lane % 5 - 2);
        ok = !(crmask[j] || (inmask && inmask[j]));
        if (ok) v = img[j];
    }
    const unsigned int usable = __ballot_sync(0xffffffffu, ok);
    const int m = __popc(usable);
    if (m == 0) {                      // no usable neighbour: the global background level
        if (lane == 0) {
            if (w.cnt->bg_valid) img[p] = *w.background;
            else atomicOr((unsigned long long *)&info[INFO_STATUS], (unsigned long long)LAC_STATUS_NEED_BG);
        }
        return;
    }
    // position in the stable ascending order of the usable values (NaNs last)
    const float mine = (v != v) ? INFINITY : v;
    int rank = 0;
#pragma unroll
    for (int k = 0; k < 25; k++) {
        float vk = __shfl_sync(0xffffffffu, v, k);
        vk = (vk != vk) ? INFINITY : vk;
        const bool okk = (usable >> k) & 1u;
        rank += okk && (vk < mine || (vk == mine && k < lane));
    }
    if (ok && rank == (m - 1) / 2) img[p] = v;
}

template <bool ALL>
__global__ void __launch_bounds__(128)
sp_clean_kernel(float *img, const uint8_t *__restrict__ crmask, const uint8_t *__restrict__ inmask, int H, int W,
                SparseWork w, long long *info)
{
    if (!info[INFO_ACTIVE]) return;
    const unsigned int n = list_len(&w.cnt->nCR, w.capCR);
    const int lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    if (ALL) {
        for (unsigned long long k = warp; k < n; k += nwarps) {
            const unsigned int p = w.listCR[k];
            const int y = (int)(p / (unsigned int)W), x = (int)(p - (unsigned int)y * (unsigned int)W);
            sp_clean_pixel(img, crmask, inmask, H, W, y, x, w, info, lane);
        }
    } else {
        const unsigned int lo = min(w.cnt->clean_lo, n), hi = min(w.cnt->clean_hi, n);
        const unsigned long long total = (unsigned long long)(hi - lo) * 25ull;
        for (unsigned long long t = warp; t < total; t += nwarps) {
            const unsigned int p = w.listCR[lo + (unsigned int)(t / 25ull)];
            const int o = (int)(t % 25ull);
            const int y = (int)(p / (unsigned int)W) + o / 5 - 2, x = (int)(p % (unsigned int)W) + o % 5 - 2;
            if (x < 0 || x >= W || y < 0 || y >= H) continue;
            if (!crmask[(size_t)y * W + x]) continue;                 // crmask is final here: warp-uniform
            sp_clean_pixel(img, crmask, inmask, H, W, y, x, w, info, lane);
        }
    }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static int sparse_begin(const float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, int niter,
                        bool with_background, void *work, long long *info, cudaStream_t st)
{
    const size_t n = (size_t)H * W;
    const SparseWork w = carve_sparse(work, n);
    // crmask and the flag bytes are cleared by the dense scan of iteration 0 (it visits every
    // pixel anyway and has store bandwidth to spare); niter == 0 never scans
    if (niter <= 0) BBX_CUDA(cudaMemsetAsync(crmask, 0, n, st));
    if (with_background) {
        if (bbx_masked_lower_median(img, inmask, n, w.sel, w.background, st)) return -2;
    } else {
        BBX_CUDA(cudaMemsetAsync(w.bghist, 0, 4ull * BG_BINS, st));
        BBX_CUDA(cudaFuncSetAttribute(sp_bg_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(unsigned int) * BG_SAMPLES)));
        sp_bg_gather_kernel<<<BG_SAMPLES / 1024, 1024, 0, st>>>(img, inmask, n, w);
        sp_bg_sample_kernel<<<1, 1024, sizeof(unsigned int) * BG_SAMPLES, st>>>(w);
    }
    sp_init_kernel<<<1, 32, 0, st>>>(info, INFO_NCR + niter, w.cnt, with_background ? 1u : 0u);
    BBX_CHECK_LAUNCH("sparse_begin");
    return 0;
}

static int sparse_iteration(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W, const LacParams &prm,
                            int it, bool with_background, void *work, long long *info, cudaStream_t st,
                            bool prescanned = false, bool stats_taken = false)
{
    const size_t n = (size_t)H * W;
    SparseWork w = carve_sparse(work, n);
    unsigned int stamp = (unsigned int)(it % STAMP_PERIOD) + 1;
    BBX_REQUIRE(ceil_div(H, SCAN_ROWS) <= 65535, "lazy LACosmic: %d rows exceed the scan grid (use the dense mode)", H);
    const dim3 scan_blocks(ceil_div((W + 3) / 4, SCAN_THREADS), ceil_div(H, SCAN_ROWS));
    // iterations after the first work on short lists (what the cleaning step touched): a full
    // grid would mostly launch and retire idle CTAs
    const int list_blocks = it == 0 ? BBX_SM_COUNT * 8 : BBX_SM_COUNT * 2;
    const int warp_blocks = it == 0 ? BBX_SM_COUNT * 16 : BBX_SM_COUNT * 4;
    if (it == 0 && prescanned) {
        // bbx_reduce_apply_scan has made list A, the statistics and the cleared byte maps; the mask
        // morphology has corrected the statistics for the pixels it masked since
        sp_bg_rank_kernel<<<1, 1024, 0, st>>>(w);
    } else if (it == 0 && stats_taken) {
        // bbx_reduce_apply_stats has taken the statistics (and the morphology has corrected them):
        // the scan is the Laplacian alone
        sp_scan_kernel<false><<<scan_blocks, SCAN_THREADS, 0, st>>>(img, inmask, crmask, H, W, prm, w, info);
        sp_bg_rank_kernel<<<1, 1024, 0, st>>>(w);
    } else if (it == 0) {
        if (with_background) {
            sp_scan_kernel<false><<<scan_blocks, SCAN_THREADS, 0, st>>>(img, inmask, crmask, H, W, prm, w, info);
        } else {
            sp_scan_kernel<true><<<scan_blocks, SCAN_THREADS, 0, st>>>(img, inmask, crmask, H, W, prm, w, info);
            sp_bg_rank_kernel<<<1, 1024, 0, st>>>(w);
        }
    } else sp_rescan_kernel<<<list_blocks, 128, 0, st>>>(img, H, W, prm, w, it, stamp, info);
    sp_cand1_kernel<<<list_blocks, 128, 0, st>>>(img, inmask, H, W, prm, w, it, info);
    sp_cand2_kernel<<<warp_blocks, 128, 0, st>>>(img, H, W, prm, w, stamp, info);
    sp_grow1_kernel<<<list_blocks, 128, 0, st>>>(H, W, w, stamp, info);
    sp_grow2a_kernel<<<list_blocks, 128, 0, st>>>(inmask, crmask, H, W, prm, w, stamp, it, info);
    sp_grow2b_kernel<<<warp_blocks, 128, 0, st>>>(img, crmask, H, W, prm, w, stamp, it, info);
    sp_flags_cleanup_kernel<<<list_blocks, 256, 0, st>>>(w, it, info);
    sp_control_kernel<<<1, 1, 0, st>>>(info, it, w.cnt);
    if (it == 0) sp_clean_kernel<true><<<warp_blocks, 128, 0, st>>>(img, crmask, inmask, H, W, w, info);
    else sp_clean_kernel<false><<<warp_blocks, 128, 0, st>>>(img, crmask, inmask, H, W, w, info);
    BBX_CHECK_LAUNCH("sparse_iteration");
    return 0;
}

BgTrack lac_sparse_bg_track(const float *img, void *lac_work, int H, int W)
{
    BgTrack t = {nullptr, nullptr, nullptr};
    if (img && lac_work) {
        const SparseWork w = carve_sparse(lac_work, (size_t)H * W);
        t.img = img; t.bg = w.bg; t.hist = w.bghist;
    }
    return t;
}

static inline bool fuse_aligned(const void *p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; }

// bbx_reduce_apply + the dense scan of LACosmic's first iteration in one pass (see include/bbx.h)
extern "C" int bbx_reduce_apply_scan(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                                     const double *vos_fit, const double *oscan, const float *mbias, const float *mflat,
                                     const uint8_t *bpm, const double *satlevel, const bbx_maskbits *bits,
                                     float *out_img, uint8_t *out_mask, unsigned int *seeds, unsigned int *seed_count,
                                     unsigned int seed_cap, uint8_t *crmask, float sigclip, float sigfrac, float objlim,
                                     float readnoise, const double *readnoise_dev, int niter, void *lac_work,
                                     long long *lac_info, void *stream)
{
    BBX_REQUIRE(g && raw && out_img && out_mask && bits && crmask && lac_work && lac_info,
                "bbx_reduce_apply_scan: null argument");
    BBX_REQUIRE(g->ny == 2 && g->nx * g->ny == BBX_NCHAN, "bbx_reduce_apply_scan: expected 2 x 8 channels");
    BBX_REQUIRE((const void *)out_img != raw, "bbx_reduce_apply_scan: output must not alias the raw frame");
    BBX_REQUIRE((seeds == nullptr) == (seed_count == nullptr), "bbx_reduce_apply_scan: seeds and seed_count go together");
    BBX_REQUIRE(niter > 0 && sigclip >= 0.f && sigfrac >= 0.f, "bbx_reduce_apply_scan: needs niter > 0 and non-negative thresholds");
    const long long RW = (long long)g->nx * g->xsize_chan, RH = (long long)g->ny * g->ysize_chan;
    BBX_REQUIRE(RH * RW < 2147483647LL, "bbx_reduce_apply_scan: frame too large for 31-bit pixel indices");
    const size_t esz = raw_type == BBX_RAW_U16 ? 2 : 4;
    const bool vec4 = (g->xsize_chan % 4 == 0) && (g->dx % 4 == 0) && (g->W % 4 == 0) && fuse_aligned(raw, 4 * esz) &&
                      fuse_aligned(mbias, 16) && fuse_aligned(mflat, 16) && fuse_aligned(out_img, 16) &&
                      fuse_aligned(bpm, 4) && fuse_aligned(out_mask, 4) && fuse_aligned(crmask, 4) &&
                      (RH + FUSE_ROWS - 1) / FUSE_ROWS <= 65535;
    BBX_REQUIRE(vec4, "bbx_reduce_apply_scan: layout not 4-pixel aligned (use bbx_reduce_apply + bbx_lacosmic)");
    cudaStream_t st = (cudaStream_t)stream;
    ChanF32 gn;
    for (int i = 0; i < BBX_NCHAN; i++) gn.v[i] = gain_h ? gain_h[i] : 1.0f;
    ApplyArgs a = {vos_fit, oscan, mbias, mflat, bpm, satlevel, out_img, out_mask, bits->bad, bits->saturated,
                   seeds, seed_count, seed_cap, (unsigned int)(bits->saturated | bits->satcon)};
    const size_t n = (size_t)(RH * RW);
    SparseWork w = carve_sparse(lac_work, n);
    FuseArgs f;
    f.crmask = crmask;
    f.prm = lac_make_params(sigclip, sigfrac, objlim, readnoise, readnoise_dev);
    f.w = w;
    f.info = lac_info;
    if (seeds) BBX_CUDA(cudaMemsetAsync(seed_count, 0, sizeof(unsigned int), st));
    // what bbx_lacosmic_begin does for the lazy path, with the sample taken through the fused pass's
    // own arithmetic on the raw frame
    BBX_CUDA(cudaMemsetAsync(w.bghist, 0, 4ull * BG_BINS, st));
    static bool attr_set = false;
    if (!attr_set) {
        BBX_CUDA(cudaFuncSetAttribute(sp_bg_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(unsigned int) * BG_SAMPLES)));
        attr_set = true;
    }
    if (raw_type == BBX_RAW_U16)
        sp_bg_gather_raw_kernel<uint16_t><<<BG_SAMPLES / 1024, 1024, 0, st>>>((const uint16_t *)raw, *g, gn, a, w);
    else
        sp_bg_gather_raw_kernel<float><<<BG_SAMPLES / 1024, 1024, 0, st>>>((const float *)raw, *g, gn, a, w);
    sp_bg_sample_kernel<<<1, 1024, sizeof(unsigned int) * BG_SAMPLES, st>>>(w);
    sp_init_kernel<<<1, 32, 0, st>>>(lac_info, INFO_NCR + niter, w.cnt, 0u);
    const long long warps_x = (RW / 4 + 29) / 30;             // a warp owns 30 groups of 4 columns
    const dim3 grid((unsigned int)((warps_x + FUSE_THREADS / 32 - 1) / (FUSE_THREADS / 32)), (unsigned int)((RH + FUSE_ROWS - 1) / FUSE_ROWS));
    if (raw_type == BBX_RAW_U16)
        reduce_apply_scan_kernel<uint16_t><<<grid, FUSE_THREADS, 0, st>>>((const uint16_t *)raw, *g, gn, a, f);
    else
        reduce_apply_scan_kernel<float><<<grid, FUSE_THREADS, 0, st>>>((const float *)raw, *g, gn, a, f);
    BBX_CHECK_LAUNCH("bbx_reduce_apply_scan");
    return 0;
}

// bbx_reduce_apply with LACosmic's background statistics taken on the way (see include/bbx.h): the
// middle road between the separate passes and bbx_reduce_apply_scan -- the per-pixel pass keeps its
// shape (HBM-bound, 64 registers), the dense scan keeps the Laplacian but sheds the mask read and the
// key arithmetic.
extern "C" int bbx_reduce_apply_stats(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                                      const double *vos_fit, const double *oscan, const float *mbias, const float *mflat,
                                      const uint8_t *bpm, const double *satlevel, const bbx_maskbits *bits,
                                      float *out_img, uint8_t *out_mask, unsigned int *seeds, unsigned int *seed_count,
                                      unsigned int seed_cap, int niter, void *lac_work, long long *lac_info, void *stream)
{
    BBX_REQUIRE(g && raw && out_img && out_mask && bits && lac_work && lac_info, "bbx_reduce_apply_stats: null argument");
    BBX_REQUIRE(niter != 0, "bbx_reduce_apply_stats: niter %d", niter);
    const long long RW = (long long)g->nx * g->xsize_chan, RH = (long long)g->ny * g->ysize_chan;
    BBX_REQUIRE(RH * RW < 2147483647LL, "bbx_reduce_apply_stats: frame too large for 31-bit pixel indices");
    cudaStream_t st = (cudaStream_t)stream;
    ChanF32 gn;
    for (int i = 0; i < BBX_NCHAN; i++) gn.v[i] = gain_h ? gain_h[i] : 1.0f;
    ApplyArgs a = {vos_fit, oscan, mbias, mflat, bpm, satlevel, out_img, out_mask, bits->bad, bits->saturated,
                   seeds, seed_count, seed_cap, (unsigned int)(bits->saturated | bits->satcon)};
    SparseWork w = carve_sparse(lac_work, (size_t)(RH * RW));
    if (niter < 0)              // the per-pixel kernel alone, with the bracket of an earlier call (for timing it)
        return apply_launch(raw, raw_type, g, gain_h, vos_fit, oscan, mbias, mflat, bpm, satlevel, bits, out_img, out_mask,
                            seeds, seed_count, seed_cap, w.bg, w.bghist, stream);
    BBX_CUDA(cudaMemsetAsync(w.bghist, 0, 4ull * BG_BINS, st));
    static bool attr_set = false;
    if (!attr_set) {
        BBX_CUDA(cudaFuncSetAttribute(sp_bg_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(unsigned int) * BG_SAMPLES)));
        attr_set = true;
    }
    if (raw_type == BBX_RAW_U16)
        sp_bg_gather_raw_kernel<uint16_t><<<BG_SAMPLES / 1024, 1024, 0, st>>>((const uint16_t *)raw, *g, gn, a, w);
    else
        sp_bg_gather_raw_kernel<float><<<BG_SAMPLES / 1024, 1024, 0, st>>>((const float *)raw, *g, gn, a, w);
    sp_bg_sample_kernel<<<1, 1024, sizeof(unsigned int) * BG_SAMPLES, st>>>(w);
    sp_init_kernel<<<1, 32, 0, st>>>(lac_info, INFO_NCR + niter, w.cnt, 0u);
    BBX_CHECK_LAUNCH("bbx_reduce_apply_stats");
    return apply_launch(raw, raw_type, g, gain_h, vos_fit, oscan, mbias, mflat, bpm, satlevel, bits, out_img, out_mask,
                        seeds, seed_count, seed_cap, w.bg, w.bghist, stream);
}

extern "C" size_t bbx_lacosmic_work_bytes(int H, int W)
{
    const size_t a = lac_dense_work_bytes(H, W), b = lac_sparse_work_bytes(H, W);
    return a > b ? a : b;
}

extern "C" int bbx_lacosmic_begin(const float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                                  int niter, int mode, void *work, long long *out_info, void *stream)
{
    BBX_REQUIRE(img && crmask && work && out_info, "bbx_lacosmic_begin: null argument");
    BBX_REQUIRE(H > 0 && W > 0 && niter >= 0, "bbx_lacosmic_begin: bad shape %d x %d or niter %d", H, W, niter);
    BBX_REQUIRE(mode >= 0 && mode <= 4, "bbx_lacosmic_begin: mode %d (0 = lazy, 1 = dense, 2 = lazy with background level, 3 = lazy, scanned by bbx_reduce_apply_scan, 4 = lazy, statistics by bbx_reduce_apply_stats)", mode);
    BBX_REQUIRE((long long)H * W < 4294967295LL, "bbx_lacosmic_begin: image too large for 32-bit pixel indices");
    if (mode == 3 || mode == 4) return 0;        // bbx_reduce_apply_scan / _stats have done it
    if (mode == 1) return lac_dense_begin(img, inmask, crmask, H, W, niter, work, out_info, (cudaStream_t)stream);
    return sparse_begin(img, inmask, crmask, H, W, niter, mode == 2, work, out_info, (cudaStream_t)stream);
}

extern "C" int bbx_lacosmic_iteration(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                                      float sigclip, float sigfrac, float objlim, float readnoise,
                                      const double *readnoise_dev, int iter, int mode, void *work,
                                      long long *out_info, void *stream)
{
    BBX_REQUIRE(img && crmask && work && out_info, "bbx_lacosmic_iteration: null argument");
    const LacParams prm = lac_make_params(sigclip, sigfrac, objlim, readnoise, readnoise_dev);
    if (mode == 1) return lac_dense_iteration(img, inmask, crmask, H, W, prm, iter, work, out_info, (cudaStream_t)stream);
    return sparse_iteration(img, inmask, crmask, H, W, prm, iter, mode == 2, work, out_info, (cudaStream_t)stream,
                            mode == 3, mode == 4);
}

extern "C" int bbx_lacosmic(float *img, const uint8_t *inmask, uint8_t *crmask, int H, int W,
                            float sigclip, float sigfrac, float objlim, float readnoise,
                            const double *readnoise_dev, int niter, int mode, void *work, long long *out_info,
                            void *stream)
{
    if (bbx_lacosmic_begin(img, inmask, crmask, H, W, niter, mode, work, out_info, stream)) return -2;
    for (int it = 0; it < niter; it++)
        if (bbx_lacosmic_iteration(img, inmask, crmask, H, W, sigclip, sigfrac, objlim, readnoise, readnoise_dev,
                                   it, mode, work, out_info, stream)) return -2;
    return 0;
}

// --------------------------------------------------------------------------------------------
// after the iterations: cosmic-ray bit into the frame mask and NCOSMICS (8-connected components
// of crmask; ndimage.label, blackbox.py:4349-4355) from the CR list instead of dense passes
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int sp_uf_find(const int *L, int i)
{
    int p = L[i];
    while (p != i) { i = p; p = L[i]; }
    return i;
}
__device__ __forceinline__ void sp_uf_union(int *L, int a, int b)
{
    bool done;
    do {
        a = sp_uf_find(L, a);
        b = sp_uf_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__global__ void __launch_bounds__(128)
sp_finish_mark_kernel(uint8_t *mask, unsigned int bit, SparseWork w, int *__restrict__ L, int32_t *out_n)
{
    const unsigned int n = list_len(&w.cnt->nCR, w.capCR);
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_n = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int p = w.listCR[k];
        if (mask) {
            unsigned int *word = reinterpret_cast<unsigned int *>(mask + (p & ~3u));
            atomicOr(word, bit << ((p & 3u) * 8));
        }
        L[p] = (int)p;
    }
}

__global__ void __launch_bounds__(128)
sp_finish_merge_kernel(const uint8_t *__restrict__ crmask, int H, int W, SparseWork w, int *__restrict__ L)
{
    if (w.cnt->nCR > w.capCR) return;
    const unsigned int n = w.cnt->nCR;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int p = (int)w.listCR[k];
        const int y = p / W, x = p - y * W;
        if (x > 0 && crmask[p - 1]) sp_uf_union(L, p, p - 1);
        if (y > 0) {
            const int q = p - W;
            if (crmask[q]) sp_uf_union(L, p, q);
            if (x > 0 && crmask[q - 1]) sp_uf_union(L, p, q - 1);
            if (x + 1 < W && crmask[q + 1]) sp_uf_union(L, p, q + 1);
        }
    }
}

__global__ void __launch_bounds__(128)
sp_finish_count_kernel(SparseWork w, const int *__restrict__ L, int32_t *out_n)
{
    const unsigned int n = list_len(&w.cnt->nCR, w.capCR);
    int c = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int p = w.listCR[k];
        if (L[p] == (int)p) c++;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out_n, c);
}

extern "C" int bbx_lacosmic_finish(const uint8_t *crmask, uint8_t *mask, int cosmic_bit, int H, int W, int mode,
                                   void *work, int32_t *labels, int32_t *out_ncosmics, void *stream)
{
    BBX_REQUIRE(crmask && work && labels && out_ncosmics, "bbx_lacosmic_finish: null argument");
    BBX_REQUIRE(mask == nullptr || ((uintptr_t)mask & 3) == 0, "bbx_lacosmic_finish: mask must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        if (mask && bbx_mask_or(mask, crmask, (size_t)H * W, cosmic_bit, stream)) return -2;
        return bbx_count_objects(crmask, 0xff, H, W, labels, out_ncosmics, stream);
    }
    const SparseWork w = carve_sparse(work, (size_t)H * W);
    const int lb = BBX_SM_COUNT * 4;
    sp_finish_mark_kernel<<<lb, 128, 0, st>>>(mask, (unsigned int)cosmic_bit, w, labels, out_ncosmics);
    sp_finish_merge_kernel<<<lb, 128, 0, st>>>(crmask, H, W, w, labels);
    sp_finish_count_kernel<<<lb, 128, 0, st>>>(w, labels, out_ncosmics);
    BBX_CHECK_LAUNCH("bbx_lacosmic_finish");
    return 0;
}
