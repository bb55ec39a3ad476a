// bg_track.cuh -- the running statistics of LACosmic's background level (lacosmic_sparse.cu) and
// the hook the mask morphology (mask.cu) uses to keep them right when the dense scan has been
// fused into the per-pixel pass and therefore ran BEFORE the mask was final.
#pragma once
#include "bbx_common.cuh"

// order-preserving 32-bit keys of float32 values
__device__ __forceinline__ unsigned int f32_key(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- background statistics of the lazy path (lacosmic_sparse.cu), shared with mask.cu ----------
struct BgState {
    unsigned int key_a, width;           // bracket: keys key_a .. key_a + width - 1 (width 0: no bracket)
    unsigned long long n_valid, n_below; // unmasked pixels; unmasked pixels with key < key_a
};

// For a pixel whose mask byte the mask morphology has just turned from zero to non-zero: the fused
// pass (reduce_apply_scan_kernel) counted it as unmasked, detect_cosmics' inmask says it is not.
struct BgTrack {
    const float *img;            // null: nothing to track (the scan runs after the morphology)
    BgState *bg;
    unsigned int *hist;
};
__device__ __forceinline__ void bg_untrack(const BgTrack &t, size_t p)
{
    if (t.img == nullptr) return;
    const unsigned int key = f32_key(t.img[p]);
    atomicAdd(&t.bg->n_valid, ~0ull);                           // - 1
    if (key < t.bg->key_a) atomicAdd(&t.bg->n_below, ~0ull);
    else if (key - t.bg->key_a < t.bg->width) atomicSub(&t.hist[key - t.bg->key_a], 1u);
}
BgTrack lac_sparse_bg_track(const float *img, void *lac_work, int H, int W);

