// apply_common.cuh -- pieces of the fused per-pixel pass shared by apply.cu (the pass on its own)
// and lacosmic_sparse.cu (the pass fused with LACosmic's dense Laplacian scan).
#pragma once
#include "bbx_common.cuh"

struct ApplyArgs {
    const double *vos_fit;     // [16][dy] or null
    const double *oscan;       // [16][xsize_chan] or null
    const float *mbias;        // [red] or null
    const float *mflat;        // [red] or null
    const uint8_t *bpm;        // [red] or null
    const double *satlevel;    // [16] device, or null
    float *out_img;            // [red]
    uint8_t *out_mask;         // [red] or null
    int bit_bad, bit_sat;
    unsigned int *seeds;       // optional list of pixels that seed the mask morphology
    unsigned int *seed_count;  // [0] entries appended (may exceed seed_cap: overflow)
    unsigned int seed_cap;
    unsigned int seed_bits;    // saturated | saturated-connected bit values
};

template <typename T> struct RawVec4;
// u16 -> f32 through the exponent trick (2^23 + n) - 2^23: exact, and on the ALU / FMA pipes
// instead of the quarter-rate conversion unit
template <> struct RawVec4<uint16_t> {
    static __device__ __forceinline__ void load(const uint16_t *p, float v[4]) {
        const uint2 u = __ldcs(reinterpret_cast<const uint2 *>(p));
        v[0] = __uint_as_float(0x4b000000u | (u.x & 0xffffu)) - 8388608.0f;
        v[1] = __uint_as_float(0x4b000000u | (u.x >> 16)) - 8388608.0f;
        v[2] = __uint_as_float(0x4b000000u | (u.y & 0xffffu)) - 8388608.0f;
        v[3] = __uint_as_float(0x4b000000u | (u.y >> 16)) - 8388608.0f;
    }
};
template <> struct RawVec4<float> {
    static __device__ __forceinline__ void load(const float *p, float v[4]) {
        const uint4 u = __ldcs(reinterpret_cast<const uint4 *>(p));
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y);
        v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
};

// smallest float32 t with (double)t >= level: for float32 v, (double)v >= level <=> v >= t
__device__ __forceinline__ float f32_ceil_of(double level)
{
    float t = (float)level;                       // round to nearest
    if ((double)t < level) t = nextafterf(t, INFINITY);
    return t;
}


// The reduced value and the seed-mask byte of ONE pixel (y, x) of the reduced frame, by the same
// sequence of separately rounded operations as the vectorised kernels (see apply.cu's header):
// used where a thread needs a pixel that is not among its own four columns -- the neighbour
// across a warp / CTA boundary in the fused Laplacian scan, the sample of the background level.
template <typename T>
__device__ __forceinline__ float apply_value_at(const T *__restrict__ raw, const bbx_geom &g, const ChanF32 &gain,
                                                const ApplyArgs &a, int y, int x, uint32_t &mbyte)
{
    const int RW = g.nx * g.xsize_chan;
    const int r = (y >= g.ysize_chan) ? 1 : 0;                  // ny == 2
    const int c = x / g.xsize_chan, lx = x - c * g.xsize_chan;
    const int ch = r * g.nx + c;
    const int rr = (r == 0 ? g.data_y0_bot : g.data_y0_top) + (y - r * g.ysize_chan);
    const size_t ro = (size_t)rr * g.W + (size_t)c * g.dx + lx;
    const size_t oo = (size_t)y * RW + x;
    float w = raw_to_f32<T>(raw[ro]) * gain.v[ch];
    if (a.vos_fit) w = sub_f64(w, a.vos_fit[(size_t)ch * g.dy + (rr - r * g.dy)]);
    else w = sub_f64(w, 0.0);
    w = sub_f64(w, a.oscan ? a.oscan[(size_t)ch * g.xsize_chan + lx] : 0.0);
    if (a.mbias) w = w - a.mbias[oo];
    uint32_t m = a.bpm ? (uint32_t)a.bpm[oo] : 0u;
    if (a.out_mask != nullptr) {
        if (!isfinite(w)) { w = 0.f; if (m == 0) m |= (uint32_t)a.bit_bad; }
        if (a.satlevel != nullptr) {
            const double lv = a.satlevel[ch];
            if (lv == lv && w >= f32_ceil_of(lv)) m |= (uint32_t)(a.bit_sat | BBX_TMP_SAT);
        }
    }
    if (a.mflat) w = w / a.mflat[oo];
    mbyte = m;
    return w;
}

// the per-pixel pass; bg_state / bghist: optional background statistics of LACosmic (apply.cu)
int apply_launch(const void *raw, int raw_type, const bbx_geom *g, const float *gain_h,
                 const double *vos_fit, const double *oscan, const float *mbias, const float *mflat,
                 const uint8_t *bpm, const double *satlevel, const bbx_maskbits *bits,
                 float *out_img, uint8_t *out_mask, unsigned int *seeds, unsigned int *seed_count,
                 unsigned int seed_cap, void *bg_state, unsigned int *bghist, void *stream);
