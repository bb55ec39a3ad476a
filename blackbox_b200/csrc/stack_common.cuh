// stack_common.cuh -- argument block shared by the master-combine kernels
#pragma once

#define STACK_MAX 64

struct StackArgs {
    const float *frames[STACK_MAX];     // device pointers, kernel parameter space
    float scale[STACK_MAX];             // divisor per frame (0 = none)
    float rcp[STACK_MAX];               // RN(1 / scale), correctly rounded (host), for stack_div_by
};

// x / d, correctly rounded, for a divisor that is the same for a whole frame.  The compiler's IEEE
// division is a reciprocal approximation, a Newton step on it, and two residual corrections of the
// quotient -- about a dozen instructions per value, twenty values per pixel in a master flat.  With
// r = RN(1/d) known per frame only the two residual corrections are left (Markstein: q' = RN(q + r
// (x - d q)) is the correctly rounded quotient when r is the correctly rounded reciprocal, q is
// within an ulp and nothing over- or underflows; d with an all-ones significand is excluded by the
// host).  Values whose exponent is outside [2^-60, 2^60] (zeros, denormals, inf, NaN included) take
// the division itself.  Checked exhaustively over all 2^32 bit patterns of x for a set of divisors
// (bbx_debug_div_check, tests/test_steps_gpu.py).
__device__ __forceinline__ bool stack_div_in_range(float x)
{
    const unsigned int ex = (__float_as_uint(x) >> 23) & 0xffu;
    return ex - 67u <= 120u;
}
__device__ __forceinline__ float stack_div_fast(float x, float d, float r)     // x in range (stack_div_in_range)
{
    float q = __fmul_rn(x, r);
    float e = __fmaf_rn(-d, q, x);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-d, q, x);
    return __fmaf_rn(e, r, q);
}
__device__ __forceinline__ float stack_div_by(float x, float d, float r)
{
    return stack_div_in_range(x) ? stack_div_fast(x, d, r) : x / d;
}

// host: RN(1/d) for a normal positive float d, and whether d qualifies for stack_div_by
static inline bool stack_rcp(float d, float *r_out)
{
    union { float f; unsigned int u; } v;
    v.f = d;
    const unsigned int ex = (v.u >> 23) & 0xffu, man = v.u & 0x7fffffu;
    if ((v.u >> 31) || ex < 97u || ex > 157u || man == 0x7fffffu) return false;       // d in [2^-30, 2^30], not 1.11..1
    // 1/d in double, rounded to float, then settled exactly: r * d is exact in double (48 bits)
    float r = (float)(1.0 / (double)d);
    float best = r;
    double berr = 1.0 - (double)r * (double)d;
    if (berr < 0) berr = -berr;
    for (int s = -1; s <= 1; s += 2) {
        union { float f; unsigned int u; } c;
        c.f = r;
        c.u += s;
        double err = 1.0 - (double)c.f * (double)d;
        if (err < 0) err = -err;
        if (err < berr) { berr = err; best = c.f; }
    }
    *r_out = best;
    return true;
}
