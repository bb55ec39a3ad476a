// bbx_common.cuh -- shared helpers for the sm_100a kernels of libbbx.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/bbx.h"

// SM count of the current device (queried once per device, api.cu); grids are sized from it
int bbx_sm_count(void);
#define BBX_SM_COUNT bbx_sm_count()

void bbx_set_error(const char *fmt, ...);

#define BBX_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            bbx_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));     \
            return -2;                                                                \
        }                                                                             \
    } while (0)

#define BBX_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            bbx_set_error("%s: %s", #call, cudaGetErrorString(_e));                   \
            return -2;                                                                \
        }                                                                             \
    } while (0)

#define BBX_REQUIRE(cond, ...)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            bbx_set_error(__VA_ARGS__);                                               \
            return -1;                                                                \
        }                                                                             \
    } while (0)

// per-channel scalars passed by value (kernel parameter space, no device allocation)
struct ChanF32 { float v[BBX_NCHAN]; };
struct ChanF64 { double v[BBX_NCHAN]; };

// ---- raw pixel access: u16 counts or f32 (already converted) ---------------------------
template <typename T> __device__ __forceinline__ float raw_to_f32(T v);
template <> __device__ __forceinline__ float raw_to_f32<uint16_t>(uint16_t v) { return (float)v; }
template <> __device__ __forceinline__ float raw_to_f32<float>(float v) { return v; }

// f32 - f64 -> f32, the way numpy evaluates `f32_array -= f64_array`
__device__ __forceinline__ float sub_f64(float a, double b) { return (float)((double)a - b); }

// ---- warp / block reductions (fixed order => deterministic) ----------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result broadcast to every thread; scratch: >= 33 elements of T in smem.
// blockDim.x must be a multiple of 32 and <= 1024.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // scratch may still be read from a previous call
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        T t = (lane < nw) ? scratch[lane] : (T)0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---- 128-bit / 64-bit streaming loads and stores --------------------------------------
__device__ __forceinline__ uint2 ld_stream_u2(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v)
{
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
