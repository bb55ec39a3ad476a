// fitscodec.cu -- FITS data-unit byte order on the device: the host moves the
// bytes of the file, the GPU turns them into the arrays the kernels work on and back.
//   decode: BITPIX 16, BZERO 32768, BSCALE 1 (raw MeerLICHT / BlackGEM frames: unsigned counts,
//           what read_hdulist hands to blackbox_reduce, blackbox.py:1451) -> uint16, in place or
//           out of place:  value = bswap16(stored) ^ 0x8000;
//           BITPIX -32 -> float32 (master frames, reduced images).
//   encode: float32 -> big-endian float32 (the _red.fits image, blackbox.py:1987), 8-bit data are
//           byte-order free.
#include "bbx_common.cuh"

__global__ void __launch_bounds__(256)
fits_swap16_kernel(const uint4 *in, uint4 *out, size_t n16, unsigned int flip)
{
    // 8 pixels per thread: swap the bytes of every 16-bit lane, then flip the sign bit (BZERO)
    const size_t n = n16 / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        v.x = __byte_perm(v.x, 0, 0x2301) ^ flip; v.y = __byte_perm(v.y, 0, 0x2301) ^ flip;
        v.z = __byte_perm(v.z, 0, 0x2301) ^ flip; v.w = __byte_perm(v.w, 0, 0x2301) ^ flip;
        out[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint16_t *i16 = reinterpret_cast<const uint16_t *>(in);
        uint16_t *o16 = reinterpret_cast<uint16_t *>(out);
        for (size_t k = n * 8; k < n16; k++) {
            const uint16_t x = i16[k];
            o16[k] = (uint16_t)(((x >> 8) | (x << 8)) ^ (flip & 0xffffu));
        }
    }
}

__global__ void __launch_bounds__(256)
fits_swap32_kernel(const uint4 *in, uint4 *out, size_t n32)
{
    const size_t n = n32 / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        v.x = __byte_perm(v.x, 0, 0x0123); v.y = __byte_perm(v.y, 0, 0x0123);
        v.z = __byte_perm(v.z, 0, 0x0123); v.w = __byte_perm(v.w, 0, 0x0123);
        out[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t *i32 = reinterpret_cast<const uint32_t *>(in);
        uint32_t *o32 = reinterpret_cast<uint32_t *>(out);
        for (size_t k = n * 4; k < n32; k++) o32[k] = __byte_perm(i32[k], 0, 0x0123);
    }
}

static int fits_swap(const void *in, void *out, int bitpix, size_t n, unsigned int flip, cudaStream_t st,
                     const char *who)
{
    BBX_REQUIRE(in && out, "%s: null argument", who);
    BBX_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "%s: buffers must be 16-byte aligned", who);
    if (n == 0) return 0;
    const int blocks = BBX_SM_COUNT * 16;
    if (bitpix == 16) fits_swap16_kernel<<<blocks, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n, flip);
    else fits_swap32_kernel<<<blocks, 256, 0, st>>>((const uint4 *)in, (uint4 *)out, n);
    BBX_CHECK_LAUNCH(who);
    return 0;
}

extern "C" int bbx_fits_decode(const void *be, int bitpix, int unsigned16, size_t n, void *out, void *stream)
{
    BBX_REQUIRE(bitpix == 16 || bitpix == -32, "bbx_fits_decode: BITPIX %d (16 and -32 are supported)", bitpix);
    return fits_swap(be, out, bitpix, n, (bitpix == 16 && unsigned16) ? 0x80008000u : 0u, (cudaStream_t)stream,
                     "bbx_fits_decode");
}

extern "C" int bbx_fits_encode(const void *in, int bitpix, int unsigned16, size_t n, void *out_be, void *stream)
{
    BBX_REQUIRE(bitpix == 16 || bitpix == -32, "bbx_fits_encode: BITPIX %d (16 and -32 are supported)", bitpix);
    // the sign-bit flip commutes with the byte swap up to its position: flip the (native) high byte
    // first, i.e. after the swap the low byte of every 16-bit lane
    return fits_swap(in, out_be, bitpix, n, (bitpix == 16 && unsigned16) ? 0x00800080u : 0u, (cudaStream_t)stream,
                     "bbx_fits_encode");
}
