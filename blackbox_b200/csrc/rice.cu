// rice.cu -- tile-compressed FITS images (".fits.fz", ZCMPTYPE = 'RICE_1') decoded on the device
// (the raw frames the reference reads with read_hdulist, blackbox.py:1451, are
// fpacked; astropy / CFITSIO unpack them on the host there).  The host parses the binary-table
// header and hands over the heap plus one (offset, length) descriptor per tile
// (blackbox_b200/fitsio.py:read_compressed); the compressed bytes -- about half the size of the
// frame -- are what crosses PCIe.
//
// Format (FITS tiled-image convention, Rice algorithm as published with CFITSIO and in
// White & Becker / Pence et al. 2009; 16-bit pixels, BYTEPIX = 2, BLOCKSIZE = 32):
//   tile   = first pixel as a big-endian 16-bit value, then blocks of 32 pixel DIFFERENCES
//   block  = 4-bit code FS+1, then per pixel
//              code 0        : all differences are 0 (no further bits)
//              code 15       : the difference as 16 raw bits
//              otherwise     : (diff >> FS) zeros, a one, then the low FS bits
//   diff   = zig-zag mapped (even = +d/2, odd = ~(d >> 1)) difference to the previous pixel,
//            modulo 2^16; bits are packed MSB first.
// The bit position of a block is only known once the block before it is decoded, so a tile is a
// serial job: one thread per tile (a frame has 10600 of them), and every step of it depends on
// the one before, so what limits the kernel is instruction latency, not bandwidth.  Hence only
// RICE_TILES = 4 lanes of a warp decode (a 32-pixel block each, into shared memory): 2650 warps
// instead of 332 give every scheduler several warps to switch between, and fewer lanes means
// fewer divergent paths per warp (measured on B200, full frame: 32 tiles per warp 3.7 ms).  The
// whole warp then stores the staged row segments, 64 contiguous bytes per row -- whole sectors,
// although every decoding thread works on its own row.  Reads are byte loads through the
// read-only path (each 32-byte sector serves ~25 pixels).
#include "bbx_common.cuh"

#define RICE_WARPS 4
#define RICE_TILES 4             // tiles (= decoding lanes) per warp
#define RICE_BLOCK 32
#define RICE_ROW   34            // uint16 per staging row: 17 words, odd, so lanes spread over the banks

struct RiceReader {
    const uint8_t *c, *end;
    unsigned int b;
    int nbits;
    bool overrun;
    __device__ __forceinline__ unsigned int next()
    {
        if (c < end) return __ldg(c++);
        overrun = true;
        return 0xffu;            // a one bit ends every unary run: no endless loop on a truncated tile
    }
};

template <bool FLIP>
__global__ void __launch_bounds__(RICE_WARPS * 32)
rice16_decode_kernel(const uint8_t *__restrict__ heap, size_t heap_bytes, const long long *__restrict__ offs,
                     const int *__restrict__ lens, int ntiles, int nx, uint16_t *__restrict__ out,
                     int *__restrict__ status)
{
    __shared__ uint16_t stage[RICE_WARPS][RICE_TILES][RICE_ROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile0 = (blockIdx.x * RICE_WARPS + warp) * RICE_TILES;
    if (tile0 >= ntiles) return;
    const int tile = tile0 + lane;
    const bool live = lane < RICE_TILES && tile < ntiles;

    RiceReader r;
    r.c = r.end = heap;
    r.b = 0; r.nbits = 8; r.overrun = false;
    bool bad = false;
    unsigned int lastpix = 0;
    if (live) {
        const long long o = offs[tile];
        const int n = lens[tile];
        if (o < 0 || n < 3 || (unsigned long long)o + (unsigned long long)n > heap_bytes) {
            bad = true;
        } else {
            r.c = heap + o; r.end = r.c + n;
            lastpix = (r.next() << 8) | r.next();
            r.b = r.next();
        }
    }
    const int fsbits = 4, fsmax = 14, bbits = 16;
    for (int i = 0; i < nx; i += RICE_BLOCK) {
        const int nthis = min(RICE_BLOCK, nx - i);
        uint16_t *row = stage[warp][lane & (RICE_TILES - 1)];
        if (live && !bad) {
            r.nbits -= fsbits;
            while (r.nbits < 0) { r.b = (r.b << 8) | r.next(); r.nbits += 8; }
            const int fs = (int)(r.b >> r.nbits) - 1;
            r.b &= (1u << r.nbits) - 1u;
            if (fs < 0) {
                for (int k = 0; k < nthis; k++) row[k] = (uint16_t)lastpix;
            } else if (fs == fsmax) {
                for (int k = 0; k < nthis; k++) {
                    int s = bbits - r.nbits;
                    unsigned int diff = r.b << s;
                    for (s -= 8; s >= 0; s -= 8) { r.b = r.next(); diff |= r.b << s; }
                    if (r.nbits > 0) { r.b = r.next(); diff |= r.b >> (-s); r.b &= (1u << r.nbits) - 1u; }
                    else r.b = 0;
                    diff &= 0xffffu;
                    diff = (diff & 1u) ? ~(diff >> 1) : (diff >> 1);
                    lastpix = (lastpix + diff) & 0xffffu;
                    row[k] = (uint16_t)lastpix;
                }
            } else {
                for (int k = 0; k < nthis; k++) {
                    unsigned int nzero = 0;
                    while (r.b == 0) {
                        nzero += r.nbits;            // the rest of the window is zeros
                        r.nbits = 8; r.b = r.next();
                        if (r.overrun) break;
                    }
                    const int top = 32 - __clz(r.b);            // position of the leading one, 1-based
                    nzero += r.nbits - top;
                    r.nbits = top - 1;
                    r.b ^= 1u << r.nbits;
                    r.nbits -= fs;
                    while (r.nbits < 0) { r.b = (r.b << 8) | r.next(); r.nbits += 8; }
                    unsigned int diff = (nzero << fs) | (r.b >> r.nbits);
                    r.b &= (1u << r.nbits) - 1u;
                    diff = (diff & 1u) ? ~(diff >> 1) : (diff >> 1);
                    lastpix = (lastpix + diff) & 0xffffu;
                    row[k] = (uint16_t)lastpix;
                }
            }
        }
        __syncwarp();
        if (lane < nthis) {
            const int rows = min(RICE_TILES, ntiles - tile0);
            for (int t = 0; t < rows; t++) {
                uint16_t v = stage[warp][t][lane];
                if (FLIP) v ^= 0x8000u;                         // BZERO = 32768: stored int16 -> counts
                out[(size_t)(tile0 + t) * nx + i + lane] = v;
            }
        }
        __syncwarp();
    }
    if (live && (bad || r.overrun)) atomicOr(status, bad ? 2 : 1);
}

// ---------------------------------------------------------------------------------------------
// bbx_rice_decode16: heap (device, heap_bytes), offs / lens (device, one per tile: byte offset into
// the heap and compressed length), ntiles tiles of nx pixels each (row tiles: ZTILE1 = ZNAXIS1,
// ZTILE2 = 1), blocksize 32, BYTEPIX 2.  unsigned16 != 0: the stored values are int16 with
// BZERO 32768 and come out as uint16 counts.  status (device int, zeroed by the call): bit 0 = a
// tile ran past its compressed bytes, bit 1 = a descriptor points outside the heap; the affected
// rows are undefined.  The caller reads it after synchronising.
// ---------------------------------------------------------------------------------------------
extern "C" int bbx_rice_decode16(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                                 int ntiles, int nx, int blocksize, int unsigned16, void *out, int *status,
                                 void *stream)
{
    BBX_REQUIRE(heap && offs && lens && out && status, "bbx_rice_decode16: null argument");
    BBX_REQUIRE(ntiles > 0 && nx > 0, "bbx_rice_decode16: %d tiles of %d pixels", ntiles, nx);
    BBX_REQUIRE(blocksize == RICE_BLOCK, "bbx_rice_decode16: BLOCKSIZE %d (32 is supported)", blocksize);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
    BBX_REQUIRE(e == cudaSuccess, "bbx_rice_decode16: %s", cudaGetErrorString(e));
    const int per_block = RICE_WARPS * RICE_TILES;
    const int blocks = (ntiles + per_block - 1) / per_block;
    if (unsigned16)
        rice16_decode_kernel<true><<<blocks, RICE_WARPS * 32, 0, st>>>((const uint8_t *)heap, heap_bytes, offs, lens,
                                                                       ntiles, nx, (uint16_t *)out, status);
    else
        rice16_decode_kernel<false><<<blocks, RICE_WARPS * 32, 0, st>>>((const uint8_t *)heap, heap_bytes, offs, lens,
                                                                        ntiles, nx, (uint16_t *)out, status);
    BBX_CHECK_LAUNCH("bbx_rice_decode16");
    return 0;
}
