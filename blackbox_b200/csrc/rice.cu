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
// serial job: one thread per tile (a frame has 10600 of them), every step depending on the one
// before.  What decides the speed is therefore (1) that the lanes of a warp, each in its own
// tile, run the SAME instructions -- a decoder written as CFITSIO's byte-at-a-time loops diverges
// completely and the warp executes its 32 tiles one after the other (measured on B200, full
// 10600 x 12000 frame: 3.7 ms; this version 2.45 ms, profiles/r01_rice_bench.txt) -- and (2) the
// length of the dependent chain per pixel, which is what is left (about 390 cycles per pixel and
// tile; two tiles per thread would give the scheduler independent work).  So:
//   * the stream is read through a 64-bit window (bit 63 = next bit) refilled with one aligned
//     32-bit load whenever 32 bits or fewer are left -- one predicated block, no loops;
//   * a pixel is decoded without branches: count-leading-zeros of the top 32 bits gives the unary
//     part, two shifts the FS low bits; the "all zero" and "raw 16 bit" block types are selects
//     on the same values; only a code longer than 32 bits (a rare outlier) takes a side path;
//   * RICE_TILES = 8 lanes of a warp decode (1325 warps for a frame, so every scheduler has
//     warps to switch between); all 32 lanes then store the staged 32-pixel row segments, 64
//     contiguous bytes per row -- whole sectors, although every decoder works on its own row.
#include "bbx_common.cuh"

#define RICE_WARPS 4
#define RICE_TILES 8             // tiles (= decoding lanes) per warp
#define RICE_BLOCK 32
#define RICE_ROW   34            // uint16 per staging row: 17 words, odd, so lanes spread over the banks

struct RiceReader {
    const uint8_t *heap, *heap_end;
    const uint8_t *next;         // next aligned word to load
    unsigned long long win;
    int have;                    // valid bits in win
    long long loaded;            // bits loaded so far

    // 32 bits at an aligned address, big-endian; bytes outside the heap read as 0xff (a one bit
    // ends every unary run, so a corrupt tile cannot loop past the end of the heap)
    __device__ __forceinline__ unsigned int word(const uint8_t *a) const
    {
        if (a >= heap && a + 4 <= heap_end)
            return __byte_perm(__ldg(reinterpret_cast<const unsigned int *>(a)), 0, 0x0123);
        unsigned int w = 0;
        for (int k = 0; k < 4; k++) w = (w << 8) | ((a + k >= heap && a + k < heap_end) ? __ldg(a + k) : 0xffu);
        return w;
    }
    __device__ __forceinline__ void refill()
    {
        if (have <= 32) {
            win |= (unsigned long long)word(next) << (32 - have);
            next += 4; have += 32; loaded += 32;
        }
    }
    __device__ __forceinline__ void drop(int n) { win <<= n; have -= n; }              // n in [0, 32]
    __device__ __forceinline__ unsigned int take(int n)                                // n in [1, 32]
    {
        const unsigned int v = (unsigned int)(win >> (64 - n));
        drop(n);
        return v;
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
    r.heap = heap; r.heap_end = heap + heap_bytes;
    r.next = heap; r.win = 0; r.have = 0; r.loaded = 0;
    bool bad = false;
    long long tile_bits = 0;
    int skip = 0;
    unsigned int lastpix = 0;
    if (live) {
        const long long o = offs[tile];
        const int n = lens[tile];
        if (o < 0 || n < 3 || (unsigned long long)o + (unsigned long long)n > heap_bytes) {
            bad = true;
        } else {
            const uintptr_t start = (uintptr_t)(heap + o);
            skip = (int)(start & 3) * 8;                        // bits in front of the tile in its first word
            r.next = reinterpret_cast<const uint8_t *>(start & ~(uintptr_t)3);
            tile_bits = 8ll * n;
            r.refill();
            r.drop(skip);
            r.refill();
            lastpix = r.take(16);
        }
    }
    const bool work = live && !bad;
    for (int i = 0; i < nx; i += RICE_BLOCK) {
        const int nthis = min(RICE_BLOCK, nx - i);
        uint16_t *row = stage[warp][lane & (RICE_TILES - 1)];
        if (work) {
            r.refill();
            const int fs = (int)r.take(4) - 1;
            const bool raw = fs == 14, zero = fs < 0;
            const int fsn = max(fs, 0);
            for (int k = 0; k < nthis; k++) {
                r.refill();                                     // at least 33 bits in the window
                const unsigned int top = (unsigned int)(r.win >> 32);
                const int z = __clz(top);                       // 32 if the top half is all zeros
                const int len = raw ? 16 : zero ? 0 : z + 1 + fsn;
                unsigned int diff;
                if (len <= 32) {
                    const unsigned long long rest = r.win << (z + 1);
                    const unsigned int low = (unsigned int)((rest >> 1) >> (63 - fsn));   // fs = 0: nothing
                    diff = raw ? (top >> 16) : zero ? 0u : (((unsigned int)z << fsn) | low);
                    r.drop(len);
                } else {
                    // a code longer than 32 bits: count the zeros across refills, then the low bits
                    unsigned int nz = 0;
                    for (;;) {
                        r.refill();
                        if (r.win == 0) { nz += r.have; r.have = 0; continue; }
                        const int zz = __clzll((long long)r.win);
                        nz += zz;
                        r.win = (r.win << zz) << 1; r.have -= zz + 1;
                        break;
                    }
                    r.refill();
                    diff = nz << fsn;
                    if (fsn > 0) diff |= r.take(fsn);
                }
                diff = (diff & 1u) ? ~(diff >> 1) : (diff >> 1);
                lastpix = (lastpix + diff) & 0xffffu;
                row[k] = (uint16_t)lastpix;
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
    // bits consumed beyond the tile's own bytes: a truncated or corrupt tile
    const bool overrun = work && (r.loaded - r.have - skip > tile_bits);
    if (live && (bad || overrun)) atomicOr(status, bad ? 2 : 1);
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
