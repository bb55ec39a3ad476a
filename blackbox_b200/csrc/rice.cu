// rice.cu -- tile-compressed FITS images (".fits.fz", ZCMPTYPE = 'RICE_1') on the device, both
// ways.  Decoding: the raw frames the reference reads with read_hdulist
// (blackbox.py:1451), its fpacked bad-pixel masks (Settings/set_blackbox.py:187-193, read in
// mask_init, blackbox.py:4386-4398) and the fpacked reduced calibration frames master_prep lists
// (blackbox.py:4698-4730) -- astropy / CFITSIO unpack them on the host there.  Encoding: the mask
// the reference writes as a losslessly fpacked uint8 image (blackbox.py:826-827: `fpack -D -Y`),
// so that what crosses PCIe on the way out is the compressed product, not 111 MB of mostly zeros.
// The host parses / writes the binary-table header and moves bytes (blackbox_b200/fitsio.py).
//
// Format (FITS tiled-image convention; Rice algorithm as published with CFITSIO -- fits_rcomp /
// fits_rdecomp, _short, _byte -- and in White & Becker / Pence et al. 2009), BYTEPIX = 1 / 2 / 4:
//   tile   = first pixel as a big-endian 8 / 16 / 32-bit value, then blocks of 32 pixel DIFFERENCES
//   block  = a 3 / 4 / 5-bit code FS+1, then per pixel
//              code 0             : all differences are 0 (no further bits)
//              code FSMAX+1       : the difference as 8 / 16 / 32 raw bits   (FSMAX = 6 / 14 / 25)
//              otherwise          : (diff >> FS) zeros, a one, then the low FS bits
//   diff   = zig-zag mapped (even = +d/2, odd = ~(d >> 1)) difference to the previous pixel,
//            modulo 2^(8 BYTEPIX); bits are packed MSB first.
//   FS     = number of bits of ((sum(diff) - nblock/2 - 1) / nblock) >> 1   (the encoder's choice)
//
// DECODER.  The bit position of a block is only known once the block before it is decoded, so a
// tile is a serial job: one thread per tile (a frame has 10600 of them), all 32 lanes of a warp
// decoding, each its own tile.  What decides the speed is the dependent chain per pixel, so:
//   * the stream is read through a 64-bit window (bit 63 = next bit), refilled 32 bits at a time
//     from a per-lane ring in shared memory that cp.async keeps 31 chunks ahead of the reader --
//     no global-memory latency sits on the chain (round 1 loaded a word when the window ran dry
//     and the whole warp waited for it: 390 cycles per pixel);
//   * a pixel is decoded without branches: count-leading-zeros of the top 32 bits gives the unary
//     part, two shifts the FS low bits; "all zero" and "raw" blocks are selects on the same
//     values; only a code longer than 32 bits (a rare outlier) takes a side path;
//   * pixels are packed in registers and leave as one 16-byte store per 16 / 8 / 4 pixels.
// ENCODER.  A warp per tile, a lane per pixel of the 32-pixel block: differences, block sum and FS
// by shuffles, code lengths by a warp scan, the codes OR-ed into a shared-memory bit buffer,
// whole words flushed to a fixed-stride scratch row; a scan over the tile sizes and a compaction
// pass then pack the tiles back to back into the heap (what a FITS binary table wants).
#include "bbx_common.cuh"

template <int BP> struct RiceP;
template <> struct RiceP<1> { static constexpr int FSBITS = 3, FSMAX = 6, BBITS = 8; typedef uint8_t T; };
template <> struct RiceP<2> { static constexpr int FSBITS = 4, FSMAX = 14, BBITS = 16; typedef uint16_t T; };
template <> struct RiceP<4> { static constexpr int FSBITS = 5, FSMAX = 25, BBITS = 32; typedef uint32_t T; };

#define RICE_BLOCK 32

// ---------------------------------------------------------------------------------------------
// decoder
// ---------------------------------------------------------------------------------------------
// Per-lane ring of compressed bytes in shared memory: RDEC_CHUNKS 16-byte chunks per lane, chunk c
// of lane l at ring[c % RDEC_CHUNKS][l].  It is topped up with cp.async (global -> shared, no
// register in between) once per 32-pixel block -- the one place where all lanes of the warp are at
// the same point of the program -- and a block later the data are there: no global-memory latency
// on the per-pixel chain.  (Round 1 loaded a word into a register when a lane's window ran dry;
// with 32 lanes running dry at different pixels the warp waited for SOME lane's load at almost
// every step: 390-800 cycles per pixel.)  A block consumes at most 9 chunks (see rice_block_bytes),
// the ring is refilled to 31: what a block reads has always landed one block earlier.
#define RDEC_CHUNKS 32

struct RiceIn {
    const uint8_t *heap, *heap_end;
    const uint8_t *next;         // global address of the next chunk to fetch (16-byte aligned)
    uint4 *ring;                 // this lane's column of the ring: chunk c at ring[(c % RDEC_CHUNKS) * 32]
    uint32_t fetched;            // chunks requested so far
    uint32_t rdw;                // index (in 32-bit words from the first chunk) of the word held in w0
    uint32_t w0, w1;             // words rdw and rdw + 1, read from the ring well before they are needed
    unsigned long long win;      // bit 63 = next bit of the stream
    int have;                    // valid bits in win
    long long popped;            // bits handed to the window so far

    // one chunk global -> ring; bytes outside the heap read as 0xff (a one bit ends every unary
    // run, so a corrupt tile cannot run away) and are never touched
    __device__ __forceinline__ void fetch()
    {
        uint4 *dst = ring + (size_t)(fetched % RDEC_CHUNKS) * 32;
        const uint8_t *a = next;
        if (a >= heap && a + 16 <= heap_end) {
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(a) : "memory");
        } else {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 3; k >= 0; k--) {
                    const uint8_t *b = a + 4 * i + k;
                    v = (v << 8) | ((b >= heap && b < heap_end) ? (uint32_t)__ldg(b) : 0xffu);
                }
                w[i] = v;
            }
            *dst = make_uint4(w[0], w[1], w[2], w[3]);
        }
        next += 16; fetched++;
    }
    __device__ __forceinline__ void landed() const { asm volatile("cp.async.wait_all;" ::: "memory"); }
    // refill the ring up to RDEC_CHUNKS - 1 chunks ahead of the chunk being read
    __device__ __forceinline__ void top_up()
    {
        const uint32_t want = (rdw >> 2) + (RDEC_CHUNKS - 1);
        while (fetched < want) fetch();
    }
    __device__ __forceinline__ uint32_t ring_word(uint32_t wi) const
    {
        const uint32_t *c = reinterpret_cast<const uint32_t *>(ring + (size_t)((wi >> 2) % RDEC_CHUNKS) * 32);
        return c[wi & 3];
    }
    __device__ __forceinline__ void open(const uint8_t *start)
    {
        const uintptr_t s = (uintptr_t)start;
        next = reinterpret_cast<const uint8_t *>(s & ~(uintptr_t)15);
        fetched = 0;
        rdw = (uint32_t)((s & 15) >> 2);
        top_up();
        landed();
        w0 = ring_word(rdw);
        w1 = ring_word(rdw + 1);
        win = 0; have = 0; popped = 0;
        refill();
        const int skip = (int)(s & 3) * 8;            // bytes of the first word in front of the tile
        win <<= skip; have -= skip; popped -= skip;
        refill();
    }
    // Branch-free: every lane runs the same instructions whether or not its window needs a word
    // (lanes run dry at different pixels; a branch here would split the warp at every step).  The
    // shared-memory read is kept OFF the dependent chain window -> code length -> window: the word
    // that goes into the window was read two refills ago (w0), the read issued here (w1) is only
    // moved between registers at the next step.
    __device__ __forceinline__ void refill()
    {
        const bool need = have <= 32;
        const unsigned long long add = (unsigned long long)__byte_perm(w0, 0, 0x0123) << ((32 - have) & 63);   // big-endian bit order
        win |= need ? add : 0ull;
        have += need ? 32 : 0;
        popped += need ? 32 : 0;
        rdw += need ? 1u : 0u;
        w0 = need ? w1 : w0;
        w1 = ring_word(rdw + 1);
    }
    __device__ __forceinline__ void drop(int n) { win <<= n; have -= n; }              // n in [0, 32]
    __device__ __forceinline__ uint32_t take(int n)                                    // n in [1, 32]
    {
        const uint32_t v = (uint32_t)(win >> (64 - n));
        drop(n);
        return v;
    }
    __device__ __forceinline__ long long consumed() const { return popped - have; }
};

// FLIP: xor of the sign bit (BZERO = 32768 frames: stored int16 -> unsigned counts)
template <int BP, bool FLIP>
__global__ void __launch_bounds__(32)
rice_decode_kernel(const uint8_t *__restrict__ heap, size_t heap_bytes, const long long *__restrict__ offs,
                   const int *__restrict__ lens, int ntiles, int nx, typename RiceP<BP>::T *__restrict__ out,
                   int vec_ok, int *__restrict__ status)
{
    typedef RiceP<BP> P;
    constexpr int G = 16 / BP;                                  // pixels per 16-byte store
    constexpr uint32_t VMASK = BP == 4 ? 0xffffffffu : ((1u << (P::BBITS & 31)) - 1u);
    __shared__ uint4 ring[RDEC_CHUNKS][32];
    const int tile = blockIdx.x * 32 + threadIdx.x;
    if (tile >= ntiles) return;

    RiceIn r;
    r.heap = heap; r.heap_end = heap + heap_bytes;
    r.ring = &ring[0][threadIdx.x];
    const long long o = offs[tile];
    const int n = lens[tile];
    if (o < 0 || n < BP + 1 || (unsigned long long)o + (unsigned long long)n > heap_bytes) {
        atomicOr(status, 2);
        return;
    }
    r.open(heap + o);
    uint32_t lastpix = r.take(P::BBITS);
    typename P::T *row = out + (size_t)tile * nx;

    for (int i = 0; i < nx; i += RICE_BLOCK) {
        const int nthis = min(RICE_BLOCK, nx - i);
        r.landed();                                             // what the last block asked for is there
        r.top_up();                                             // ask for what the next blocks will read
        r.refill();
        const int fs = (int)r.take(P::FSBITS) - 1;
        const bool raw = fs == P::FSMAX, zero = fs < 0;
        const int fsn = max(fs, 0);
        // one pixel: window -> difference -> pixel value
        auto pixel = [&]() -> uint32_t {
            r.refill();                                         // at least 33 bits in the window
            const uint32_t top = (uint32_t)(r.win >> 32);
            const int z = __clz(top);                           // 32 if the top half is all zeros
            const int len = raw ? P::BBITS : zero ? 0 : z + 1 + fsn;
            uint32_t diff;
            if (len <= 32) {
                const unsigned long long rest = r.win << (z + 1);
                const uint32_t low = (uint32_t)((rest >> 1) >> (63 - fsn));               // fs = 0: nothing
                diff = raw ? (top >> (32 - P::BBITS)) : zero ? 0u : (((uint32_t)z << fsn) | low);
                r.drop(len);
            } else {
                // a code longer than 32 bits: count the zeros across refills, then the low bits.
                // (Bounded by the ring: a valid block never needs more than 9 chunks.)
                uint32_t nz = 0;
                for (int guard = 0; guard < 64; guard++) {
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
            lastpix = (lastpix + diff) & VMASK;
            return FLIP ? (lastpix ^ (1u << (P::BBITS - 1))) : lastpix;
        };
        if (nthis == RICE_BLOCK && vec_ok) {
            // a whole block, rows 16-byte aligned: no per-pixel bookkeeping, 16-byte stores
#pragma unroll 1
            for (int g = 0; g < RICE_BLOCK; g += G) {
                uint32_t pw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < G; k++) pw[(k * BP) >> 2] |= pixel() << (8 * ((k * BP) & 3));
                *reinterpret_cast<uint4 *>(row + i + g) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < nthis; k++) row[i + k] = (typename P::T)pixel();
        }
    }
    // bits consumed beyond the tile's own bytes: a truncated or corrupt tile
    if (r.consumed() > 8ll * n) atomicOr(status, 1);
}

template <int BP>
static int rice_decode_launch(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                              int ntiles, int nx, int flip, void *out, int *status, cudaStream_t st)
{
    typedef typename RiceP<BP>::T T;
    const int vec_ok = (((size_t)nx * BP) % 16 == 0) && (((uintptr_t)out) % 16 == 0);
    const int blocks = (ntiles + 31) / 32;
    if (flip)
        rice_decode_kernel<BP, true><<<blocks, 32, 0, st>>>((const uint8_t *)heap, heap_bytes, offs, lens, ntiles, nx,
                                                           (T *)out, vec_ok, status);
    else
        rice_decode_kernel<BP, false><<<blocks, 32, 0, st>>>((const uint8_t *)heap, heap_bytes, offs, lens, ntiles, nx,
                                                            (T *)out, vec_ok, status);
    BBX_CHECK_LAUNCH("rice_decode_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bbx_rice_decode: heap (device, heap_bytes), offs / lens (device, one per tile: byte offset into
// the heap and compressed length), ntiles row tiles of nx pixels each (ZTILE1 = ZNAXIS1, ZTILE2 =
// 1), BLOCKSIZE 32, BYTEPIX 1 / 2 / 4.  flip_sign != 0: the sign bit of every pixel is inverted
// (BYTEPIX 2 with BZERO 32768: stored int16 -> uint16 counts).  status (device int, zeroed by the
// call): bit 0 = a tile ran past its compressed bytes, bit 1 = a descriptor points outside the
// heap; the affected rows are undefined.  The caller reads it after synchronising.
// ---------------------------------------------------------------------------------------------
extern "C" int bbx_rice_decode(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                               int ntiles, int nx, int blocksize, int bytepix, int flip_sign, void *out, int *status,
                               void *stream)
{
    BBX_REQUIRE(heap && offs && lens && out && status, "bbx_rice_decode: null argument");
    BBX_REQUIRE(ntiles > 0 && nx > 0, "bbx_rice_decode: %d tiles of %d pixels", ntiles, nx);
    BBX_REQUIRE(blocksize == RICE_BLOCK, "bbx_rice_decode: BLOCKSIZE %d (32 is supported)", blocksize);
    BBX_REQUIRE(bytepix == 1 || bytepix == 2 || bytepix == 4, "bbx_rice_decode: BYTEPIX %d", bytepix);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
    BBX_REQUIRE(e == cudaSuccess, "bbx_rice_decode: %s", cudaGetErrorString(e));
    if (bytepix == 1) return rice_decode_launch<1>(heap, heap_bytes, offs, lens, ntiles, nx, flip_sign, out, status, st);
    if (bytepix == 2) return rice_decode_launch<2>(heap, heap_bytes, offs, lens, ntiles, nx, flip_sign, out, status, st);
    return rice_decode_launch<4>(heap, heap_bytes, offs, lens, ntiles, nx, flip_sign, out, status, st);
}

extern "C" int bbx_rice_decode16(const void *heap, size_t heap_bytes, const long long *offs, const int *lens,
                                 int ntiles, int nx, int blocksize, int unsigned16, void *out, int *status,
                                 void *stream)
{
    return bbx_rice_decode(heap, heap_bytes, offs, lens, ntiles, nx, blocksize, 2, unsigned16, out, status, stream);
}

// ---------------------------------------------------------------------------------------------
// Un-quantising a float image (FITS tiled-image convention, "quantization of floating-point
// data", with the subtractive dithering of Pence, White & Seaman 2010 and the random sequence
// of the FITS standard's appendix on the tiled-image random number generator):
//   NO_DITHER               value = q * ZSCALE + ZZERO
//   SUBTRACTIVE_DITHER_1/2  value = (q - R[i] + 0.5) * ZSCALE + ZZERO, R the 10000 published random
//                           numbers, restarted per tile at R[(tile + ZDITHER0 - 1) mod 10000] * 500
//                           (tile counted from 0); DITHER_2: q = -2147483646 is an exact 0
//   q = ZBLANK (-2147483647 by default) -> NaN
// One tile = one image row; zscale / zzero: float64 per tile (the table's ZSCALE / ZZERO columns).
// ---------------------------------------------------------------------------------------------
#define RICE_NRANDOM 10000
#define RICE_NULL_VALUE (-2147483647)
#define RICE_ZERO_VALUE (-2147483646)

__global__ void __launch_bounds__(256)
unquantize_kernel(const int32_t *__restrict__ q, int ntiles, int nx, const double *__restrict__ zscale,
                  const double *__restrict__ zzero, const float *__restrict__ rnd, int dither, int zdither0,
                  int zblank, int have_blank, float *__restrict__ out)
{
    const int tile = blockIdx.y;
    if (tile >= ntiles) return;
    const double scale = zscale[tile], zero = zzero[tile];
    const int32_t *src = q + (size_t)tile * nx;
    float *dst = out + (size_t)tile * nx;
    // position in the random sequence of pixel x of this tile: the sequence restarts at
    // nextrand0 = int(R[iseed] * 500) and, each time it runs off the end of the table, moves to
    // the next iseed and restarts from int(R[iseed] * 500) -- walked here in closed form per
    // segment, so that pixels are independent of each other
    int iseed = (int)(((long long)tile + zdither0 - 1) % RICE_NRANDOM);
    if (iseed < 0) iseed += RICE_NRANDOM;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nx; x += gridDim.x * blockDim.x) {
        const int32_t v = src[x];
        float r;
        if (have_blank && v == zblank) {
            r = nanf("");
        } else if (dither == 0) {
            r = (float)((double)v * scale + zero);
        } else if (dither == 2 && v == RICE_ZERO_VALUE) {
            r = 0.0f;
        } else {
            int seed = iseed, nextrand = (int)(rnd[seed] * 500.0f), left = x;
            while (left >= RICE_NRANDOM - nextrand) {          // at most a few segments per row
                left -= RICE_NRANDOM - nextrand;
                seed = seed + 1 == RICE_NRANDOM ? 0 : seed + 1;
                nextrand = (int)(rnd[seed] * 500.0f);
            }
            r = (float)(((double)v - (double)rnd[nextrand + left] + 0.5) * scale + zero);
        }
        dst[x] = r;
    }
}

extern "C" int bbx_unquantize(const int32_t *q, int ntiles, int nx, const double *zscale, const double *zzero,
                              const float *rand10000, int dither, int zdither0, int zblank, int have_blank,
                              float *out, void *stream)
{
    BBX_REQUIRE(q && zscale && zzero && out, "bbx_unquantize: null argument");
    BBX_REQUIRE(ntiles > 0 && nx > 0, "bbx_unquantize: %d tiles of %d pixels", ntiles, nx);
    BBX_REQUIRE(dither >= 0 && dither <= 2, "bbx_unquantize: dither method %d", dither);
    BBX_REQUIRE(dither == 0 || rand10000, "bbx_unquantize: dithering needs the random table");
    dim3 grid((unsigned)min(8, (nx + 255) / 256), (unsigned)ntiles);
    unquantize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, ntiles, nx, zscale, zzero, rand10000, dither,
                                                              zdither0, zblank, have_blank, out);
    BBX_CHECK_LAUNCH("unquantize_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// encoder
// ---------------------------------------------------------------------------------------------
// Longest possible block.  With FS taken from the block sum S (n = 32 pixels): FS >= 1 means
// floor((S - n/2 - 1)/n) < 2^(FS+1), so the unary parts hold sum(diff >> FS) <= S / 2^FS < 2n + n/4 + 1
// = 73 zeros; FS = 0 means S < 2n + n/2 + 1 = 81.  A coded block is therefore at most
// FSBITS + 32 (FSMAX - 1 + 1) + 80 bits, a raw one FSBITS + 32 BBITS:
//   BYTEPIX 1: 275 bits (35 B)   BYTEPIX 2: 532 bits (67 B)   BYTEPIX 4: 1029 bits (129 B)
// plus at most 31 carried bits in front: 34 words.
#define RENC_WARPS 4
#define RENC_WORDS 40
__host__ __device__ static inline int rice_block_bytes(int bp) { return bp == 1 ? 35 : bp == 2 ? 67 : 129; }

// OR the `nbits` (1..32) low bits of `value` into a cleared MSB-first bit buffer at bit `pos`
__device__ __forceinline__ void put_bits(uint32_t *buf, int pos, uint32_t value, int nbits)
{
    const int w = pos >> 5, off = pos & 31;
    const unsigned long long v = (unsigned long long)value << (64 - off - nbits);       // off + nbits <= 63
    atomicOr(&buf[w], (uint32_t)(v >> 32));
    const uint32_t lo = (uint32_t)v;
    if (lo) atomicOr(&buf[w + 1], lo);
}

template <int BP> __device__ __forceinline__ int rice_signed(typename RiceP<BP>::T v)
{
    // fits_rcomp_byte / _short / fits_rcomp see the pixels as signed char / short / int
    return BP == 1 ? (int)(int8_t)v : BP == 2 ? (int)(int16_t)v : (int)v;
}

// One warp per tile (grid-stride), a lane per pixel of the block.  scratch: ntiles rows of
// `stride` bytes (16-byte multiples); out_lens[t] = compressed bytes of tile t.
template <int BP>
__global__ void __launch_bounds__(RENC_WARPS * 32)
rice_encode_kernel(const typename RiceP<BP>::T *__restrict__ img, int ntiles, int nx, uint8_t *__restrict__ scratch,
                   size_t stride, int *__restrict__ out_lens)
{
    typedef RiceP<BP> P;
    typedef typename P::T T;
    __shared__ uint32_t sbuf[RENC_WARPS][RENC_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *buf = sbuf[warp];
    const unsigned FULL = 0xffffffffu;
    for (int tile = blockIdx.x * RENC_WARPS + warp; tile < ntiles; tile += gridDim.x * RENC_WARPS) {
        const T *row = img + (size_t)tile * nx;
        uint32_t *dst = reinterpret_cast<uint32_t *>(scratch + (size_t)tile * stride);
        int wpos = 0;                       // whole 32-bit words already written for this tile
        uint32_t carry;                     // the partly filled word (its top cbits bits are valid)
        int cbits;
        const T first = row[0];
        if (BP == 4) {
            if (lane == 0) dst[0] = __byte_perm((uint32_t)first, 0, 0x0123);
            wpos = 1; carry = 0; cbits = 0;
        } else {
            carry = (uint32_t)first << (32 - P::BBITS); cbits = P::BBITS;
        }
        int lastpix = rice_signed<BP>(first);
        int cur = (lane < nx) ? rice_signed<BP>(row[lane]) : 0;
        for (int i = 0; i < nx; i += RICE_BLOCK) {
            const int nthis = min(RICE_BLOCK, nx - i);
            // the next block's pixel is requested before this block is coded
            const int inext = i + RICE_BLOCK + lane;
            const int nxt = (inext < nx) ? rice_signed<BP>(row[inext]) : 0;
            int prev = __shfl_up_sync(FULL, cur, 1);
            if (lane == 0) prev = lastpix;
            // difference in the pixel's own width (it wraps), zig-zag mapped
            int pd = (int)((unsigned)cur - (unsigned)prev);
            if (BP == 1) pd = (int)(int8_t)pd;
            if (BP == 2) pd = (int)(int16_t)pd;
            uint32_t diff = (pd < 0) ? ~((uint32_t)pd << 1) : ((uint32_t)pd << 1);
            if (lane >= nthis) diff = 0;
            lastpix = __shfl_sync(FULL, cur, nthis - 1);
            cur = nxt;
            if (__ballot_sync(FULL, diff != 0) == 0) {
                // all differences zero: the code 0 and nothing else
                cbits += P::FSBITS;
                if (cbits >= 32) {
                    if (lane == 0) dst[wpos] = __byte_perm(carry, 0, 0x0123);
                    wpos++; carry = 0; cbits -= 32;
                }
                continue;
            }
            // block sum (exact: 32 values below 2^32), FS as fits_rcomp computes it in double
            unsigned long long sum = diff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
            double dpsum = ((double)sum - (double)(nthis / 2) - 1.0) / (double)nthis;
            if (dpsum < 0) dpsum = 0.0;
            const unsigned long long ip = (unsigned long long)dpsum;
            uint32_t psum = (BP == 1 ? (uint32_t)(uint8_t)ip : BP == 2 ? (uint32_t)(uint16_t)ip : (uint32_t)ip) >> 1;
            int fs = 0;
            while (psum > 0) { psum >>= 1; fs++; }
            const bool raw = fs >= P::FSMAX;
            const uint32_t top = raw ? 0u : (diff >> fs);
            const int len = (lane < nthis) ? (raw ? P::BBITS : (int)top + 1 + fs) : 0;
            // bit position of this lane's code: carry + FS code + the codes of the lanes before it
            int pos = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, pos, o);
                if (lane >= o) pos += t;
            }
            const int total = cbits + P::FSBITS + __shfl_sync(FULL, pos, 31);
            pos = cbits + P::FSBITS + pos - len;
            buf[lane] = (lane == 0) ? carry : 0u;
            if (lane + 32 < RENC_WORDS) buf[lane + 32] = 0u;
            __syncwarp();
            if (lane == 0) put_bits(buf, cbits, raw ? (uint32_t)(P::FSMAX + 1) : (uint32_t)(fs + 1), P::FSBITS);
            if (lane < nthis) {
                // what has to be written of a code: raw -> the BBITS bits of diff; else a one and
                // the low FS bits, ending at pos + len (the zeros in front are the cleared buffer)
                if (raw) put_bits(buf, pos, BP == 4 ? diff : (diff & ((1u << (P::BBITS & 31)) - 1u)), P::BBITS);
                else put_bits(buf, pos + (int)top, (1u << fs) | (diff & ((1u << fs) - 1u)), fs + 1);
            }
            __syncwarp();
            const int nfull = total >> 5;
            for (int w = lane; w < nfull; w += 32) dst[wpos + w] = __byte_perm(buf[w], 0, 0x0123);
            wpos += nfull;
            carry = buf[nfull];
            cbits = total & 31;
            __syncwarp();
        }
        // flush the partly filled word: the stream ends on a byte boundary, zero padded
        const int tail_bytes = (cbits + 7) >> 3;
        if (lane == 0) {
            uint8_t *tb = reinterpret_cast<uint8_t *>(dst + wpos);
            for (int b = 0; b < tail_bytes; b++) tb[b] = (uint8_t)(carry >> (24 - 8 * b));
            out_lens[tile] = 4 * wpos + tail_bytes;
        }
    }
}

// exclusive prefix sum of the tile sizes (one CTA; ntiles is ~10^4): offs[t], hdr = {total bytes,
// ntiles, status}
struct RiceOutHdr { long long total; int ntiles; int status; };

__global__ void __launch_bounds__(1024)
rice_scan_kernel(const int *__restrict__ lens, int ntiles, long long *__restrict__ offs, RiceOutHdr *hdr,
                 long long heap_cap)
{
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int per = (ntiles + 1023) / 1024;
    const int a = min(t * per, ntiles), b = min(a + per, ntiles);
    long long s = 0;
    for (int i = a; i < b; i++) s += lens[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                    // Hillis-Steele inclusive scan
        const long long v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;
    for (int i = a; i < b; i++) { offs[i] = run; run += lens[i]; }
    if (t == 1023) {
        hdr->total = part[1023];
        hdr->ntiles = ntiles;
        hdr->status = part[1023] > heap_cap ? 1 : 0;
    }
}

// tiles back to back into the heap: a warp per tile, 32-bit words aligned on the DESTINATION
// (the source rows are 16-byte aligned), the few bytes either side of them one by one
__global__ void __launch_bounds__(256)
rice_compact_kernel(const uint8_t *__restrict__ scratch, size_t stride, const int *__restrict__ lens,
                    const long long *__restrict__ offs, int ntiles, uint8_t *__restrict__ heap, long long heap_cap)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += warps) {
        const int len = lens[tile];
        const long long o = offs[tile];
        if (o + len > heap_cap) continue;                   // reported by the scan kernel's status
        const uint8_t *src = scratch + (size_t)tile * stride;
        uint8_t *dst = heap + o;
        const int head = min(len, (int)((4 - ((uintptr_t)dst & 3)) & 3));
        if (lane < head) dst[lane] = src[lane];
        const int nwords = (len - head) >> 2;
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + head);
        const int sh = 8 * (head & 3);                      // source byte offset of destination word 0 is `head`
        for (int k = lane; k < nwords; k += 32) {
            const int sidx = (head >> 2) + k;               // head < 4 -> 0
            d32[k] = sh ? __funnelshift_r(s32[sidx], s32[sidx + 1], sh) : s32[sidx];
        }
        const int done = head + 4 * nwords;
        if (lane < len - done) dst[done + lane] = src[done + lane];
    }
}

static size_t rice_stride(int nx, int bp)
{
    const size_t nblocks = ((size_t)nx + RICE_BLOCK - 1) / RICE_BLOCK;
    return (bp + nblocks * rice_block_bytes(bp) + 8 + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------------------------
// bbx_rice_encode: img (device; ntiles rows of nx pixels of BYTEPIX 1 / 2 / 4 bytes: row tiles,
// BLOCKSIZE 32) -> out (device):
//     [0:8)   int64  total heap bytes          [8:12) int32 ntiles
//     [12:16) int32  status (bit 0: the heap did not fit into out_bytes; sizes are still valid)
//     [16 : 16 + 4 ntiles)  int32 compressed bytes of every tile (the table's descriptors; the
//                            offsets are their running sum), padded to a multiple of 16
//     then the heap: the tiles back to back.
// work >= bbx_rice_encode_work_bytes(ntiles, nx, bytepix); out_bytes >= 16 + 16 ceil(ntiles / 4),
// bbx_rice_encode_out_bytes(...) is the size that always fits (incompressible data).
// The bytes are those fits_rcomp / _short / _byte produce for the same row.
// ---------------------------------------------------------------------------------------------
extern "C" size_t bbx_rice_encode_work_bytes(int ntiles, int nx, int bytepix)
{
    if (ntiles <= 0 || nx <= 0 || (bytepix != 1 && bytepix != 2 && bytepix != 4)) return 0;
    return (size_t)ntiles * rice_stride(nx, bytepix) + (size_t)ntiles * sizeof(long long) + 16;
}

static size_t rice_heap_offset(int ntiles) { return 16 + ((size_t)ntiles * 4 + 15) / 16 * 16; }

extern "C" size_t bbx_rice_encode_out_bytes(int ntiles, int nx, int bytepix)
{
    if (ntiles <= 0 || nx <= 0 || (bytepix != 1 && bytepix != 2 && bytepix != 4)) return 0;
    return rice_heap_offset(ntiles) + (size_t)ntiles * rice_stride(nx, bytepix);
}

extern "C" int bbx_rice_encode(const void *img, int ntiles, int nx, int bytepix, void *work, size_t work_bytes,
                               void *out, size_t out_bytes, void *stream)
{
    BBX_REQUIRE(img && work && out, "bbx_rice_encode: null argument");
    BBX_REQUIRE(ntiles > 0 && nx > 0, "bbx_rice_encode: %d tiles of %d pixels", ntiles, nx);
    BBX_REQUIRE(bytepix == 1 || bytepix == 2 || bytepix == 4, "bbx_rice_encode: BYTEPIX %d", bytepix);
    BBX_REQUIRE(work_bytes >= bbx_rice_encode_work_bytes(ntiles, nx, bytepix), "bbx_rice_encode: work buffer too small");
    BBX_REQUIRE(out_bytes >= rice_heap_offset(ntiles), "bbx_rice_encode: output buffer too small for the descriptors");
    BBX_REQUIRE(((uintptr_t)work % 16) == 0 && ((uintptr_t)out % 16) == 0, "bbx_rice_encode: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = rice_stride(nx, bytepix);
    uint8_t *scratch = (uint8_t *)work;
    long long *offs = reinterpret_cast<long long *>(scratch + ((size_t)ntiles * stride + 15) / 16 * 16);
    RiceOutHdr *hdr = (RiceOutHdr *)out;
    int *lens = reinterpret_cast<int *>((uint8_t *)out + 16);
    uint8_t *heap = (uint8_t *)out + rice_heap_offset(ntiles);
    const long long heap_cap = (long long)(out_bytes - rice_heap_offset(ntiles));
    const int want = (ntiles + RENC_WARPS - 1) / RENC_WARPS;
    const int blocks = want < BBX_SM_COUNT * 16 ? want : BBX_SM_COUNT * 16;
    if (bytepix == 1) rice_encode_kernel<1><<<blocks, RENC_WARPS * 32, 0, st>>>((const uint8_t *)img, ntiles, nx, scratch, stride, lens);
    else if (bytepix == 2) rice_encode_kernel<2><<<blocks, RENC_WARPS * 32, 0, st>>>((const uint16_t *)img, ntiles, nx, scratch, stride, lens);
    else rice_encode_kernel<4><<<blocks, RENC_WARPS * 32, 0, st>>>((const uint32_t *)img, ntiles, nx, scratch, stride, lens);
    BBX_CHECK_LAUNCH("rice_encode_kernel");
    rice_scan_kernel<<<1, 1024, 0, st>>>(lens, ntiles, offs, hdr, heap_cap);
    BBX_CHECK_LAUNCH("rice_scan_kernel");
    const int cblocks = (ntiles + 7) / 8 < BBX_SM_COUNT * 8 ? (ntiles + 7) / 8 : BBX_SM_COUNT * 8;
    rice_compact_kernel<<<cblocks, 256, 0, st>>>(scratch, stride, lens, offs, ntiles, heap, heap_cap);
    BBX_CHECK_LAUNCH("rice_compact_kernel");
    return 0;
}
