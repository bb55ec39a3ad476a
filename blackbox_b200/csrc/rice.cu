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
// ENCODER.  A warp per tile, 1024 pixels at a time, a lane per 32-pixel block: differences, block
// sum, FS and code lengths in the lane's registers, the bit positions of the 32 blocks by a warp
// scan, the codes OR-ed into a shared-memory bit buffer, whole words flushed to a fixed-stride
// scratch row; a scan over the tile sizes and a compaction pass then pack the tiles back to back
// into the heap (what a FITS binary table wants).
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
    if (n == 0) return;                     // stored in the table's fall-back column: the host fills the row
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
// The encoder's bit buffer holds 32 blocks and at most 31 carried bits in front: 1031 words (+ the
// word the last lane's window may flush behind its last bit).
#define RENC_WARPS 4
#define RENC_CHUNK (32 * RICE_BLOCK)    // pixels a warp codes at a time: one block per lane
#define RENC_WORDS 1034
__host__ __device__ static inline int rice_block_bytes(int bp) { return bp == 1 ? 35 : bp == 2 ? 67 : 129; }

// Bit writer of one lane: codes are appended MSB first to a 64-bit window; whole 32-bit words go
// to the warp's cleared bit buffer by atomicOr (the first and the last word of a block are shared
// with the neighbouring lanes' blocks).
struct BitOut {
    uint32_t *buf;
    int w, n;                               // word index; n (< 32 between calls) pending bits at the top of acc
    unsigned long long acc;
    __device__ __forceinline__ void open(uint32_t *b, int pos) { buf = b; w = pos >> 5; n = pos & 31; acc = 0; }
    __device__ __forceinline__ void flush_word()
    {
        const uint32_t hi = (uint32_t)(acc >> 32);
        if (hi) atomicOr(&buf[w], hi);
        w++; acc <<= 32; n -= 32;
    }
    __device__ __forceinline__ void put(uint32_t value, int nbits)             // 1 <= nbits <= 32
    {
        acc |= (unsigned long long)value << (64 - n - nbits);
        n += nbits;
        if (n >= 32) flush_word();
    }
    __device__ __forceinline__ void zeros(int nbits)
    {
        n += nbits;
        while (n >= 32) flush_word();
    }
    __device__ __forceinline__ void close() { if (n > 0) { n += 32; flush_word(); } }
};

template <int BP> __device__ __forceinline__ int rice_signed(typename RiceP<BP>::T v)
{
    // fits_rcomp_byte / _short / fits_rcomp see the pixels as signed char / short / int
    return BP == 1 ? (int)(int8_t)v : BP == 2 ? (int)(int16_t)v : (int)v;
}

// Where the encoder's pixels come from.  PlainSrc: an integer image.  QuantSrc: a float32 image
// quantised on the fly as fits_quantize_float does (SUBTRACTIVE_DITHER_1; zscale[tile] == 0 marks
// a row that is not quantised: it gets no Rice-coded bytes) -- the int32 image never exists.
template <int BP> struct PlainSrc {
    const typename RiceP<BP>::T *img;
    int nx;
    const typename RiceP<BP>::T *row;
    __device__ __forceinline__ bool open(int tile, int) { row = img + (size_t)tile * nx; return true; }
    __device__ __forceinline__ int px(int x) { return rice_signed<BP>(row[x]); }
    // the 1024 pixels from i0 on into vals (block k at vals[33 k]); pixels beyond the row repeat
    // `lastpix`.  -> does any of them differ from lastpix?
    __device__ __forceinline__ bool fill(int *vals, int i0, int lane, int lastpix)
    {
        bool differs = false;
#pragma unroll 16
        for (int k = 0; k < 32; k++) {
            const int x = i0 + 32 * k + lane;
            const int v = (x < nx) ? px(x) : lastpix;
            differs |= v != lastpix;
            vals[k * 33 + lane] = v;
        }
        return differs;
    }
};

struct QuantSrc {
    const float *img;
    const double *zscale, *zzero;
    const float *rnd;
    int nx, zdither0;
    const float *row;
    double scale, zero;
    float zero_f, inv_f, guard0;
    int seed, nextrand, base;               // position in the random sequence: R[nextrand + (x - base)]
    __device__ __forceinline__ bool open(int tile, int)
    {
        scale = zscale[tile];
        if (scale == 0.0) return false;
        zero = zzero[tile];
        const double inv = 1.0 / scale;
        zero_f = (float)zero;
        inv_f = (float)inv;
        guard0 = (float)((fabs(zero) * inv + 8.0) * 1.1920928955078125e-07);      // 2^-23
        row = img + (size_t)tile * nx;
        seed = (int)(((long long)tile + zdither0 - 1) % RICE_NRANDOM);
        if (seed < 0) seed += RICE_NRANDOM;
        nextrand = (int)(rnd[seed] * 500.0f);
        base = 0;
        return true;
    }
    // x never decreases from call to call (per lane)
    __device__ __forceinline__ void advance(int x)
    {
        while (x - base >= RICE_NRANDOM - nextrand) {
            base += RICE_NRANDOM - nextrand;
            seed = seed + 1 == RICE_NRANDOM ? 0 : seed + 1;
            nextrand = (int)(rnd[seed] * 500.0f);
        }
    }
    // NINT((v - zero) / scale + R - 0.5) as CFITSIO evaluates it, in double
    __device__ __forceinline__ int exact(float pix, float rr) const
    {
        const double v = ((double)pix - zero) / scale + (double)rr - 0.5;
        return (v >= 0.0) ? (int)(v + 0.5) : (int)(v - 0.5);
    }
    // The same in float32 first: with Z = |zero| / scale the float value is off by less than
    // (Z + 5 |x| + 3) 2^-24, which can change the integer only if x sits that close to a
    // half-integer -- the guard is twice that, and the few pixels inside it (one in some thousands)
    // take the double-precision statement (`unsure`).
    __device__ __forceinline__ int quick(float pix, float rr, bool &unsure) const
    {
        const float xf = (pix - zero_f) * inv_f + rr - 0.5f;
        const float g = guard0 + fabsf(xf) * 9.5367431640625e-07f;               // 8 * 2^-23
        const float yf = xf + 0.5f, ff = yf - floorf(yf);
        unsure = ff < g || ff > 1.0f - g;
        return (int)(xf + copysignf(0.5f, xf));
    }
    __device__ __forceinline__ int px(int x)
    {
        advance(x);
        const float pix = row[x], rr = rnd[nextrand + (x - base)];
        bool unsure;
        const int v = quick(pix, rr, unsure);
        return unsure ? exact(pix, rr) : v;
    }
    __device__ __forceinline__ bool fill(int *vals, int i0, int lane, int lastpix)
    {
        bool differs = false;
        const int x0 = i0 + lane;
        bool straight = i0 + RENC_CHUNK <= nx;
        if (straight) {
            advance(x0);
            straight = (x0 + 31 * 32 - base) < RICE_NRANDOM - nextrand;          // no restart of the sequence inside
        }
        if (__all_sync(0xffffffffu, straight)) {
            // a full chunk, one run of the random sequence: straight-line code, the unsure pixels afterwards
            const float *rp = rnd + (nextrand - base);
            unsigned int redo = 0;
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const int x = x0 + 32 * k;
                bool unsure;
                const int v = quick(row[x], rp[x], unsure);
                redo |= (unsigned int)unsure << k;
                differs |= v != lastpix;
                vals[k * 33 + lane] = v;
            }
            while (redo) {
                const int k = __ffs(redo) - 1, x = x0 + 32 * k;
                redo &= redo - 1;
                const int v = exact(row[x], rp[x]);
                differs |= v != lastpix;
                vals[k * 33 + lane] = v;
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < 32; k++) {
                const int x = x0 + 32 * k;
                const int v = (x < nx) ? px(x) : lastpix;
                differs |= v != lastpix;
                vals[k * 33 + lane] = v;
            }
        }
        return differs;
    }
};

// One warp per tile (grid-stride).  The warp codes 1024 pixels at a time: first, with the lanes
// side by side along the row (coalesced loads; QuantSrc quantises here), the pixel values go into
// shared memory, one 32-pixel block per row of a 33-word-pitch array; then every lane codes ONE
// block on its own -- differences, block sum, FS and the code lengths in registers -- a warp scan
// of the 32 block lengths gives every lane its bit position, and the lanes OR their codes into the
// warp's (cleared) bit buffer; whole words are flushed to a fixed-stride scratch row.  Round 2's
// first encoder gave a lane one PIXEL of a block: seven warp-wide steps (scan, shuffles, syncs)
// per 32 pixels, ~170 instructions; this way it is ~30.
// scratch: ntiles rows of `stride` bytes (16-byte multiples); out_lens[t] = compressed bytes of tile t.
template <int BP, typename SRC>
__global__ void __launch_bounds__(RENC_WARPS * 32, 6)
rice_encode_kernel(SRC src, int ntiles, int nx, uint8_t *__restrict__ scratch, size_t stride, int *__restrict__ out_lens)
{
    typedef RiceP<BP> P;
    __shared__ int svals[RENC_WARPS][32 * 33];
    __shared__ uint32_t sbuf[RENC_WARPS][RENC_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *vals = svals[warp];
    uint32_t *buf = sbuf[warp];
    const unsigned FULL = 0xffffffffu;
    for (int w = lane; w < RENC_WORDS; w += 32) buf[w] = 0u;       // kept clear from chunk to chunk
    __syncwarp();
    for (int tile = blockIdx.x * RENC_WARPS + warp; tile < ntiles; tile += gridDim.x * RENC_WARPS) {
        if (!src.open(tile, lane)) {
            if (lane == 0) out_lens[tile] = 0;
            continue;
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(scratch + (size_t)tile * stride);
        int wpos = 0;                       // whole 32-bit words already written for this tile
        int cbits = 0;                      // bits of the partly filled word, which sits in buf[0]
        int lastpix = 0;
        for (int i0 = 0; i0 < nx; i0 += RENC_CHUNK) {
            // ---- the chunk's pixel values, block k in vals[33 k ..]
            if (i0 == 0) lastpix = src.px(0);            // the first pixel: block 0 starts with difference 0
            const bool differs = src.fill(vals, i0, lane, lastpix);
            if (i0 == 0) {
                // the first pixel of the tile as it is
                if (BP == 4) {
                    if (lane == 0) dst[0] = __byte_perm((uint32_t)lastpix, 0, 0x0123);
                    wpos = 1;
                } else {
                    if (lane == 0) buf[0] = (uint32_t)lastpix << (32 - P::BBITS);
                    cbits = P::BBITS;
                }
            }
            __syncwarp();
            int total;
            if (!__any_sync(FULL, differs)) {
                // a chunk of one value (most of a mask): every block is the code 0, nothing to write
                const int nblocks = min(32, (nx - i0 + RICE_BLOCK - 1) / RICE_BLOCK);
                total = cbits + nblocks * P::FSBITS;
            } else {
                // ---- lane = block; the differences replace the values in shared memory
                const int nthis = max(0, min(RICE_BLOCK, nx - (i0 + RICE_BLOCK * lane)));
                int prev = (lane == 0) ? lastpix : vals[(lane - 1) * 33 + 31];
                lastpix = vals[31 * 33 + 31];   // (of use only if another chunk follows: then the chunk was full)
                __syncwarp();
                int *mine = vals + lane * 33;
                unsigned long long sum = 0;
#pragma unroll
                for (int j = 0; j < RICE_BLOCK; j++) {
                    const int cur = mine[j];
                    // difference in the pixel's own width (it wraps), zig-zag mapped
                    int pd = (int)((unsigned)cur - (unsigned)prev);
                    if (BP == 1) pd = (int)(int8_t)pd;
                    if (BP == 2) pd = (int)(int16_t)pd;
                    uint32_t d = (pd < 0) ? ~((uint32_t)pd << 1) : ((uint32_t)pd << 1);
                    if (j >= nthis) d = 0;
                    mine[j] = (int)d;
                    sum += d;
                    prev = cur;
                }
                int fs = 0, nbits = 0;
                bool raw = false;
                if (nthis > 0) {
                    nbits = P::FSBITS;          // all differences zero: the code 0 and nothing else
                    if (sum != 0) {
                        // FS from the block sum, as fits_rcomp computes it in double
                        // (a full block divides by 32: the same double, by a multiplication)
                        double dpsum = (double)sum - (double)(nthis / 2) - 1.0;
                        dpsum = nthis == RICE_BLOCK ? dpsum * (1.0 / RICE_BLOCK) : dpsum / (double)nthis;
                        if (dpsum < 0) dpsum = 0.0;
                        const unsigned long long ip = (unsigned long long)dpsum;
                        const uint32_t psum = (BP == 1 ? (uint32_t)(uint8_t)ip : BP == 2 ? (uint32_t)(uint16_t)ip : (uint32_t)ip) >> 1;
                        fs = 32 - __clz((int)psum);                  // number of bits of psum
                        raw = fs >= P::FSMAX;
                        if (raw) {
                            nbits += nthis * P::BBITS;
                        } else {
                            uint32_t tops = 0;
#pragma unroll
                            for (int j = 0; j < RICE_BLOCK; j++) tops += (uint32_t)mine[j] >> fs;
                            nbits += (int)tops + nthis * (fs + 1);
                        }
                    }
                }
                // ---- bit position of every block: the carried bits, then the blocks in order
                int incl = nbits;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                total = cbits + __shfl_sync(FULL, incl, 31);
                int pos = cbits + incl - nbits;
                if (nthis > 0 && sum != 0) {
                    BitOut out;
                    out.open(buf, pos);
                    out.put(raw ? (uint32_t)(P::FSMAX + 1) : (uint32_t)(fs + 1), P::FSBITS);
                    if (raw) {
#pragma unroll 8
                        for (int j = 0; j < nthis; j++) {
                            const uint32_t d = (uint32_t)mine[j];
                            out.put(BP == 4 ? d : (d & ((1u << (P::BBITS & 31)) - 1u)), P::BBITS);
                        }
                    } else {
                        // a code = (diff >> FS) zeros, a one, the low FS bits
                        const uint32_t low = (1u << fs) - 1u;
#pragma unroll 8
                        for (int j = 0; j < nthis; j++) {
                            const uint32_t d = (uint32_t)mine[j];
                            out.zeros((int)(d >> fs));
                            out.put((1u << fs) | (d & low), fs + 1);
                        }
                    }
                    out.close();
                }
                __syncwarp();
            }
            // ---- whole words out, the rest stays in front of the next chunk
            const int nfull = total >> 5;
            for (int w = lane; w < nfull; w += 32) dst[wpos + w] = __byte_perm(buf[w], 0, 0x0123);
            const uint32_t carry = buf[nfull];
            __syncwarp();
            for (int w = lane; w <= nfull + 1; w += 32) buf[w] = 0u;
            __syncwarp();
            if (lane == 0) buf[0] = carry;
            wpos += nfull;
            cbits = total & 31;
            __syncwarp();
        }
        // flush the partly filled word: the stream ends on a byte boundary, zero padded
        const int tail_bytes = (cbits + 7) >> 3;
        if (lane == 0) {
            const uint32_t carry = buf[0];
            uint8_t *tb = reinterpret_cast<uint8_t *>(dst + wpos);
            for (int b = 0; b < tail_bytes; b++) tb[b] = (uint8_t)(carry >> (24 - 8 * b));
            out_lens[tile] = 4 * wpos + tail_bytes;
            buf[0] = 0u;
        }
        __syncwarp();
    }
}

// exclusive prefix sum of the tile sizes (one CTA; ntiles is ~10^4): offs[t], hdr = {total bytes,
// ntiles, status}
struct RiceOutHdr { long long total; int ntiles; int status; };

__global__ void __launch_bounds__(1024)
rice_scan_kernel(const int *__restrict__ lens, int ntiles, long long *__restrict__ offs, RiceOutHdr *hdr,
                 long long heap_cap, const int *__restrict__ nskipped)
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
        hdr->status = (part[1023] > heap_cap ? 1 : 0) | (nskipped ? (*nskipped << 8) : 0);
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
    if (bytepix == 1) {
        PlainSrc<1> src = {(const uint8_t *)img, nx, nullptr};
        rice_encode_kernel<1><<<blocks, RENC_WARPS * 32, 0, st>>>(src, ntiles, nx, scratch, stride, lens);
    } else if (bytepix == 2) {
        PlainSrc<2> src = {(const uint16_t *)img, nx, nullptr};
        rice_encode_kernel<2><<<blocks, RENC_WARPS * 32, 0, st>>>(src, ntiles, nx, scratch, stride, lens);
    } else {
        PlainSrc<4> src = {(const uint32_t *)img, nx, nullptr};
        rice_encode_kernel<4><<<blocks, RENC_WARPS * 32, 0, st>>>(src, ntiles, nx, scratch, stride, lens);
    }
    BBX_CHECK_LAUNCH("rice_encode_kernel");
    rice_scan_kernel<<<1, 1024, 0, st>>>(lens, ntiles, offs, hdr, heap_cap, nullptr);
    BBX_CHECK_LAUNCH("rice_scan_kernel");
    const int cblocks = (ntiles + 7) / 8 < BBX_SM_COUNT * 8 ? (ntiles + 7) / 8 : BBX_SM_COUNT * 8;
    rice_compact_kernel<<<cblocks, 256, 0, st>>>(scratch, stride, lens, offs, ntiles, heap, heap_cap);
    BBX_CHECK_LAUNCH("rice_compact_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// `fpack -q <q>` of a float32 image on the device (blackbox.py:826-836: every reduced image is
// written as `fpack -q 16 -D -Y`: quantised with subtractive dither, Rice-coded) -- so that the
// compressed product, a fifth of the float32 image, is what crosses PCIe on the way out.
//
// Per row tile (fits_quantize_float + FnNoise5_float of CFITSIO's quantize.c, restated in
// oracle/rice.py: fn_noise5_row / fpack_quantize_row):
//   c = 4 .. nx-5, v1..v9 = row[c-4 .. c+4], float32 arithmetic left to right
//     d2 = |v5 - v7|                       unless v5 == v6 == v7
//     d3 = |2 v5 - v3 - v7|                unless v3 == v4 == v5 == v6 == v7
//     d5 = |6 v5 - 4 v3 - 4 v7 + v1 + v9|  (as d3)
//   med = element (m - 1) / 2 of the m = count(d3) sorted values (d2: over m entries of a
//   zero-filled array holding count(d2) <= m values, and only if count(d2) > 1, or == 1 when m == 1)
//   sigma = 0.6052697 med3, replaced by 1.0483579 med2 / 0.1772048 med5 where non-zero and smaller
//   ZSCALE = sigma / q, ZZERO = trunc(min / ZSCALE + 0.5) ZSCALE  (mid-range if the span needs > 31 bits)
//   ZSCALE = 0 marks a row that is not quantised: sigma == 0, span > 32 bits, or a non-finite value.
// fq_row_stats_kernel: one CTA per row, the row in shared memory, the differences recomputed in
// every pass.  The three medians are exact selections: a linear histogram scaled by a 32-value
// sample finds the bin of the wanted rank, the handful of values in it are ranked directly;
// rows where that does not work (hundreds of tied values, a misleading sample) take a radix
// select on the float bits (non-negative: the bits order like the values), 11 + 11 + 9 bits.
// The quantised values are made inside the Rice encoder (QuantSrc).
// ---------------------------------------------------------------------------------------------
#define FQ_THREADS 256
#define FQ_BINS 2048
#define FQ_MAX_NX 16384
#define FQ_CAND 128                     // candidates per array the fast path ranks directly

struct FqSel { unsigned int prefix; unsigned int rank; int use; int bin; unsigned int in_bin; float scale; };

// the three differences around pixel c (v1..v9 = row[c-4 .. c+4]) and whether they count
__device__ __forceinline__ void fq_diffs(const float *srow, int c, bool &keep2, bool &keep3, float &d2, float &d3, float &d5)
{
    const float v1 = srow[c - 4], v3 = srow[c - 2], v4 = srow[c - 1], v5 = srow[c], v6 = srow[c + 1], v7 = srow[c + 2],
                v9 = srow[c + 4];
    const bool e567 = (v5 == v6) && (v6 == v7);
    keep2 = !e567;
    keep3 = !((v3 == v4) && (v4 == v5) && e567);
    d2 = fabsf(v5 - v7);
    float a = 2.0f * v5;
    a = a - v3;
    a = a - v7;
    float b = 6.0f * v5;
    b = b - 4.0f * v3;
    b = b - 4.0f * v7;
    b = b + v1;
    b = b + v9;
    d3 = fabsf(a);
    d5 = fabsf(b);
}

// monotone in d (d >= 0 or NaN): the bin of the linear histogram
__device__ __forceinline__ int fq_linear_bin(float d, float scale)
{
    const float x = d * scale;
    return x < (float)(FQ_BINS - 1) ? (int)x : FQ_BINS - 1;       // NaN -> last bin
}

// Block-wide: the bin of hist[0..FQ_BINS) that holds rank s.rank (if s.use); s.bin = that bin,
// s.in_bin = its count, s.rank becomes the rank inside the bin.
__device__ __forceinline__ void fq_find_bin(const unsigned int *hist, FqSel &s, unsigned int *s_scan)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = FQ_BINS / FQ_THREADS;
    unsigned int mine = 0;
#pragma unroll
    for (int i = 0; i < per; i++) mine += hist[tid * per + i];
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    unsigned int before = inc - mine;
    for (int w = 0; w < warp; w++) before += s_scan[w];
    const unsigned int k = s.rank;
    const int use = s.use;
    __syncthreads();                                    // everybody has read the rank
    if (use && k >= before && k < before + mine) {
        unsigned int acc = before;
        int b = tid * per;
        for (; b < tid * per + per - 1; b++) { if (acc + hist[b] > k) break; acc += hist[b]; }
        s.bin = b;
        s.in_bin = hist[b];
        s.rank = k - acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FQ_THREADS)
fq_row_stats_kernel(const float *__restrict__ img, int ntiles, int nx, float qlevel, double *__restrict__ zscale,
                    double *__restrict__ zzero, int *__restrict__ nskipped)
{
    extern __shared__ __align__(16) unsigned char fq_smem[];
    unsigned int (*hist)[FQ_BINS] = reinterpret_cast<unsigned int (*)[FQ_BINS]>(fq_smem);
    float *srow = reinterpret_cast<float *>(fq_smem + 3 * FQ_BINS * sizeof(unsigned int));
    __shared__ float s_lo[FQ_THREADS / 32], s_hi[FQ_THREADS / 32];
    __shared__ unsigned int s_cnt[2][FQ_THREADS / 32];
    __shared__ unsigned int s_scan[FQ_THREADS / 32];
    __shared__ FqSel sel[3];
    __shared__ unsigned int cand[3][FQ_CAND];
    __shared__ unsigned int ncand[3];
    __shared__ int s_bad, s_slow;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;

    for (int i = tid; i < 3 * FQ_BINS; i += FQ_THREADS) (&hist[0][0])[i] = 0u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const float *row = img + (size_t)tile * nx;
        // ---- the row, its range, non-finite values
        float lo = INFINITY, hi = -INFINITY;
        int bad = 0;
        if (tid == 0) { s_bad = 0; s_slow = 0; }
        if (tid < 3) ncand[tid] = 0;
        if ((nx & 3) == 0 && ((uintptr_t)row & 15) == 0) {
            const float4 *r4 = reinterpret_cast<const float4 *>(row);
            for (int i = tid; i < nx / 4; i += FQ_THREADS) {
                const float4 v = __ldcs(r4 + i);
                reinterpret_cast<float4 *>(srow)[i] = v;
                lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
                hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
                bad |= !isfinite(v.x) | !isfinite(v.y) | !isfinite(v.z) | !isfinite(v.w);
            }
        } else {
            for (int i = tid; i < nx; i += FQ_THREADS) {
                const float v = row[i];
                srow[i] = v;
                lo = fminf(lo, v); hi = fmaxf(hi, v);
                bad |= !isfinite(v);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(FULL, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o));
        }
        __syncthreads();                                   // flags reset, row stored
        if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
        if (bad) s_bad = 1;
        __syncthreads();
        lo = s_lo[0]; hi = s_hi[0];
#pragma unroll
        for (int w = 1; w < FQ_THREADS / 32; w++) { lo = fminf(lo, s_lo[w]); hi = fmaxf(hi, s_hi[w]); }
        const bool unusable = s_bad != 0 || nx < 9;
        const int nd = nx - 8;                             // centres c = 4 .. nx - 5

        if (!unusable) {
            // ---- a first look: 32 differences spread over the row give the scale of a LINEAR
            // histogram (its bins hold a handful of values each; the float bits of a noise-like
            // quantity would pile thousands of values onto a few exponent bins and the
            // shared-memory atomics would queue up behind each other)
            if (warp == 0) {
                const int c = 4 + (int)(((long long)(2 * lane + 1) * nd) / 64);
                bool k2, k3;
                float d[3];
                fq_diffs(srow, c, k2, k3, d[0], d[1], d[2]);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    int rank = 0;
                    for (int j = 0; j < 32; j++) {
                        const float o = __shfl_sync(FULL, d[a], j);
                        rank += (o < d[a]) || (o == d[a] && j < lane);
                    }
                    const unsigned int who = __ballot_sync(FULL, rank == 16);
                    const float med = __shfl_sync(FULL, d[a], who ? __ffs(who) - 1 : 0);
                    // four sample medians span the histogram; no usable sample -> the slow path
                    const float sc = (float)(FQ_BINS - 1) / (4.0f * med);
                    if (lane == 0) sel[a].scale = (who && med > 0.f && isfinite(sc)) ? sc : 0.f;
                }
            }
            __syncthreads();
            const float sc2 = sel[0].scale, sc3 = sel[1].scale, sc5 = sel[2].scale;
            unsigned int n2 = 0, n3 = 0;
            for (int c = 4 + tid; c < nx - 4; c += FQ_THREADS) {
                bool k2, k3;
                float d2, d3, d5;
                fq_diffs(srow, c, k2, k3, d2, d3, d5);
                if (k2) { n2++; atomicAdd(&hist[0][fq_linear_bin(d2, sc2)], 1u); }
                if (k3) {
                    n3++;
                    atomicAdd(&hist[1][fq_linear_bin(d3, sc3)], 1u);
                    atomicAdd(&hist[2][fq_linear_bin(d5, sc5)], 1u);
                }
            }
            n2 = __reduce_add_sync(FULL, n2);
            n3 = __reduce_add_sync(FULL, n3);
            if (lane == 0) { s_cnt[0][warp] = n2; s_cnt[1][warp] = n3; }
            __syncthreads();
            if (tid == 0) {
                // the ranks asked for: m = count(d3); d2 sits in a zero-filled array of m entries
                unsigned int m2 = 0, m = 0;
                for (int w = 0; w < FQ_THREADS / 32; w++) { m2 += s_cnt[0][w]; m += s_cnt[1][w]; }
                const unsigned int r = m ? (m - 1) / 2 : 0, nz = m - m2;
                sel[1].rank = r; sel[1].use = m > 0;
                sel[2].rank = r; sel[2].use = m > 0;
                sel[0].use = (m == 1 ? m2 == 1 : m2 > 1) && r >= nz;
                sel[0].rank = r >= nz ? r - nz : 0;
                s_cnt[0][0] = sel[0].rank; s_cnt[0][1] = r;          // kept for the slow path
            }
            __syncthreads();
            for (int a = 0; a < 3; a++) fq_find_bin(hist[a], sel[a], s_scan);
            for (int i = tid; i < 3 * FQ_BINS; i += FQ_THREADS) (&hist[0][0])[i] = 0u;
            if (tid < 3 && sel[tid].use &&
                (sel[tid].scale == 0.f || sel[tid].bin == FQ_BINS - 1 || sel[tid].in_bin > FQ_CAND)) s_slow = 1;
            __syncthreads();
            if (!s_slow) {
                // ---- the few values of the three bins, ranked directly
                const int b2 = sel[0].use ? sel[0].bin : -1, b3 = sel[1].use ? sel[1].bin : -1,
                          b5 = sel[2].use ? sel[2].bin : -1;
                for (int c = 4 + tid; c < nx - 4; c += FQ_THREADS) {
                    bool k2, k3;
                    float d2, d3, d5;
                    fq_diffs(srow, c, k2, k3, d2, d3, d5);
                    if (k2 && fq_linear_bin(d2, sc2) == b2) cand[0][atomicAdd(&ncand[0], 1u)] = __float_as_uint(d2);
                    if (k3 && fq_linear_bin(d3, sc3) == b3) cand[1][atomicAdd(&ncand[1], 1u)] = __float_as_uint(d3);
                    if (k3 && fq_linear_bin(d5, sc5) == b5) cand[2][atomicAdd(&ncand[2], 1u)] = __float_as_uint(d5);
                }
                __syncthreads();
                for (int a = 0; a < 3; a++) {
                    const unsigned int n = ncand[a];
                    if (!sel[a].use || (unsigned int)tid >= n) continue;
                    const unsigned int mine = cand[a][tid];
                    unsigned int rank = 0;
                    for (unsigned int j = 0; j < n; j++) {
                        const unsigned int o = cand[a][j];
                        rank += (o < mine) || (o == mine && j < (unsigned int)tid);
                    }
                    if (rank == sel[a].rank) sel[a].prefix = mine;
                }
                __syncthreads();
            } else {
                // ---- the slow path (ties by the hundred, or a sample that misled): exact radix
                // select on the float bits, 11 + 11 + 9 bits, the differences recomputed each pass
                if (tid == 0) {
                    sel[0].rank = s_cnt[0][0]; sel[1].rank = s_cnt[0][1]; sel[2].rank = s_cnt[0][1];
                    sel[0].prefix = sel[1].prefix = sel[2].prefix = 0;
                }
                __syncthreads();
                for (int level = 0; level < 3; level++) {
                    const int shift = level == 0 ? 20 : level == 1 ? 9 : 0;
                    const int pshift = level == 1 ? 20 : 9;        // bits fixed so far = key >> pshift
                    const unsigned int mask = level == 2 ? 0x1ffu : 0x7ffu;
                    const unsigned int p2 = sel[0].prefix, p3 = sel[1].prefix, p5 = sel[2].prefix;
                    for (int c = 4 + tid; c < nx - 4; c += FQ_THREADS) {
                        bool k2, k3;
                        float d2, d3, d5;
                        fq_diffs(srow, c, k2, k3, d2, d3, d5);
                        const unsigned int q2 = __float_as_uint(d2), q3 = __float_as_uint(d3), q5 = __float_as_uint(d5);
                        if (k2 && (level == 0 || (q2 >> pshift) == p2)) atomicAdd(&hist[0][(q2 >> shift) & mask], 1u);
                        if (k3 && (level == 0 || (q3 >> pshift) == p3)) atomicAdd(&hist[1][(q3 >> shift) & mask], 1u);
                        if (k3 && (level == 0 || (q5 >> pshift) == p5)) atomicAdd(&hist[2][(q5 >> shift) & mask], 1u);
                    }
                    __syncthreads();
                    for (int a = 0; a < 3; a++) {
                        fq_find_bin(hist[a], sel[a], s_scan);
                        if (tid == 0 && sel[a].use) sel[a].prefix = (sel[a].prefix << (level == 2 ? 9 : 11)) | (unsigned int)sel[a].bin;
                    }
                    for (int i = tid; i < 3 * FQ_BINS; i += FQ_THREADS) (&hist[0][0])[i] = 0u;
                    __syncthreads();
                }
            }
        }
        if (tid == 0) {
            double delta = 0.0, zero = 0.0;
            if (!unusable && sel[1].use) {
                const double n3v = 0.6052697 * (double)__uint_as_float(sel[1].prefix);
                const double n5v = 0.1772048 * (double)__uint_as_float(sel[2].prefix);
                const double n2v = sel[0].use ? 1.0483579 * (double)__uint_as_float(sel[0].prefix) : 0.0;
                double sigma = n3v;
                if (n2v != 0.0 && n2v < sigma) sigma = n2v;
                if (n5v != 0.0 && n5v < sigma) sigma = n5v;
                delta = sigma / (double)qlevel;
                if (delta != 0.0) {
                    const double span = ((double)hi - (double)lo) / delta;
                    if (!(span <= 2.0 * 2147483647.0 - 10.0)) delta = 0.0;       // (also a NaN / infinite sigma)
                    else if (span < 2147483647.0 - 10.0) zero = trunc((double)lo / delta + 0.5) * delta;
                    else zero = ((double)lo + (double)hi) / 2.0;
                }
            }
            zscale[tile] = delta;
            zzero[tile] = zero;
            if (delta == 0.0) atomicAdd(nskipped, 1);
        }
        __syncthreads();
    }
}

static size_t fq_lens_bytes(int ntiles) { return ((size_t)ntiles * 4 + 15) / 16 * 16; }
static size_t fq_col_bytes(int ntiles) { return ((size_t)ntiles * 8 + 15) / 16 * 16; }
static size_t fq_heap_offset(int ntiles) { return 16 + fq_lens_bytes(ntiles) + 2 * fq_col_bytes(ntiles); }

// bbx_fpack_f32: img (device float32, ntiles rows of nx pixels) -> out (device):
//     [0:8)  int64 total heap bytes    [8:12) int32 ntiles
//     [12:16) int32 status: bit 0 the heap did not fit into out_bytes; bits 8.. = number of rows NOT
//             quantised (ZSCALE 0, no Rice-coded bytes: the caller stores them losslessly)
//     int32 compressed bytes per tile (padded to 16 B), float64 ZSCALE per tile (padded to 16 B),
//     float64 ZZERO per tile (padded to 16 B), then the heap (bbx_fpack_f32_heap_offset).
// qlevel: fpack's -q (16 for the reduced images); zdither0: the ZDITHER0 keyword to write (1..10000;
// fpack takes it from the clock); rand10000: the FITS standard's random table on the device.
extern "C" size_t bbx_fpack_f32_work_bytes(int ntiles, int nx) { return bbx_rice_encode_work_bytes(ntiles, nx, 4); }

extern "C" size_t bbx_fpack_f32_heap_offset(int ntiles) { return ntiles > 0 ? fq_heap_offset(ntiles) : 0; }

extern "C" size_t bbx_fpack_f32_out_bytes(int ntiles, int nx)
{
    if (ntiles <= 0 || nx <= 0) return 0;
    return fq_heap_offset(ntiles) + (size_t)ntiles * rice_stride(nx, 4);
}

extern "C" int bbx_fpack_f32(const float *img, int ntiles, int nx, float qlevel, int zdither0, const float *rand10000,
                             void *work, size_t work_bytes, void *out, size_t out_bytes, void *stream)
{
    BBX_REQUIRE(img && work && out && rand10000, "bbx_fpack_f32: null argument");
    BBX_REQUIRE(ntiles > 0 && nx > 0 && nx <= FQ_MAX_NX, "bbx_fpack_f32: %d tiles of %d pixels (rows of up to %d)", ntiles, nx, FQ_MAX_NX);
    BBX_REQUIRE(qlevel > 0.f, "bbx_fpack_f32: quantisation level %g", (double)qlevel);
    BBX_REQUIRE(zdither0 >= 1 && zdither0 <= RICE_NRANDOM, "bbx_fpack_f32: ZDITHER0 %d", zdither0);
    BBX_REQUIRE(work_bytes >= bbx_fpack_f32_work_bytes(ntiles, nx), "bbx_fpack_f32: work buffer too small");
    BBX_REQUIRE(out_bytes >= fq_heap_offset(ntiles), "bbx_fpack_f32: output buffer too small for the table columns");
    BBX_REQUIRE(((uintptr_t)work % 16) == 0 && ((uintptr_t)out % 16) == 0, "bbx_fpack_f32: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = rice_stride(nx, 4);
    uint8_t *scratch = (uint8_t *)work;
    long long *offs = reinterpret_cast<long long *>(scratch + ((size_t)ntiles * stride + 15) / 16 * 16);
    RiceOutHdr *hdr = (RiceOutHdr *)out;
    int *lens = reinterpret_cast<int *>((uint8_t *)out + 16);
    double *zscale = reinterpret_cast<double *>((uint8_t *)out + 16 + fq_lens_bytes(ntiles));
    double *zzero = reinterpret_cast<double *>((uint8_t *)zscale + fq_col_bytes(ntiles));
    uint8_t *heap = (uint8_t *)out + fq_heap_offset(ntiles);
    const long long heap_cap = (long long)(out_bytes - fq_heap_offset(ntiles));
    int *nskipped = reinterpret_cast<int *>(offs + ntiles);               // the work buffer's last 16 bytes
    BBX_CUDA(cudaMemsetAsync(nskipped, 0, sizeof(int), st));
    const size_t smem = 3 * FQ_BINS * sizeof(unsigned int) + ((size_t)nx * 4 + 15) / 16 * 16;
    static bool attr_set = false;
    if (!attr_set) {
        BBX_CUDA(cudaFuncSetAttribute(fq_row_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * FQ_BINS * 4 + FQ_MAX_NX * 4));
        attr_set = true;
    }
    const int sblocks = ntiles < BBX_SM_COUNT * 3 ? ntiles : BBX_SM_COUNT * 3;
    fq_row_stats_kernel<<<sblocks, FQ_THREADS, smem, st>>>(img, ntiles, nx, qlevel, zscale, zzero, nskipped);
    BBX_CHECK_LAUNCH("fq_row_stats_kernel");
    const int want = (ntiles + RENC_WARPS - 1) / RENC_WARPS;
    const int blocks = want < BBX_SM_COUNT * 16 ? want : BBX_SM_COUNT * 16;
    QuantSrc src = {img, zscale, zzero, rand10000, nx, zdither0, nullptr, 0.0, 0.0, 0.f, 0.f, 0.f, 0, 0, 0};
    rice_encode_kernel<4><<<blocks, RENC_WARPS * 32, 0, st>>>(src, ntiles, nx, scratch, stride, lens);
    BBX_CHECK_LAUNCH("rice_encode_kernel(quantised)");
    rice_scan_kernel<<<1, 1024, 0, st>>>(lens, ntiles, offs, hdr, heap_cap, nskipped);
    BBX_CHECK_LAUNCH("rice_scan_kernel");
    const int cblocks = (ntiles + 7) / 8 < BBX_SM_COUNT * 8 ? (ntiles + 7) / 8 : BBX_SM_COUNT * 8;
    rice_compact_kernel<<<cblocks, 256, 0, st>>>(scratch, stride, lens, offs, ntiles, heap, heap_cap);
    BBX_CHECK_LAUNCH("rice_compact_kernel");
    return 0;
}
