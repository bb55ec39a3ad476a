"""Rice coding of 8 / 16 / 32-bit image tiles and float quantisation -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference reads its raw frames through zogy's read_hdulist (blackbox.py:1451; copy at
blackbox_slurm_google.py:1144-1253), i.e. astropy.io.fits on fpacked files, which hands the tiles
to CFITSIO's ricecomp.c.  Neither astropy nor CFITSIO is in /root/reference or in this image
(third-party, versions unpinned), so this module restates the published algorithm
(R. White's Rice coder as distributed with CFITSIO, functions fits_rcomp_short /
fits_rdecomp_short; FITS tiled-image convention, Pence, White & Seaman 2010):

  tile  = first pixel, 8 / 16 / 32 bits (BYTEPIX 1 / 2 / 4), big-endian; then blocks of ``nblock`` = 32 pixels
  block = 3 / 4 / 5 bits FS+1 (0: every difference is 0; FSMAX+1 = 7 / 15 / 26: differences as raw
          bits), then per pixel (diff >> FS) zero bits, a one bit, the low FS bits of diff
  diff  = zig-zag mapped difference to the previous pixel in the pixel's own width
  FS    = number of bits of  ((sum(diff) - nblock/2 - 1) / nblock) >> 1   (the encoder's choice;
          any FS decodes)

PARITY UNPINNED: no fpack / CFITSIO / astropy here to produce or read a real .fz file; the two
hand-derived known-answer vectors in tests/test_zz_rice_fz.py follow from the format text above.
"""
import numpy as np

# BYTEPIX -> (FSBITS, FSMAX, BBITS) of fits_rcomp_byte / fits_rcomp_short / fits_rcomp
PARAMS = {1: (3, 6, 8), 2: (4, 14, 16), 4: (5, 25, 32)}
FSBITS, FSMAX, BBITS = PARAMS[2]


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value, nbits):
        if nbits == 0:
            return
        self.acc = (self.acc << nbits) | (int(value) & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xff)
        self.acc &= (1 << self.n) - 1

    def done(self):
        if self.n:
            self.out.append((self.acc << (8 - self.n)) & 0xff)
            self.acc = self.n = 0
        return bytes(self.out)


def encode_tile(a, bytepix=2, nblock=32, fast=None):
    """Stored (signed char / short / int) pixel values of one tile -> bytes
    (fits_rcomp_byte / fits_rcomp_short / fits_rcomp).  ``fast``: use the C copy of the same
    statements (oracle/csrc/bbo.c; default for rows of 256 pixels and more -- the tests hold the
    two against each other)."""
    if (fast or (fast is None and np.size(a) >= 256)) and nblock == 32:
        from . import clib
        return clib.rice_encode(_stored(a, bytepix), bytepix)
    fsbits, fsmax, bbits = PARAMS[bytepix]
    half, full = 1 << (bbits - 1), 1 << bbits
    a = ((np.asarray(a).astype(np.int64) + half) % full) - half          # the signed view of the stored bits
    w = _BitWriter()
    w.put(int(a[0]) & (full - 1), bbits)
    lastpix = int(a[0])
    nx = a.size
    for i in range(0, nx, nblock):
        blk = a[i:i + nblock]
        thisblock = blk.size
        prev = np.concatenate(([lastpix], blk[:-1]))
        pdiff = ((blk - prev + half) % full) - half                      # arithmetic in the pixel's own width
        diff = np.where(pdiff < 0, ~(pdiff << 1), pdiff << 1) & 0xffffffff
        lastpix = int(blk[-1])
        pixelsum = float(diff.sum())
        dpsum = (pixelsum - (thisblock // 2) - 1) / thisblock
        if dpsum < 0:
            dpsum = 0.0
        psum = (int(dpsum) & (full - 1)) >> 1                            # (unsigned char / short / int) dpsum
        fs = 0
        while psum > 0:
            psum >>= 1
            fs += 1
        if fs >= fsmax:
            w.put(fsmax + 1, fsbits)
            for v in diff:
                w.put(int(v), bbits)
        elif fs == 0 and pixelsum == 0:
            w.put(0, fsbits)
        else:
            w.put(fs + 1, fsbits)
            for v in diff:
                v = int(v)
                top = v >> fs
                while top >= 24:                                   # long runs of zeros in pieces
                    w.put(0, 24)
                    top -= 24
                w.put(1, top + 1)
                w.put(v & ((1 << fs) - 1), fs)
    return w.done()


def encode_tile16(a, nblock=32):
    """int16 (stored) pixel values of one tile -> bytes (fits_rcomp_short)."""
    return encode_tile(a, 2, nblock)


def _stored(a, bytepix):
    """any integer array -> the int8 / int16 / int32 bit pattern of its low bytes"""
    full = 1 << (8 * bytepix)
    half = full >> 1
    v = ((np.asarray(a).astype(np.int64) + half) % full) - half
    return v.astype({1: np.int8, 2: np.int16, 4: np.int32}[bytepix])


def decode_tile(buf, nx, bytepix=2, nblock=32, fast=None):
    """bytes -> nx stored pixel values as int64 in [0, 2^bbits) (fits_rdecomp_byte / _short /
    fits_rdecomp, statement by statement).  ``fast``: the C copy, as for ``encode_tile``."""
    if (fast or (fast is None and nx >= 256)) and nblock == 32:
        from . import clib
        return clib.rice_decode(buf, nx, bytepix).astype(np.int64)
    fsbits, fsmax, bbits = PARAMS[bytepix]
    vmask = (1 << bbits) - 1
    c = memoryview(bytes(buf))
    out = np.zeros(nx, dtype=np.int64)
    lastpix = 0
    for k in range(bytepix):
        lastpix = (lastpix << 8) | c[k]
    pos = bytepix
    b = c[pos]
    pos += 1
    nbits = 8
    i = 0
    while i < nx:
        nbits -= fsbits
        while nbits < 0:
            b = (b << 8) | c[pos]
            pos += 1
            nbits += 8
        fs = (b >> nbits) - 1
        b &= (1 << nbits) - 1
        imax = min(i + nblock, nx)
        if fs < 0:
            out[i:imax] = lastpix
            i = imax
        elif fs == fsmax:
            while i < imax:
                k = bbits - nbits
                diff = b << k
                k -= 8
                while k >= 0:
                    b = c[pos]
                    pos += 1
                    diff |= b << k
                    k -= 8
                if nbits > 0:
                    b = c[pos]
                    pos += 1
                    diff |= b >> (-k)
                    b &= (1 << nbits) - 1
                else:
                    b = 0
                diff &= vmask
                diff = ~(diff >> 1) if diff & 1 else diff >> 1
                lastpix = (lastpix + diff) & vmask
                out[i] = lastpix
                i += 1
        else:
            while i < imax:
                while b == 0:
                    nbits += 8
                    b = c[pos]
                    pos += 1
                nzero = nbits - b.bit_length()
                nbits -= nzero + 1
                b ^= 1 << nbits
                nbits -= fs
                while nbits < 0:
                    b = (b << 8) | c[pos]
                    pos += 1
                    nbits += 8
                diff = (nzero << fs) | (b >> nbits)
                b &= (1 << nbits) - 1
                diff = ~(diff >> 1) if diff & 1 else diff >> 1
                lastpix = (lastpix + diff) & vmask
                out[i] = lastpix
                i += 1
    return out


def decode_tile16(buf, nx, nblock=32):
    """bytes -> int16 array of nx pixels (fits_rdecomp_short)."""
    return decode_tile(buf, nx, 2, nblock).astype(np.uint16).view(np.int16)


# -------------------------------------------------------------------------------------------
# float images: quantisation with subtractive dithering (FITS standard 4.0, section 10.2 and
# the appendix that defines the random number sequence; Pence, White & Seaman 2010)
# -------------------------------------------------------------------------------------------
N_RANDOM = 10000
NULL_VALUE, ZERO_VALUE = -2147483647, -2147483646
_rand = None


def random_table():
    """The 10000 published random numbers: Park & Miller's minimal standard generator
    (a = 16807, m = 2^31 - 1, seed 1), value = seed / m as float32; the 10000th seed must be
    1043618065 (the standard's own check value)."""
    global _rand
    if _rand is None:
        a, m = 16807.0, 2147483647.0
        seed = 1.0
        out = np.empty(N_RANDOM, dtype=np.float32)
        for i in range(N_RANDOM):
            temp = a * seed
            seed = temp - m * int(temp / m)
            out[i] = np.float32(seed / m)
        assert int(seed) == 1043618065
        _rand = out
    return _rand


def _dither_sequence(tile_index, zdither0, n):
    """The n random numbers pixel 0 .. n-1 of tile ``tile_index`` (0-based) are dithered with."""
    r = random_table()
    iseed = (tile_index + zdither0 - 1) % N_RANDOM
    nextrand = int(r[iseed] * np.float32(500))
    out = np.empty(n, dtype=np.float32)
    i = 0
    while i < n:                                     # one run of the table per pass, as fits_unquantize walks it
        m = min(n - i, N_RANDOM - nextrand)
        out[i:i + m] = r[nextrand:nextrand + m]
        i += m
        nextrand += m
        if nextrand == N_RANDOM:
            iseed = (iseed + 1) % N_RANDOM
            nextrand = int(r[iseed] * np.float32(500))
    return out


def quantize_tile(values, tile_index, scale, zero, dither=1, zdither0=1):
    """float32 row -> int32 (the encoder's side: q = nint((v - zero) / scale + R - 0.5);
    NaN -> NULL_VALUE; dither 2: exact zeros -> ZERO_VALUE)."""
    v = np.asarray(values, dtype=np.float64)
    if dither == 0:
        q = np.floor((v - zero) / scale + 0.5)
    else:
        r = _dither_sequence(tile_index, zdither0, v.size).astype(np.float64)
        q = np.floor((v - zero) / scale + r - 0.5 + 0.5)
    q = np.where(np.isnan(v), NULL_VALUE, q)
    if dither == 2:
        q = np.where(v == 0.0, ZERO_VALUE, q)
    return q.astype(np.int64).astype(np.int32)


def unquantize_tile(q, tile_index, scale, zero, dither=1, zdither0=1, zblank=NULL_VALUE):
    """int32 row -> float32, as fits_unquantize_i4r4 evaluates it:
    (float)(((double) q - R + 0.5) * scale + zero)."""
    q = np.asarray(q, dtype=np.int64)
    if dither == 0:
        out = (q.astype(np.float64) * scale + zero).astype(np.float32)
    else:
        r = _dither_sequence(tile_index, zdither0, q.size).astype(np.float64)
        out = ((q.astype(np.float64) - r + 0.5) * scale + zero).astype(np.float32)
        if dither == 2:
            out[q == ZERO_VALUE] = 0.0
    out[q == zblank] = np.nan
    return out


def _fz_file(path, tiles, shape, zbitpix, bytepix, header, pointer, lead_column, extra_cols=None, extra_cards=(),
             gzip_tiles=None):
    """Lay out primary HDU + BINTABLE + heap the way fpack does.  ``extra_cols``: list of
    (name, float64 array per tile) stored as 1D columns behind COMPRESSED_DATA."""
    from blackbox_b200 import fitsio
    H, W = shape
    gzip_tiles = gzip_tiles or {}
    stored = [gzip_tiles.get(r, t) for r, t in enumerate(tiles)]       # heap order: row by row
    slens = np.array([len(t) for t in stored], dtype=np.int64)
    offs = np.concatenate(([0], np.cumsum(slens)[:-1]))
    heap = b''.join(stored)
    isgz = np.array([r in gzip_tiles for r in range(H)])
    ptype = '>i4' if pointer == 'P' else '>i8'
    desc = np.stack([np.where(isgz, 0, slens), np.where(isgz, 0, offs)], axis=1).astype(ptype)
    lens = np.where(isgz, 0, slens)
    rows = [desc.view(np.uint8).reshape(H, -1)]
    names = ['COMPRESSED_DATA']
    forms = ['1{}B({})'.format(pointer, int(lens.max()))]
    if gzip_tiles:
        gdesc = np.stack([np.where(isgz, slens, 0), np.where(isgz, offs, 0)], axis=1).astype(ptype)
        rows.append(gdesc.view(np.uint8).reshape(H, -1))
        names.append('GZIP_COMPRESSED_DATA')
        forms.append('1{}B({})'.format(pointer, int(np.where(isgz, slens, 0).max())))
    if lead_column:
        rows.insert(0, np.zeros_like(rows[0]))
        names.insert(0, 'GZIP_COMPRESSED_DATA')
        forms.insert(0, forms[0])
    for name, vals in (extra_cols or []):
        rows.append(np.asarray(vals, dtype='>f8').reshape(H, 1).view(np.uint8).reshape(H, 8))
        names.append(name)
        forms.append('1D')
    table = np.concatenate(rows, axis=1)
    width = table.shape[1]
    card = fitsio._card
    primary = [card('SIMPLE', True), card('BITPIX', 16), card('NAXIS', 0), card('EXTEND', True), 'END'.ljust(80)]
    ext = [("XTENSION= 'BINTABLE'").ljust(80), card('BITPIX', 8), card('NAXIS', 2), card('NAXIS1', width),
           card('NAXIS2', H), card('PCOUNT', len(heap)), card('GCOUNT', 1), card('TFIELDS', len(names))]
    for n, (name, form) in enumerate(zip(names, forms), 1):
        ext += [card('TTYPE{}'.format(n), name), card('TFORM{}'.format(n), form)]
    ext += [card('ZIMAGE', True), card('ZSIMPLE', True), card('ZBITPIX', zbitpix), card('ZNAXIS', 2),
            card('ZNAXIS1', W), card('ZNAXIS2', H), card('ZTILE1', W), card('ZTILE2', 1), card('ZCMPTYPE', 'RICE_1'),
            card('ZNAME1', 'BLOCKSIZE'), card('ZVAL1', 32), card('ZNAME2', 'BYTEPIX'), card('ZVAL2', bytepix)]
    ext += list(extra_cards)
    for k, v in (header or {}).items():
        ext.append(card(k, v))
    ext.append('END'.ljust(80))
    with open(path, 'wb') as fh:
        for cards in (primary, ext):
            text = ''.join(cards)
            fh.write((text + ' ' * (-len(text) % fitsio.BLOCK)).encode('ascii'))
        body = table.tobytes() + heap
        fh.write(body + b'\0' * (-len(body) % fitsio.BLOCK))
    return path


def write_fz(path, counts, header=None, pointer='P', lead_column=False):
    """A tile-compressed FITS file of a uint16 image as fpack lays it out: empty primary HDU,
    BINTABLE with one COMPRESSED_DATA column of row tiles (BZERO 32768, RICE_1, BLOCKSIZE 32,
    BYTEPIX 2).  ``pointer``: 'P' (32-bit descriptors) or 'Q' (64-bit).  ``lead_column``: put an
    (empty) GZIP_COMPRESSED_DATA column in front, as CFITSIO does when it keeps a fall-back column."""
    from blackbox_b200 import fitsio
    counts = np.asarray(counts, dtype=np.uint16)
    stored = (counts.astype(np.int32) - 32768).astype(np.int16)
    tiles = [encode_tile(stored[r], 2) for r in range(counts.shape[0])]
    card = fitsio._card
    return _fz_file(path, tiles, counts.shape, 16, 2, header, pointer, lead_column,
                    extra_cards=[card('BSCALE', 1), card('BZERO', 32768)])


def write_fz_u8(path, img, header=None, pointer='P'):
    """uint8 image (the bad-pixel mask, the data mask): `fpack -D -Y` -> RICE_1, BYTEPIX 1."""
    img = np.asarray(img, dtype=np.uint8)
    tiles = [encode_tile(img[r], 1) for r in range(img.shape[0])]
    return _fz_file(path, tiles, img.shape, 8, 1, header, pointer, False)


# -------------------------------------------------------------------------------------------
# `fpack -q <q>` of a float image (blackbox.py:826-836 runs `fpack -q 16 -D -Y` on every reduced
# image): CFITSIO's fits_quantize_float with the noise estimate of FnNoise5_float (quantize.c of
# CFITSIO 4.x; not in /root/reference -- restated from the published source, PARITY UNPINNED:
# neither CFITSIO nor astropy is in this image.  fpack seeds the dither from the clock by
# default, so no two runs of the reference itself agree bit for bit; what a reader needs is in
# the file: ZSCALE / ZZERO per row and ZDITHER0).
#   row of n >= 9 pixels, c = 4 .. n-5, v1..v9 = row[c-4 .. c+4]:
#     d2 = |v5 - v7|                          unless v5 == v6 == v7
#     d3 = |2 v5 - v3 - v7|                   unless v3 == v4 == v5 == v6 == v7
#     d5 = |6 v5 - 4 v3 - 4 v7 + v1 + v9|     (same condition as d3)
#   (float32 arithmetic, left to right); median = element (m - 1) / 2 of the m = count(d3)
#   sorted values -- for d2 ALSO over m entries of its zero-initialised array (it holds
#   count(d2) <= m values: the rest are zeros), and only if count(d2) > 1;
#   noise2 = 1.0483579 med2, noise3 = 0.6052697 med3, noise5 = 0.1772048 med5;
#   sigma = noise3, replaced by noise2 / noise5 where those are non-zero and smaller.
#   delta = sigma / q; zero = trunc(min / delta + 0.5) * delta;
#   value = NINT((v - zero) / delta + R - 0.5), NINT(x) = trunc(x + 0.5) / trunc(x - 0.5) for x >= 0 / < 0.
# A row is NOT quantised (-> None; the writer stores it losslessly in GZIP_COMPRESSED_DATA) if
# delta == 0 or the quantised range would not fit 32 bits -- and, here, if it holds a non-finite
# value (CFITSIO would code NaN as ZBLANK; the reduced image has none: they were zeroed and masked).
# -------------------------------------------------------------------------------------------
N_RESERVED_VALUES = 10


def _lower_median(vals, m):
    """Element (m - 1) // 2 of the m sorted entries of a zero-initialised array holding ``vals``."""
    nz = m - vals.size
    r = (m - 1) // 2
    if r < nz:
        return np.float32(0.0)
    return np.partition(vals, r - nz)[r - nz]


def fn_noise5_row(row):
    """-> (min, max, noise2, noise3, noise5) of one row (float32 in, float64 noise out)."""
    v = np.asarray(row, dtype=np.float32)
    n = v.size
    lo, hi = v.min(), v.max()
    if n < 9:
        return lo, hi, 0.0, 0.0, 0.0
    f = np.float32
    v1, v3, v4, v5, v6, v7, v9 = v[0:n - 8], v[2:n - 6], v[3:n - 5], v[4:n - 4], v[5:n - 3], v[6:n - 2], v[8:n]
    keep2 = ~((v5 == v6) & (v6 == v7))
    keep3 = ~((v3 == v4) & (v4 == v5) & (v5 == v6) & (v6 == v7))
    d2 = np.abs(v5 - v7)[keep2]
    d3 = np.abs((f(2) * v5 - v3) - v7)[keep3]
    d5 = np.abs((((f(6) * v5 - f(4) * v3) - f(4) * v7) + v1) + v9)[keep3]
    m = d3.size
    if m == 0:
        return lo, hi, 0.0, 0.0, 0.0
    med3, med5 = _lower_median(d3, m), _lower_median(d5, m)
    if m == 1:
        med2 = d2[0] if d2.size == 1 else f(0)
    else:
        med2 = _lower_median(d2, m) if d2.size > 1 else f(0)
    return lo, hi, 1.0483579 * float(med2), 0.6052697 * float(med3), 0.1772048 * float(med5)


def _nint(x):
    return np.where(x >= 0, np.trunc(x + 0.5), np.trunc(x - 0.5))


def fpack_quantize_row(values, tile_index, q=16.0, zdither0=1):
    """-> (int32 row, ZSCALE, ZZERO) as `fpack -q <q>` (SUBTRACTIVE_DITHER_1) stores the row, or
    (None, 0.0, 0.0) for a row it cannot quantise."""
    v = np.asarray(values, dtype=np.float32)
    if v.size <= 1 or not np.isfinite(v).all():
        return None, 0.0, 0.0
    lo, hi, n2, n3, n5 = fn_noise5_row(v)
    sigma = n3
    if n2 != 0.0 and n2 < sigma:
        sigma = n2
    if n5 != 0.0 and n5 < sigma:
        sigma = n5
    delta = sigma / float(np.float32(q))
    if delta == 0.0:
        return None, 0.0, 0.0
    lo, hi = float(lo), float(hi)
    if (hi - lo) / delta > 2.0 * 2147483647.0 - N_RESERVED_VALUES:
        return None, 0.0, 0.0
    if (hi - lo) / delta < 2147483647.0 - N_RESERVED_VALUES:
        zero = float(np.trunc(lo / delta + 0.5)) * delta
    else:
        zero = (lo + hi) / 2.0
    r = _dither_sequence(tile_index, zdither0, v.size).astype(np.float64)
    x = (v.astype(np.float64) - zero) / delta + r - 0.5
    return _nint(x).astype(np.int64).astype(np.int32), delta, zero


def write_fz_f32(path, img, header=None, q=16.0, dither=1, zdither0=1, pointer='P'):
    """float32 image as `fpack -q <q>` stores it (``fpack_quantize_row``; ``dither`` 0 / 2: the
    same scaling without dither / with exact zeros kept): RICE_1 with BYTEPIX 4, ZSCALE / ZZERO
    columns; rows that cannot be quantised go gzipped into GZIP_COMPRESSED_DATA.  Returns (path,
    what a reader must get back)."""
    import zlib
    from blackbox_b200 import fitsio
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape
    tiles, scales, zeros, back, gz = [], [], [], np.empty_like(img), {}
    for r in range(H):
        finite = np.isfinite(img[r])
        if not finite.all() and finite.sum() > 2:
            # CFITSIO codes NaN as ZBLANK; scale and zero point from the finite pixels (their choice
            # does not matter to a reader; bbx_fpack_f32 stores such a row losslessly instead)
            good = img[r][finite].astype(np.float64)
            noise = 1.4826 * np.median(np.abs(np.diff(good))) / np.sqrt(2)
            scale, zero = (float(noise / q) if noise > 0 else 1.0), float(good.min())
            qrow = quantize_tile(img[r], r, scale, zero, dither, zdither0)
            tiles.append(encode_tile(qrow, 4))
            back[r] = unquantize_tile(qrow, r, scale, zero, dither, zdither0)
            scales.append(scale)
            zeros.append(zero)
            continue
        qrow, scale, zero = fpack_quantize_row(img[r], r, q, zdither0)
        if qrow is None:
            co = zlib.compressobj(6, zlib.DEFLATED, 31)
            gz[r] = co.compress(img[r].astype('>f4').tobytes()) + co.flush()
            tiles.append(b'')
            back[r] = img[r]
        else:
            if dither != 1:
                qrow = quantize_tile(img[r], r, scale, zero, dither, zdither0)
            tiles.append(encode_tile(qrow, 4))
            back[r] = unquantize_tile(qrow, r, scale, zero, dither, zdither0)
        scales.append(scale)
        zeros.append(zero)
    card = fitsio._card
    method = {0: 'NO_DITHER', 1: 'SUBTRACTIVE_DITHER_1', 2: 'SUBTRACTIVE_DITHER_2'}[dither]
    extra = [card('ZQUANTIZ', method), card('ZBLANK', NULL_VALUE)]
    if dither:
        extra.append(card('ZDITHER0', zdither0))
    _fz_file(path, tiles, img.shape, -32, 4, header, pointer, False,
             extra_cols=[('ZSCALE', scales), ('ZZERO', zeros)], extra_cards=extra, gzip_tiles=gz)
    return path, back
