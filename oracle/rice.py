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


def _fz_file(path, tiles, shape, zbitpix, bytepix, header, pointer, lead_column, extra_cols=None, extra_cards=()):
    """Lay out primary HDU + BINTABLE + heap the way fpack does.  ``extra_cols``: list of
    (name, float64 array per tile) stored as 1D columns behind COMPRESSED_DATA."""
    from blackbox_b200 import fitsio
    H, W = shape
    lens = np.array([len(t) for t in tiles], dtype=np.int64)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
    heap = b''.join(tiles)
    desc = np.stack([lens, offs], axis=1).astype('>i4' if pointer == 'P' else '>i8')
    rows = [desc.view(np.uint8).reshape(H, -1)]
    names = ['COMPRESSED_DATA']
    forms = ['1{}B({})'.format(pointer, int(lens.max()))]
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


def write_fz_f32(path, img, header=None, q=16.0, dither=1, zdither0=1, pointer='P'):
    """float32 image as `fpack -q` stores it: every row scaled to integers (ZSCALE = a noise
    estimate / q, ZZERO = the row minimum -- the decoder does not care how they were chosen),
    subtractive dithering, RICE_1 with BYTEPIX 4.  Returns (path, what a reader must get back)."""
    from blackbox_b200 import fitsio
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape
    tiles, scales, zeros, back = [], [], [], np.empty_like(img)
    for r in range(H):
        row = img[r].astype(np.float64)
        good = row[np.isfinite(row)]
        noise = 1.4826 * np.median(np.abs(np.diff(good))) / np.sqrt(2) if good.size > 2 else 1.0
        scale = float(noise / q) if noise > 0 else 1.0
        zero = float(good.min()) if good.size else 0.0
        qrow = quantize_tile(img[r], r, scale, zero, dither, zdither0)
        tiles.append(encode_tile(qrow, 4))
        scales.append(scale)
        zeros.append(zero)
        back[r] = unquantize_tile(qrow, r, scale, zero, dither, zdither0)
    card = fitsio._card
    method = {0: 'NO_DITHER', 1: 'SUBTRACTIVE_DITHER_1', 2: 'SUBTRACTIVE_DITHER_2'}[dither]
    extra = [card('ZQUANTIZ', method), card('ZBLANK', NULL_VALUE)]
    if dither:
        extra.append(card('ZDITHER0', zdither0))
    _fz_file(path, tiles, img.shape, -32, 4, header, pointer, False,
             extra_cols=[('ZSCALE', scales), ('ZZERO', zeros)], extra_cards=extra)
    return path, back
