"""Rice coding of 16-bit image tiles -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference reads its raw frames through zogy's read_hdulist (blackbox.py:1451; copy at
blackbox_slurm_google.py:1144-1253), i.e. astropy.io.fits on fpacked files, which hands the tiles
to CFITSIO's ricecomp.c.  Neither astropy nor CFITSIO is in /root/reference or in this image
(third-party, versions unpinned), so this module restates the published algorithm
(R. White's Rice coder as distributed with CFITSIO, functions fits_rcomp_short /
fits_rdecomp_short; FITS tiled-image convention, Pence, White & Seaman 2010):

  tile  = first pixel, 16 bits, big-endian; then blocks of ``nblock`` = 32 pixels
  block = 4 bits FS+1 (0: every difference is 0; 15: differences as 16 raw bits), then per pixel
          (diff >> FS) zero bits, a one bit, the low FS bits of diff
  diff  = zig-zag mapped difference to the previous pixel in 16-bit arithmetic
  FS    = number of bits of  ((sum(diff) - nblock/2 - 1) / nblock) >> 1   (the encoder's choice;
          any FS decodes)

PARITY UNPINNED: no fpack / CFITSIO / astropy here to produce or read a real .fz file; the two
hand-derived known-answer vectors in tests/test_zz_rice_fz.py follow from the format text above.
"""
import numpy as np

FSBITS, FSMAX, BBITS = 4, 14, 16


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


def encode_tile16(a, nblock=32):
    """int16 (stored) pixel values of one tile -> bytes (fits_rcomp_short)."""
    a = np.asarray(a).astype(np.int16).astype(np.int64)
    w = _BitWriter()
    w.put(int(a[0]) & 0xffff, 16)
    lastpix = int(a[0])
    nx = a.size
    for i in range(0, nx, nblock):
        blk = a[i:i + nblock]
        thisblock = blk.size
        prev = np.concatenate(([lastpix], blk[:-1]))
        pdiff = ((blk - prev + 32768) % 65536) - 32768            # short arithmetic
        diff = np.where(pdiff < 0, ~(pdiff << 1), pdiff << 1) & 0xffffffff
        lastpix = int(blk[-1])
        pixelsum = float(diff.sum())
        dpsum = (pixelsum - (thisblock // 2) - 1) / thisblock
        if dpsum < 0:
            dpsum = 0.0
        psum = (int(dpsum) & 0xffff) >> 1
        fs = 0
        while psum > 0:
            psum >>= 1
            fs += 1
        if fs >= FSMAX:
            w.put(FSMAX + 1, FSBITS)
            for v in diff:
                w.put(int(v), BBITS)
        elif fs == 0 and pixelsum == 0:
            w.put(0, FSBITS)
        else:
            w.put(fs + 1, FSBITS)
            for v in diff:
                v = int(v)
                top = v >> fs
                while top >= 24:                                   # long runs of zeros in pieces
                    w.put(0, 24)
                    top -= 24
                w.put(1, top + 1)
                w.put(v & ((1 << fs) - 1), fs)
    return w.done()


def decode_tile16(buf, nx, nblock=32):
    """bytes -> int16 array of nx pixels (fits_rdecomp_short, statement by statement)."""
    c = memoryview(bytes(buf))
    out = np.zeros(nx, dtype=np.int64)
    lastpix = (c[0] << 8) | c[1]
    pos = 2
    b = c[pos]
    pos += 1
    nbits = 8
    i = 0
    while i < nx:
        nbits -= FSBITS
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
        elif fs == FSMAX:
            while i < imax:
                k = BBITS - nbits
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
                diff &= 0xffff
                diff = ~(diff >> 1) if diff & 1 else diff >> 1
                lastpix = (lastpix + diff) & 0xffff
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
                lastpix = (lastpix + diff) & 0xffff
                out[i] = lastpix
                i += 1
    return out.astype(np.uint16).view(np.int16)


def write_fz(path, counts, header=None, pointer='P', lead_column=False):
    """A tile-compressed FITS file of a uint16 image as fpack lays it out: empty primary HDU,
    BINTABLE with one COMPRESSED_DATA column of row tiles (BZERO 32768, RICE_1, BLOCKSIZE 32,
    BYTEPIX 2).  ``pointer``: 'P' (32-bit descriptors) or 'Q' (64-bit).  ``lead_column``: put an
    (empty) GZIP_COMPRESSED_DATA column in front, as CFITSIO does when it keeps a fall-back column."""
    from blackbox_b200 import fitsio
    counts = np.asarray(counts, dtype=np.uint16)
    H, W = counts.shape
    stored = (counts.astype(np.int32) - 32768).astype(np.int16)
    tiles = [encode_tile16(stored[r]) for r in range(H)]
    lens = np.array([len(t) for t in tiles], dtype=np.int64)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
    heap = b''.join(tiles)
    desc = np.stack([lens, offs], axis=1).astype('>i4' if pointer == 'P' else '>i8')
    if lead_column:
        desc = np.concatenate([np.zeros_like(desc), desc], axis=1).astype(desc.dtype)   # keeps big-endian
    width = desc.dtype.itemsize * desc.shape[1]
    card = fitsio._card
    primary = [card('SIMPLE', True), card('BITPIX', 16), card('NAXIS', 0), card('EXTEND', True), 'END'.ljust(80)]
    ext = [("XTENSION= 'BINTABLE'").ljust(80), card('BITPIX', 8), card('NAXIS', 2), card('NAXIS1', width),
           card('NAXIS2', H), card('PCOUNT', len(heap)), card('GCOUNT', 1), card('TFIELDS', 2 if lead_column else 1)]
    cols = (['GZIP_COMPRESSED_DATA'] if lead_column else []) + ['COMPRESSED_DATA']
    for n, name in enumerate(cols, 1):
        ext += [card('TTYPE{}'.format(n), name), card('TFORM{}'.format(n), '1{}B({})'.format(pointer, int(lens.max())))]
    ext += [
           card('ZIMAGE', True), card('ZSIMPLE', True), card('ZBITPIX', 16), card('ZNAXIS', 2), card('ZNAXIS1', W),
           card('ZNAXIS2', H), card('ZTILE1', W), card('ZTILE2', 1), card('ZCMPTYPE', 'RICE_1'),
           card('ZNAME1', 'BLOCKSIZE'), card('ZVAL1', 32), card('ZNAME2', 'BYTEPIX'), card('ZVAL2', 2),
           card('BSCALE', 1), card('BZERO', 32768)]
    for k, v in (header or {}).items():
        ext.append(card(k, v))
    ext.append('END'.ljust(80))
    with open(path, 'wb') as fh:
        for cards in (primary, ext):
            text = ''.join(cards)
            fh.write((text + ' ' * (-len(text) % fitsio.BLOCK)).encode('ascii'))
        body = desc.tobytes() + heap
        fh.write(body + b'\0' * (-len(body) % fitsio.BLOCK))
    return path
