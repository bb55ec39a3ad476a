"""Minimal FITS primary-HDU reader / writer for the boundary of the hot path.

The reference reads raw frames with ``read_hdulist(..., dtype='float32')`` (astropy.io.fits on
fpacked files, blackbox.py:7653-7771; the raw data are 16-bit integers with BZERO = 32768, i.e.
unsigned counts) and writes the reduced float32 image and the uint8 mask with ``fits.writeto``
(blackbox.py:1987-1990).  astropy is not a dependency here; this module does the two things the
GPU path needs:

  * ``read_primary``  parse the header cards of an UNCOMPRESSED primary HDU and hand back the data
                      unit exactly as it is on disk (big-endian), as a numpy array or a pinned
                      ``torch`` tensor -- the byte swap and the BZERO offset happen on the GPU
                      (``reduce.fits_decode`` -> ``bbx_fits_decode``), so the host only moves bytes;
  * ``write_primary`` write a header + a data unit; the data may already be big-endian bytes
                      produced on the GPU (``reduce.fits_encode`` -> ``bbx_fits_encode``).

  * ``read_compressed`` parse a tile-compressed image (``.fits.fz`` as written by fpack: an empty
                      primary HDU and a binary table whose heap holds one Rice-coded tile per
                      image row) and hand back the heap bytes plus the per-tile descriptors --
                      the decoding is ``reduce.rice_decode`` -> ``bbx_rice_decode`` on the GPU.
                      ZCMPTYPE RICE_1, BLOCKSIZE 32, row tiles: 16-bit raw frames, 8-bit masks,
                      32-bit integers and quantised float images (ZSCALE / ZZERO, dithering);
                      anything else raises FitsError.
  * ``write_compressed`` the way back for tiles Rice-coded on the GPU (``reduce.rice_encode`` ->
                      ``bbx_rice_encode``): the losslessly fpacked uint8 mask of the reference.
"""
import collections
import os

import numpy as np

BLOCK = 2880
_DTYPES = {8: '>u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


class FitsError(ValueError):
    pass


def already_exists(filename, get_filename=False):
    """blackbox.py:787-807: is ``filename`` there, possibly fpacked / gzipped (or the other way
    round)?  With ``get_filename`` -> (exists, name of the file found)."""
    cands = [filename, filename + '.fz', filename + '.gz', filename.replace('.fz', ''), filename.replace('.gz', '')]
    for c in dict.fromkeys(cands):
        if os.path.isfile(c):
            return (True, c) if get_filename else True
    return (False, filename) if get_filename else False


def _parse_value(text):
    t = text.strip()
    if not t:
        return None
    if t.startswith("'"):
        end = 1
        out = []
        while end < len(t):                       # '' inside a string is a literal quote
            if t[end] == "'":
                if end + 1 < len(t) and t[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(t[end])
            end += 1
        return ''.join(out).rstrip()
    if t in ('T', 'F'):
        return t == 'T'
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t.replace('D', 'E').replace('d', 'e'))
    except ValueError:
        return t


def _split_card(card):
    """-> (key, value, comment) of one 80-character card (value None for commentary cards)."""
    key = card[:8].strip()
    if card[8:10] != '= ' or key in ('COMMENT', 'HISTORY', ''):
        return key, None, card[8:].rstrip()
    body = card[10:]
    if body.lstrip().startswith("'"):
        # the comment separator is the first '/' after the closing quote
        i = body.index("'") + 1
        while i < len(body):
            if body[i] == "'":
                if i + 1 < len(body) and body[i + 1] == "'":
                    i += 2
                    continue
                break
            i += 1
        rest = body[i + 1:]
        value = _parse_value(body[:i + 1])
        comment = rest.split('/', 1)[1].strip() if '/' in rest else ''
        return key, value, comment
    if '/' in body:
        v, c = body.split('/', 1)
        return key, _parse_value(v), c.strip()
    return key, _parse_value(body), ''


def read_header(fh):
    """Read the header of the HDU at the current file position.  -> (OrderedDict key ->
    (value, comment), number of header bytes)."""
    hdr = collections.OrderedDict()
    nbytes = 0
    while True:
        block = fh.read(BLOCK)
        if len(block) != BLOCK:
            raise FitsError('truncated FITS header')
        nbytes += BLOCK
        text = block.decode('ascii', 'replace')
        for i in range(0, BLOCK, 80):
            card = text[i:i + 80]
            key, value, comment = _split_card(card)
            if key == 'END':
                return hdr, nbytes
            if value is None:
                if key in ('COMMENT', 'HISTORY'):
                    hdr.setdefault(key, ([], ''))[0].append(comment)
                continue
            hdr[key] = (value, comment)


def read_primary(path, pinned=False):
    """-> (header, data, info).  ``data`` is the primary data unit as stored: a big-endian numpy
    array of shape (NAXIS2, NAXIS1) (a read-only memory map), or with ``pinned`` a pinned uint8
    ``torch`` tensor holding the same bytes (ready for an asynchronous host-to-device copy).
    ``info``: dict(bitpix, shape, bzero, bscale, offset).  Raises FitsError for anything but a
    2-D uncompressed image in the primary HDU."""
    with open(path, 'rb') as fh:
        hdr, hbytes = read_header(fh)
    if hdr.get('SIMPLE', (False,))[0] is not True:
        raise FitsError('{}: not a standard FITS file (SIMPLE != T)'.format(path))
    bitpix = hdr['BITPIX'][0]
    if hdr.get('NAXIS', (0,))[0] != 2 or bitpix not in _DTYPES:
        raise FitsError('{}: primary HDU is not a 2-D image (NAXIS={}, BITPIX={}); tile-compressed '
                        'files have to be funpacked first'.format(path, hdr.get('NAXIS', (None,))[0], bitpix))
    shape = (int(hdr['NAXIS2'][0]), int(hdr['NAXIS1'][0]))
    dt = np.dtype(_DTYPES[bitpix])
    nbytes = shape[0] * shape[1] * dt.itemsize
    if os.path.getsize(path) < hbytes + nbytes:
        raise FitsError('{}: truncated data unit'.format(path))
    info = dict(bitpix=bitpix, shape=shape, bzero=float(hdr.get('BZERO', (0.0,))[0]),
                bscale=float(hdr.get('BSCALE', (1.0,))[0]), offset=hbytes)
    if pinned:
        import torch
        buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        with open(path, 'rb') as fh:
            fh.seek(hbytes)
            got = fh.readinto(buf.numpy())
        if got != nbytes:
            raise FitsError('{}: short read'.format(path))
        return hdr, buf, info
    data = np.memmap(path, dtype=dt, mode='r', offset=hbytes, shape=shape)
    return hdr, data, info


_TFORM_BYTES = {'L': 1, 'X': 1, 'B': 1, 'I': 2, 'J': 4, 'K': 8, 'A': 1, 'E': 4, 'D': 8, 'C': 8, 'M': 16,
                'P': 8, 'Q': 16}


def _tform_width(tform):
    """Bytes a binary-table column takes in a row: rTa / rPt(max) / rQt(max)."""
    t = str(tform).strip()
    i = 0
    while i < len(t) and t[i].isdigit():
        i += 1
    repeat = int(t[:i]) if i else 1
    code = t[i] if i < len(t) else ''
    if code not in _TFORM_BYTES:
        raise FitsError('unsupported TFORM {!r}'.format(tform))
    if code == 'X':
        return (repeat + 7) // 8, code
    return repeat * _TFORM_BYTES[code], code


class CompressedImage:
    """A tile-compressed image as it sits in the file, ready for the GPU: ``heap`` (the table's heap:
    numpy memory map or pinned torch tensor), ``offsets`` (int64) / ``lengths`` (int32) one per
    tile, ``info`` (see ``read_compressed``), ``header``; for float images ``zscale`` / ``zzero``
    (float64 per tile)."""

    def __init__(self, header, heap, offsets, lengths, info, zscale=None, zzero=None, fallback=None):
        self.header, self.heap, self.offsets, self.lengths, self.info = header, heap, offsets, lengths, info
        self.zscale, self.zzero = zscale, zzero
        # rows that are not Rice-coded (length 0): {row: pixel values}, unpacked on the host from the
        # GZIP_COMPRESSED_DATA column (CFITSIO stores a float row there when it cannot be quantised)
        self.fallback = fallback or {}
        self._desc = None

    def __iter__(self):                 # (header, heap, offsets, lengths, info) = read_compressed(...)
        return iter((self.header, self.heap, self.offsets, self.lengths, self.info))

    def descriptors(self):
        """offsets (int64) then lengths (int32) in ONE pinned uint8 tensor: a single host-to-device
        copy next to the heap's."""
        if self._desc is None:
            import torch
            n = len(self.offsets)
            buf = torch.empty(12 * n, dtype=torch.uint8).pin_memory()
            buf[:8 * n].view(torch.int64).copy_(torch.from_numpy(np.ascontiguousarray(self.offsets, dtype=np.int64)))
            buf[8 * n:].view(torch.int32).copy_(torch.from_numpy(np.ascontiguousarray(self.lengths, dtype=np.int32)))
            self._desc = buf
        return self._desc


def read_compressed(path, pinned=False):
    """-> CompressedImage (unpacks as ``(header, heap, offsets, lengths, info)``) of the first
    tile-compressed image of ``path``.  ``heap``: the table's heap as a uint8 numpy array (a memory
    map) or, with ``pinned``, a pinned ``torch`` tensor; ``offsets`` (int64) / ``lengths`` (int32):
    one entry per tile, byte offset into the heap and compressed size; ``info``: dict(bitpix, shape,
    bzero, bscale, blocksize, bytepix, tile_shape, quantize, zdither0, zblank).  ``header`` merges the
    primary and the extension keywords.  Supported: RICE_1, BLOCKSIZE 32, row tiles, ZBITPIX 8 / 16 /
    32 with BYTEPIX 1 / 2 / 4 and ZBITPIX -32 (float images quantised to int32, ZSCALE / ZZERO
    columns, NO_DITHER / SUBTRACTIVE_DITHER_1 / _2); anything else raises FitsError."""
    with open(path, 'rb') as fh:
        hdr, hbytes = read_header(fh)
        if hdr.get('NAXIS', (0,))[0] != 0:
            raise FitsError('{}: the primary HDU holds data; not an fpacked image'.format(path))
        ext, ebytes = read_header(fh)
        table_start = hbytes + ebytes
        get = lambda k, d=None: ext[k][0] if k in ext else d
        if get('XTENSION') != 'BINTABLE' or get('ZIMAGE') is not True:
            raise FitsError('{}: the first extension is not a tile-compressed image'.format(path))
        cmptype = str(get('ZCMPTYPE', '')).strip()
        if cmptype not in ('RICE_1', 'RICE_ONE'):
            raise FitsError('{}: ZCMPTYPE {!r} (RICE_1 is supported)'.format(path, cmptype))
        zbitpix = get('ZBITPIX')
        if zbitpix not in (8, 16, 32, -32) or get('ZNAXIS') != 2:
            raise FitsError('{}: ZBITPIX {} / ZNAXIS {} (8 / 16 / 32 / -32-bit 2-D images are supported)'.format(
                path, zbitpix, get('ZNAXIS')))
        shape = (int(get('ZNAXIS2')), int(get('ZNAXIS1')))
        tile = (int(get('ZTILE2', 1)), int(get('ZTILE1', shape[1])))
        if tile != (1, shape[1]):
            raise FitsError('{}: tiles of {} (row tiles are supported)'.format(path, tile))
        blocksize, bytepix = 32, 4
        n = 1
        while 'ZNAME{}'.format(n) in ext:
            name = str(get('ZNAME{}'.format(n))).strip()
            if name == 'BLOCKSIZE':
                blocksize = int(get('ZVAL{}'.format(n)))
            elif name == 'BYTEPIX':
                bytepix = int(get('ZVAL{}'.format(n)))
            n += 1
        want_bp = {8: (1,), 16: (2,), 32: (4,), -32: (4,)}[zbitpix]
        if blocksize != 32 or bytepix not in want_bp:
            raise FitsError('{}: BLOCKSIZE {} / BYTEPIX {} for ZBITPIX {} (32 / {} are supported)'.format(
                path, blocksize, bytepix, zbitpix, want_bp[0]))
        rowlen, nrows, pcount = int(get('NAXIS1')), int(get('NAXIS2')), int(get('PCOUNT', 0))
        if nrows != shape[0]:
            raise FitsError('{}: {} table rows for {} tiles'.format(path, nrows, shape[0]))
        col_off, cols = 0, {}
        for c in range(1, int(get('TFIELDS')) + 1):
            width, code = _tform_width(get('TFORM{}'.format(c)))
            cols[str(get('TTYPE{}'.format(c), '')).strip()] = (col_off, code, width)
            col_off += width
        found = cols.get('COMPRESSED_DATA')
        if found is None or found[1] not in ('P', 'Q'):
            raise FitsError('{}: no COMPRESSED_DATA column of variable-length bytes'.format(path))
        fh.seek(table_start)
        table = fh.read(rowlen * nrows)
        if len(table) != rowlen * nrows:
            raise FitsError('{}: truncated table'.format(path))
        rows = np.frombuffer(table, dtype=np.uint8).reshape(nrows, rowlen)
        if found[1] == 'P':
            desc = rows[:, found[0]:found[0] + 8].copy().view('>i4').astype(np.int64)
        else:
            desc = rows[:, found[0]:found[0] + 16].copy().view('>i8').astype(np.int64)
        lengths, offsets = desc[:, 0], desc[:, 1]
        gz_rows = {}
        if (lengths <= 0).any():
            gz = cols.get('GZIP_COMPRESSED_DATA')
            if gz is None or gz[1] not in ('P', 'Q') or (lengths < 0).any():
                raise FitsError('{}: {} tile(s) without Rice-coded bytes and no fall-back column'.format(
                    path, int((lengths <= 0).sum())))
            w = 8 if gz[1] == 'P' else 16
            gdesc = rows[:, gz[0]:gz[0] + w].copy().view('>i4' if gz[1] == 'P' else '>i8').astype(np.int64)
            for r in np.nonzero(lengths == 0)[0]:
                if gdesc[r, 0] <= 0:
                    raise FitsError('{}: tile {} has no data in either column'.format(path, int(r)))
                gz_rows[int(r)] = (int(gdesc[r, 1]), int(gdesc[r, 0]))
            offsets = np.where(lengths == 0, 0, offsets)
        zscale = zzero = None
        quantize, zdither0, zblank = None, 0, None
        if zbitpix == -32:
            quantize = str(get('ZQUANTIZ', 'NO_DITHER')).strip().upper()
            if quantize not in ('NO_DITHER', 'SUBTRACTIVE_DITHER_1', 'SUBTRACTIVE_DITHER_2'):
                raise FitsError('{}: ZQUANTIZ {!r}'.format(path, quantize))
            per_tile = {}
            for name in ('ZSCALE', 'ZZERO'):
                if name in cols:
                    o, code, width = cols[name]
                    if code not in ('D', 'E') or width not in (8, 4):
                        raise FitsError('{}: column {} has TFORM code {}'.format(path, name, code))
                    per_tile[name] = rows[:, o:o + width].copy().view('>f8' if code == 'D' else '>f4')[:, 0].astype(np.float64)
                elif name in ext:
                    per_tile[name] = np.full(nrows, float(get(name)))
                else:
                    raise FitsError('{}: float image without {}'.format(path, name))
            zscale, zzero = per_tile['ZSCALE'], per_tile['ZZERO']
            zdither0 = int(get('ZDITHER0', 1))
            if 'ZBLANK' in cols:
                raise FitsError('{}: per-tile ZBLANK column'.format(path))
            zblank = int(get('ZBLANK')) if 'ZBLANK' in ext else None
        theap = int(get('THEAP', rowlen * nrows))
        heap_start = table_start + theap
        heap_bytes = pcount - (theap - rowlen * nrows)
        if (offsets < 0).any() or (offsets + lengths > heap_bytes).any():
            raise FitsError('{}: tile descriptors point outside the heap'.format(path))
        if os.path.getsize(path) < heap_start + heap_bytes:
            raise FitsError('{}: truncated heap'.format(path))
        fallback = {}
        if gz_rows:
            import zlib
            dt = {8: 'u1', 16: '>i2', 32: '>i4', -32: '>f4'}[zbitpix]
            for r, (o, n) in gz_rows.items():
                if o < 0 or o + n > heap_bytes:
                    raise FitsError('{}: tile descriptors point outside the heap'.format(path))
                fh.seek(heap_start + o)
                try:
                    vals = np.frombuffer(zlib.decompress(fh.read(n), 47), dtype=dt)
                except zlib.error as exc:
                    raise FitsError('{}: gzipped tile {}: {}'.format(path, r, exc))
                if vals.size != shape[1]:
                    raise FitsError('{}: gzipped tile {} holds {} pixels'.format(path, r, vals.size))
                fallback[r] = vals.astype(vals.dtype.newbyteorder('='))
        if pinned:
            import torch
            heap = torch.empty(heap_bytes, dtype=torch.uint8).pin_memory()
            fh.seek(heap_start)
            if fh.readinto(heap.numpy()) != heap_bytes:
                raise FitsError('{}: short read'.format(path))
        else:
            heap = np.memmap(path, dtype=np.uint8, mode='r', offset=heap_start, shape=(heap_bytes,))
    merged = collections.OrderedDict(hdr)
    for k, v in ext.items():
        if k not in ('XTENSION', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'PCOUNT', 'GCOUNT', 'TFIELDS', 'THEAP') \
                and not k.startswith(('TTYPE', 'TFORM', 'ZNAME', 'ZVAL', 'ZTILE', 'ZNAXIS')) \
                and k not in ('ZIMAGE', 'ZCMPTYPE', 'ZBITPIX', 'ZSIMPLE', 'ZEXTEND', 'ZQUANTIZ', 'ZDITHER0', 'ZBLANK'):
            merged[k] = v
    info = dict(bitpix=zbitpix, shape=shape, bzero=float(get('BZERO', 0.0)), bscale=float(get('BSCALE', 1.0)),
                blocksize=blocksize, bytepix=bytepix, tile_shape=tile, quantize=quantize, zdither0=zdither0,
                zblank=zblank)
    return CompressedImage(merged, heap, offsets.astype(np.int64), lengths.astype(np.int32), info, zscale, zzero,
                           fallback)


_dither_table = None


def dither_random_table():
    """The 10000 random numbers the FITS standard defines for (un)dithering quantised float
    images: Park & Miller's minimal standard generator (a = 16807, m = 2^31 - 1, seed 1), value =
    seed / m in float32; the 10000th seed is 1043618065 (the standard's check value)."""
    global _dither_table
    if _dither_table is None:
        a, m, seed = 16807.0, 2147483647.0, 1.0
        out = np.empty(10000, dtype=np.float32)
        for i in range(10000):
            temp = a * seed
            seed = temp - m * int(temp / m)
            out[i] = seed / m
        if int(seed) != 1043618065:
            raise FitsError('dither table self-check failed')
        _dither_table = out
    return _dither_table


def write_compressed(path, heap, lengths, shape, zbitpix, header=None, zscale=None, zzero=None, zdither0=None,
                     lossless_rows=None):
    """Write a tile-compressed image the way fpack lays it out (empty primary HDU; BINTABLE with a
    COMPRESSED_DATA column of row tiles, RICE_1, BLOCKSIZE 32) from Rice-coded tiles that already
    exist -- ``heap``: the tiles back to back (numpy uint8 / pinned torch tensor), ``lengths``:
    bytes per tile, as ``bbx_rice_encode`` / ``bbx_fpack_f32`` leave them.  ``zbitpix`` 8 (the mask:
    what the reference gets from ``fpack -D -Y``, blackbox.py:826-827), 16, 32, or -32: a float image
    quantised by ``bbx_fpack_f32`` (``fpack -q 16``, blackbox.py:836) with its per-row ``zscale`` /
    ``zzero`` columns and the ``zdither0`` it was dithered with; ``lossless_rows``: {row: float32
    values} of the rows that were not quantised (length 0), gzipped into GZIP_COMPRESSED_DATA."""
    if zbitpix not in (8, 16, 32, -32):
        raise FitsError('write_compressed: ZBITPIX {}'.format(zbitpix))
    H, W = shape
    lens = np.asarray(lengths, dtype=np.int64)
    if lens.shape != (H,):
        raise FitsError('write_compressed: {} tile sizes for {} rows'.format(lens.size, H))
    raw = heap.numpy() if hasattr(heap, 'numpy') else np.asarray(heap)
    raw = raw.reshape(-1).view(np.uint8)
    total = int(lens.sum())
    if raw.size < total:
        raise FitsError('write_compressed: heap holds {} bytes, the tiles need {}'.format(raw.size, total))
    offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
    gz, glens, goffs = [], np.zeros(H, dtype=np.int64), np.zeros(H, dtype=np.int64)
    if zbitpix == -32:
        if zscale is None or zzero is None or zdither0 is None:
            raise FitsError('write_compressed: a float image needs zscale, zzero and zdither0')
        zs, zz = np.asarray(zscale, dtype=np.float64), np.asarray(zzero, dtype=np.float64)
        if zs.shape != (H,) or zz.shape != (H,):
            raise FitsError('write_compressed: {} / {} scale factors for {} rows'.format(zs.size, zz.size, H))
        missing = set(np.nonzero(lens == 0)[0].tolist()) - set(lossless_rows or {})
        if missing:
            raise FitsError('write_compressed: rows {} are neither Rice-coded nor given losslessly'.format(sorted(missing)[:8]))
        import zlib
        pos = total
        for r in sorted(lossless_rows or {}):
            if lens[r] != 0:
                continue
            vals = np.asarray(lossless_rows[r], dtype=np.float32)
            if vals.shape != (W,):
                raise FitsError('write_compressed: lossless row {} has shape {}'.format(r, vals.shape))
            co = zlib.compressobj(6, zlib.DEFLATED, 31)
            blob = co.compress(vals.astype('>f4').tobytes()) + co.flush()
            gz.append(blob)
            glens[r], goffs[r] = len(blob), pos
            pos += len(blob)
        offs = np.where(lens == 0, 0, offs)
    elif (lens <= 0).any():
        raise FitsError('write_compressed: empty tiles in an integer image')
    pcount = total + int(glens.sum())
    pointer = 'P' if pcount < 2 ** 31 else 'Q'
    ptype = '>i4' if pointer == 'P' else '>i8'
    pw = 8 if pointer == 'P' else 16
    columns = [np.stack([lens, offs], axis=1).astype(ptype).view(np.uint8).reshape(H, pw)]
    fields = [('COMPRESSED_DATA', '1{}B({})'.format(pointer, int(lens.max())))]
    if zbitpix == -32:
        if gz:
            columns.append(np.stack([glens, goffs], axis=1).astype(ptype).view(np.uint8).reshape(H, pw))
            fields.append(('GZIP_COMPRESSED_DATA', '1{}B({})'.format(pointer, int(glens.max()))))
        for name, vals in (('ZSCALE', zs), ('ZZERO', zz)):
            columns.append(np.ascontiguousarray(vals.astype('>f8')).view(np.uint8).reshape(H, 8))
            fields.append((name, '1D'))
    table = np.ascontiguousarray(np.concatenate(columns, axis=1))
    width = table.shape[1]
    primary = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0),
               _card('EXTEND', True), 'END'.ljust(80)]
    ext = ["XTENSION= 'BINTABLE'".ljust(80), _card('BITPIX', 8), _card('NAXIS', 2), _card('NAXIS1', width),
           _card('NAXIS2', H), _card('PCOUNT', pcount), _card('GCOUNT', 1), _card('TFIELDS', len(fields))]
    for n, (name, form) in enumerate(fields, 1):
        ext += [_card('TTYPE{}'.format(n), name), _card('TFORM{}'.format(n), form)]
    ext += [_card('ZIMAGE', True), _card('ZSIMPLE', True), _card('ZBITPIX', zbitpix), _card('ZNAXIS', 2),
            _card('ZNAXIS1', W), _card('ZNAXIS2', H), _card('ZTILE1', W), _card('ZTILE2', 1),
            _card('ZCMPTYPE', 'RICE_1'), _card('ZNAME1', 'BLOCKSIZE'), _card('ZVAL1', 32),
            _card('ZNAME2', 'BYTEPIX'), _card('ZVAL2', abs(zbitpix) // 8)]
    if zbitpix == -32:
        ext += [_card('ZQUANTIZ', 'SUBTRACTIVE_DITHER_1'), _card('ZDITHER0', int(zdither0))]
    # keywords that describe the layout of whatever file the header came from (a raw frame's
    # BZERO 32768 on a float image would be applied by every reader) stay behind
    skip = {'SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'END', 'EXTEND', 'XTENSION', 'PCOUNT', 'GCOUNT', 'TFIELDS',
            'BSCALE', 'BZERO', 'THEAP', 'CHECKSUM', 'DATASUM', 'ZHECKSUM', 'ZDATASUM'}
    for key, val in (header.items() if header is not None else ()):
        ku = str(key).upper()
        if ku in skip or ku in ('COMMENT', 'HISTORY') or ku.startswith(('TTYPE', 'TFORM', 'ZNAME', 'ZVAL', 'ZTILE', 'ZNAXIS')) \
                or ku in ('ZIMAGE', 'ZSIMPLE', 'ZEXTEND', 'ZBITPIX', 'ZCMPTYPE', 'ZQUANTIZ', 'ZDITHER0', 'ZBLANK', 'ZSCALE', 'ZZERO'):
            continue
        value, comment = (val if isinstance(val, tuple) and len(val) == 2 else (val, ''))
        ext.append(_card(key, value, comment))
    ext.append('END'.ljust(80))
    tmp = os.path.join(os.path.dirname(path) or '.', '.{}.{}.part'.format(
        os.path.basename(path).replace('.fits', '_fits'), os.getpid()))
    with open(tmp, 'wb') as fh:
        for cards in (primary, ext):
            text = ''.join(cards)
            fh.write((text + ' ' * (-len(text) % BLOCK)).encode('ascii', 'replace'))
        fh.write(table.tobytes())
        fh.write(memoryview(np.ascontiguousarray(raw[:total])))
        for blob in gz:
            fh.write(blob)
        fh.write(b'\0' * (-(table.nbytes + pcount) % BLOCK))
    os.replace(tmp, path)
    return path


def to_native(data, info):
    """Host-side decode of ``read_primary``'s array (tests, small frames): unsigned 16-bit counts
    for BITPIX 16 / BZERO 32768 / BSCALE 1, float32 ``bscale * x + bzero`` otherwise."""
    if info['bitpix'] == 16 and info['bzero'] == 32768.0 and info['bscale'] == 1.0:
        return (np.asarray(data).astype(np.int32) + 32768).astype(np.uint16)
    out = np.asarray(data).astype(np.float32)
    if info['bscale'] != 1.0:
        out *= np.float32(info['bscale'])
    if info['bzero'] != 0.0:
        out += np.float32(info['bzero'])
    return out


# -------------------------------------------------------------------------------------------
def _format_value(value):
    if isinstance(value, (bool, np.bool_)):
        return '{:>20}'.format('T' if value else 'F')
    if isinstance(value, (int, np.integer)):
        return '{:>20d}'.format(int(value))
    if isinstance(value, (float, np.floating)):
        v = float(value)
        if not np.isfinite(v):
            return "'{:<8}'".format(str(v))
        text = repr(v).upper()
        if 'E' not in text and '.' not in text:
            text += '.'
        return '{:>20}'.format(text)
    text = str(value).replace("'", "''")
    return "'{:<8}'".format(text[:68])


def _card(key, value, comment=''):
    key = str(key).upper()
    if len(key) > 8:
        raise FitsError('keyword {!r} is longer than 8 characters'.format(key))
    card = '{:<8}= {}'.format(key, _format_value(value))
    if comment:
        card += ' / ' + str(comment)
    return card[:80].ljust(80)


def build_header(shape, bitpix, header=None, bzero=None):
    """ASCII header block(s) of a primary HDU for an image of ``shape`` (rows, columns)."""
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', bitpix, 'array data type'),
             _card('NAXIS', 2, 'number of array dimensions'), _card('NAXIS1', shape[1]), _card('NAXIS2', shape[0])]
    if bzero is not None:
        cards += [_card('BSCALE', 1), _card('BZERO', int(bzero))]
    skip = {'SIMPLE', 'BITPIX', 'NAXIS', 'NAXIS1', 'NAXIS2', 'BSCALE', 'BZERO', 'END', 'EXTEND'}
    for key, val in (header.items() if header is not None else ()):
        if str(key).upper() in skip:
            continue
        if str(key).upper() in ('COMMENT', 'HISTORY'):
            lines = val[0] if isinstance(val, tuple) else val
            for ln in ([lines] if isinstance(lines, str) else lines):
                cards.append('{:<8}{}'.format(str(key).upper(), str(ln))[:80].ljust(80))
            continue
        value, comment = (val if isinstance(val, tuple) and len(val) == 2 else (val, ''))
        cards.append(_card(key, value, comment))
    cards.append('END'.ljust(80))
    text = ''.join(cards)
    text += ' ' * (-len(text) % BLOCK)
    return text.encode('ascii', 'replace')


def write_primary(path, data, header=None, be_bytes=False, shape=None, bitpix=None, bzero=None):
    """Write a primary HDU.  ``data``: a numpy array (float32 / uint8 / int16 / uint16 -- uint16 is
    stored as int16 with BZERO 32768), or with ``be_bytes`` a buffer (numpy uint8 array / pinned
    torch tensor) that already holds the big-endian data unit of an image of ``shape`` and
    ``bitpix`` (``reduce.fits_encode``; ``bzero=32768`` for encoded uint16 counts).  The file is
    written to a temporary name (one no ``*.fits*`` listing can pick up) and renamed."""
    if be_bytes:
        if shape is None or bitpix not in _DTYPES:
            raise FitsError('be_bytes needs shape and bitpix')
        raw = data.numpy() if hasattr(data, 'numpy') else np.asarray(data)
        raw = raw.reshape(-1).view(np.uint8)
        if raw.size != shape[0] * shape[1] * abs(bitpix) // 8:
            raise FitsError('buffer size {} does not match shape {} / BITPIX {}'.format(raw.size, shape, bitpix))
    else:
        a = np.asarray(data)
        if a.ndim != 2:
            raise FitsError('only 2-D images are supported')
        shape = a.shape
        if a.dtype == np.uint16:
            bitpix, bzero = 16, 32768
            a = (a.astype(np.int32) - 32768).astype('>i2')
        else:
            table = {np.dtype(np.float32): -32, np.dtype(np.float64): -64, np.dtype(np.uint8): 8,
                     np.dtype(np.int16): 16, np.dtype(np.int32): 32}
            if a.dtype.newbyteorder('=') not in table:
                raise FitsError('unsupported dtype {}'.format(a.dtype))
            bitpix = table[a.dtype.newbyteorder('=')]
            a = a.astype(_DTYPES[bitpix], copy=False)
        raw = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    tmp = os.path.join(os.path.dirname(path) or '.', '.{}.{}.part'.format(
        os.path.basename(path).replace('.fits', '_fits'), os.getpid()))
    with open(tmp, 'wb') as fh:
        fh.write(build_header(shape, bitpix, header, bzero=bzero))
        fh.write(raw.tobytes() if not raw.flags['C_CONTIGUOUS'] else memoryview(raw))
        fh.write(b'\0' * (-raw.size % BLOCK))
    os.replace(tmp, path)
    return path
