"""Tile-compressed (.fits.fz, RICE_1) raw frames: the oracle codec against hand-derived vectors,
the host parser of blackbox_b200.fitsio, and the GPU decoder (bbx_rice_decode16)."""
import os

import numpy as np
import pytest


def _rows(seed, nx):
    """Rows that reach every branch of the coder: constant, quiet, noisy, white noise (raw 16-bit
    blocks), ramps across the int16 wrap, and a mix within one row."""
    rng = np.random.default_rng(seed)
    rows = [np.full(nx, 1234), rng.normal(3000, 2, nx), rng.normal(3000, 40, nx), rng.normal(3000, 900, nx),
            rng.integers(0, 65536, nx), np.arange(nx) * 37 % 65536, np.where(np.arange(nx) % 97 < 50, 0, 65535),
            np.concatenate([np.full(nx // 2, 7), rng.integers(0, 65536, nx - nx // 2)])]
    return np.clip(np.round(np.array(rows)), 0, 65535).astype(np.uint16)


def test_known_answers_from_the_format_text():
    from oracle import rice
    # 32 equal pixels: first pixel 0x0005, block code 0 ("all differences zero"), padded
    assert rice.encode_tile16(np.full(32, 5, np.int16)) == bytes([0x00, 0x05, 0x00])
    # [0, 1]: diffs 0 and +1 -> mapped 0 and 2; FS = 0 -> code 0001, then "1" and "001"
    assert rice.encode_tile16(np.array([0, 1], np.int16)) == bytes([0x00, 0x00, 0x19])
    assert list(rice.decode_tile16(bytes([0x00, 0x00, 0x19]), 2)) == [0, 1]
    # [-1, 0, -2] : first pixel 0xffff; diffs 0, +1, -2 -> 0, 2, 3; sum 5, (5-1-1)/3 = 1 -> psum 0, FS 0
    #   code 0001 | 1 | 001 | 0001 -> 0001 1001 0001 0000
    assert rice.encode_tile16(np.array([-1, 0, -2], np.int16)) == bytes([0xff, 0xff, 0x19, 0x10])
    assert list(rice.decode_tile16(bytes([0xff, 0xff, 0x19, 0x10]), 3)) == [-1, 0, -2]
    # [0, 3, 0, 3]: diffs 0, +3, -3, +3 -> 0, 6, 5, 6; sum 17, (17-2-1)/4 = 3.5 -> psum 1 -> FS 1
    #   code 0010 | 1 0 | 0001 0 | 001 1 | 0001 0 -> 0010 1000 0100 0110 0010 (0000)
    assert rice.encode_tile16(np.array([0, 3, 0, 3], np.int16)) == bytes.fromhex('0000284620')
    assert list(rice.decode_tile16(bytes.fromhex('0000284620'), 4)) == [0, 3, 0, 3]
    # [0, 20000]: diffs 0, 40000; (40000-1-1)/2 = 19999 -> psum 9999 -> FS 14 = FSMAX: raw block,
    #   code 1111 | 0x0000 | 0x9c40
    assert rice.encode_tile16(np.array([0, 20000], np.int16)) == bytes.fromhex('0000f00009c400')
    assert list(rice.decode_tile16(bytes.fromhex('0000f00009c400'), 2)) == [0, 20000]


@pytest.mark.parametrize('nx', [1, 31, 32, 33, 1000, 1500])
def test_oracle_round_trip(nx):
    from oracle import rice
    rows = _rows(nx, nx)
    stored = (rows.astype(np.int32) - 32768).astype(np.int16)
    for r in stored:
        buf = rice.encode_tile16(r)
        assert np.array_equal(rice.decode_tile16(buf, nx), r)
    # white noise does not compress: raw blocks, 16 bits per pixel + the block codes
    assert len(rice.encode_tile16(stored[4])) >= 2 * nx


@pytest.mark.parametrize('pointer,lead', [('P', False), ('Q', False), ('P', True), ('Q', True)])
def test_read_compressed_parses_the_table(tmp_path, pointer, lead):
    from blackbox_b200 import fitsio
    from oracle import rice
    rows = _rows(3, 700)
    path = rice.write_fz(str(tmp_path / 'raw.fits.fz'), rows, {'EXPTIME': 60.0, 'FILTER': 'q'}, pointer=pointer,
                         lead_column=lead)
    hdr, heap, offs, lens, info = fitsio.read_compressed(path)
    assert info['shape'] == rows.shape and info['bzero'] == 32768.0 and info['blocksize'] == 32
    assert hdr['EXPTIME'][0] == 60.0 and hdr['FILTER'][0] == 'q' and 'TFORM1' not in hdr
    assert offs.dtype == np.int64 and lens.dtype == np.int32 and len(offs) == rows.shape[0]
    heap = np.asarray(heap)
    for r in range(rows.shape[0]):
        got = rice.decode_tile16(heap[offs[r]:offs[r] + lens[r]].tobytes(), rows.shape[1])
        assert np.array_equal((got.astype(np.int32) + 32768).astype(np.uint16), rows[r])


def test_read_compressed_rejects_what_it_cannot_decode(tmp_path):
    from blackbox_b200 import fitsio
    from oracle import rice
    plain = str(tmp_path / 'plain.fits')
    fitsio.write_primary(plain, np.zeros((4, 4), np.uint16))
    with pytest.raises(fitsio.FitsError):
        fitsio.read_compressed(plain)
    path = rice.write_fz(str(tmp_path / 'raw.fits.fz'), _rows(1, 64))
    raw = bytearray(open(path, 'rb').read())
    for old, new in ((b"'RICE_1  '", b"'GZIP_1  '"), (b'ZTILE2  =                    1', b'ZTILE2  =                    2')):
        bad = str(tmp_path / 'bad.fits.fz')
        open(bad, 'wb').write(bytes(raw).replace(old, new))
        with pytest.raises(fitsio.FitsError):
            fitsio.read_compressed(bad)
    open(str(tmp_path / 'short.fits.fz'), 'wb').write(bytes(raw[:len(raw) - fitsio.BLOCK]))
    with pytest.raises(fitsio.FitsError):
        fitsio.read_compressed(str(tmp_path / 'short.fits.fz'))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(8, 1500), (70, 1000), (33, 31), (64, 12000)])
def test_gpu_decodes_what_the_oracle_encodes(tmp_path, shape):
    import torch
    from blackbox_b200 import fitsio, reduce as bbr
    from oracle import rice
    H, W = shape
    base = _rows(H, W)
    rows = np.concatenate([base] * (H // len(base) + 1))[:H]
    if W == 12000:                                                  # a synthetic raw frame strip
        from blackbox_b200 import synth
        rows = synth.make_raw('BG3', 4001, ysize_chan=H // 2 - 20)[0][:H].copy()
        rows[5] = base[4, :W] if W <= base.shape[1] else np.resize(base[4], W)
    path = rice.write_fz(str(tmp_path / 'raw.fits.fz'), rows)
    hdr, heap, offs, lens, info = fitsio.read_compressed(path, pinned=True)
    got = bbr.rice_decode(heap.cuda(non_blocking=True), offs, lens, info)
    assert got.dtype == torch.uint16 and tuple(got.shape) == shape
    assert np.array_equal(got.cpu().numpy(), rows)
    # a truncated tile is reported, not silently mis-decoded
    lens2 = lens.copy()
    lens2[H // 2] = max(3, lens2[H // 2] // 2)
    if lens2[H // 2] != lens[H // 2]:
        with pytest.raises(ValueError):
            bbr.rice_decode(heap.cuda(), offs, lens2, info)
    offs2 = offs.copy()
    offs2[0] = heap.numel()
    with pytest.raises(ValueError):
        bbr.rice_decode(heap.cuda(), offs2, lens, info)


# ---------------------------------------------------------------------------------------------
# 8- and 32-bit tiles, quantised float images, and the encoder on the GPU
# ---------------------------------------------------------------------------------------------
def _int_rows(seed, nx, bytepix):
    """Rows for every branch of the coder at a given pixel width (values as stored: signed)."""
    rng = np.random.default_rng(seed)
    lo, hi = -(1 << (8 * bytepix - 1)), (1 << (8 * bytepix - 1))
    rows = [np.full(nx, 7), rng.normal(100, 2, nx), rng.normal(0, 40, nx), rng.integers(lo, hi, nx),
            np.where(rng.random(nx) < 0.01, 64, 0), np.cumsum(rng.integers(-3, 4, nx)),
            np.where(np.arange(nx) % 97 < 50, lo, hi - 1),
            np.concatenate([np.zeros(nx // 2), rng.integers(lo, hi, nx - nx // 2)]),
            np.where(np.arange(nx) == nx // 3, hi - 1, 0)]            # one huge outlier in a quiet row
    rows = np.clip(np.round(np.array(rows, dtype=np.float64)), lo, hi - 1).astype(np.int64)
    return rows.astype({1: np.int8, 2: np.int16, 4: np.int32}[bytepix])


def test_known_answers_for_bytes_and_ints():
    from oracle import rice
    # BYTEPIX 1: first pixel 8 bits, 3-bit block code.  32 zeros: 0x00 | 000 -> 00 00
    assert rice.encode_tile(np.zeros(32, np.int8), 1) == bytes([0x00, 0x00])
    # [0, 1]: mapped diffs 0, 2; FS 0 -> code 001 | 1 | 001 -> 0011 0010
    assert rice.encode_tile(np.array([0, 1], np.int8), 1) == bytes([0x00, 0x32])
    assert list(rice.decode_tile(bytes([0x00, 0x32]), 2, 1)) == [0, 1]
    # a mask row: 40 zeros, one 64, zeros: block 0 all-zero (000), block 1 has diffs +64, -64 -> 128, 127
    row = np.zeros(64, np.int8); row[40] = 64
    buf = rice.encode_tile(row, 1)
    assert list(rice.decode_tile(buf, 64, 1)) == list(row.astype(np.int64))
    # BYTEPIX 4: first pixel 32 bits, 5-bit code.  [5, 5]: 00 00 00 05 | 00000 -> 5 bytes
    assert rice.encode_tile(np.array([5, 5], np.int32), 4) == bytes([0, 0, 0, 5, 0])
    # [0, -1]: mapped diffs 0, 1; (1 - 1 - 1)/2 < 0 -> FS 0 -> code 00001 | 1 | 01 -> 0000 1101
    assert rice.encode_tile(np.array([0, -1], np.int32), 4) == bytes([0, 0, 0, 0, 0x0d])
    assert list(rice.decode_tile(bytes([0, 0, 0, 0, 0x0d]), 2, 4)) == [0, 0xffffffff]


@pytest.mark.parametrize('bytepix', [1, 2, 4])
@pytest.mark.parametrize('nx', [1, 31, 32, 33, 700])
def test_oracle_round_trip_all_widths(bytepix, nx):
    from oracle import rice
    for r in _int_rows(nx + bytepix, nx, bytepix):
        buf = rice.encode_tile(r, bytepix)
        want = r.astype(np.int64) & ((1 << (8 * bytepix)) - 1)
        assert np.array_equal(rice.decode_tile(buf, nx, bytepix), want)


def test_dither_table_is_the_published_sequence():
    """Park-Miller minimal standard generator; the FITS standard's check value for the 10000th seed
    is asserted inside both implementations."""
    from blackbox_b200 import fitsio
    from oracle import rice
    a, b = fitsio.dither_random_table(), rice.random_table()
    assert a.dtype == np.float32 and a.shape == (10000,) and np.array_equal(a, b)
    assert a[0] == np.float32(16807.0 / 2147483647.0)


def test_fz_files_of_masks_and_float_images_parse(tmp_path):
    from blackbox_b200 import fitsio
    from oracle import rice
    rng = np.random.default_rng(4)
    mask = np.where(rng.random((40, 300)) < 0.01, rng.choice([1, 2, 4, 8, 32, 64], (40, 300)), 0).astype(np.uint8)
    ci = fitsio.read_compressed(rice.write_fz_u8(str(tmp_path / 'bpm.fits.fz'), mask, {'FILTER': 'q'}))
    assert ci.info['bitpix'] == 8 and ci.info['bytepix'] == 1 and ci.info['shape'] == mask.shape
    heap = np.asarray(ci.heap)
    for r in range(mask.shape[0]):
        got = rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(), mask.shape[1], 1)
        assert np.array_equal(got.astype(np.uint8), mask[r])
    img = rng.normal(300, 12, (20, 257)).astype(np.float32)
    img[3, 7] = np.nan
    for dither in (0, 1, 2):
        if dither == 2:
            img[5, 9] = 0.0
        path, back = rice.write_fz_f32(str(tmp_path / 'red{}.fits.fz'.format(dither)), img, {'MEDSEC': 301.5},
                                       dither=dither, zdither0=37)
        ci = fitsio.read_compressed(path)
        assert ci.info['bitpix'] == -32 and ci.info['bytepix'] == 4 and ci.info['zdither0'] == (37 if dither else 1)
        assert ci.header['MEDSEC'][0] == 301.5 and ci.zscale.shape == (20,) and ci.zzero.dtype == np.float64
        # quantisation error below half a step, NaN and exact zero survive
        err = np.abs(back - img)
        assert np.nanmax(err / ci.zscale[:, None]) <= 0.5 + 1e-6 and np.isnan(back[3, 7])
        if dither == 2:
            assert back[5, 9] == 0.0


def test_write_compressed_round_trip_on_the_host(tmp_path):
    """fitsio.write_compressed lays out tiles coded elsewhere (here: by the oracle) so that
    read_compressed and the oracle decoder get the image back."""
    from blackbox_b200 import fitsio
    from oracle import rice
    rng = np.random.default_rng(8)
    mask = np.where(rng.random((33, 500)) < 0.02, 4, 0).astype(np.uint8)
    tiles = [rice.encode_tile(mask[r], 1) for r in range(mask.shape[0])]
    path = fitsio.write_compressed(str(tmp_path / 'x_mask.fits.fz'), np.frombuffer(b''.join(tiles), np.uint8),
                                   [len(t) for t in tiles], mask.shape, 8, {'M-BPNUM': (3, 'number of bad pixels')})
    ci = fitsio.read_compressed(path)
    assert ci.header['M-BPNUM'][0] == 3 and ci.info['bitpix'] == 8
    heap = np.asarray(ci.heap)
    for r in range(mask.shape[0]):
        got = rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(), mask.shape[1], 1)
        assert np.array_equal(got.astype(np.uint8), mask[r])
    assert not [f for f in os.listdir(str(tmp_path)) if f.endswith('.part')]


@pytest.mark.gpu
@pytest.mark.parametrize('bytepix', [1, 2, 4])
@pytest.mark.parametrize('nx', [1, 33, 96, 1000, 1504])
def test_gpu_codec_equals_the_oracle_byte_for_byte(bytepix, nx):
    """bbx_rice_encode writes the very bytes the restated fits_rcomp_* writes, for every kind of
    block; bbx_rice_decode reads them back; both for all three pixel widths."""
    import torch
    from blackbox_b200 import reduce as bbr
    from oracle import rice
    rows = np.concatenate([_int_rows(nx + 10 * s, nx, bytepix) for s in range(4)])
    H = rows.shape[0]
    heap, lens = bbr.rice_encode(rows.view(np.uint8) if bytepix == 1 else rows)
    want = [rice.encode_tile(rows[r], bytepix) for r in range(H)]
    assert list(lens) == [len(t) for t in want]
    assert heap.tobytes() == b''.join(want)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    info = dict(shape=(H, nx), bitpix=8 * bytepix, bytepix=bytepix, bzero=0.0, bscale=1.0, blocksize=32)
    got = bbr.rice_decode(torch.from_numpy(heap).cuda(), offs, lens.astype(np.int32), info)
    assert np.array_equal(got.cpu().numpy().view(rows.dtype), rows)


@pytest.mark.gpu
@pytest.mark.parametrize('dither', [0, 1, 2])
def test_gpu_reads_fpacked_float_images(tmp_path, dither):
    """read_fits_image on a quantised float image (what `fpack -q` leaves in the reference's red
    folders): the oracle's un-quantised values bit for bit, NaN / exact-zero pixels included."""
    from blackbox_b200 import reduce as bbr
    from oracle import rice
    rng = np.random.default_rng(21)
    img = rng.normal(500, 20, (37, 12000 if dither == 1 else 333)).astype(np.float32)   # 12000: the dither sequence wraps
    img[2, 5] = np.nan
    img[4, 6] = 0.0
    path, back = rice.write_fz_f32(str(tmp_path / 'f.fits.fz'), img, {'MEDSEC': 499.0}, dither=dither, zdither0=9990)
    hdr, got = bbr.read_fits_image(path)
    assert hdr['MEDSEC'] == 499.0 and got.dtype.is_floating_point
    assert np.array_equal(got.cpu().numpy(), back, equal_nan=True)


@pytest.mark.gpu
def test_fullsize_mask_and_raw_frame_round_trip():
    """Size-independent property at BASELINE's full size: encode -> decode is the identity for a
    10560^2 mask and a 10600 x 12000 raw frame; sizes are what the bench moves over PCIe."""
    import torch
    from blackbox_b200 import reduce as bbr, synth
    rng = np.random.default_rng(3)
    mask = np.zeros((10560, 10560), np.uint8)
    ys, xs = rng.integers(0, 10560, 200000), rng.integers(0, 10560, 200000)
    mask[ys, xs] = rng.choice([1, 2, 4, 8, 64], 200000).astype(np.uint8)
    mask[:20] |= 32; mask[-20:] |= 32; mask[:, :20] |= 32; mask[:, -20:] |= 32
    heap, lens = bbr.rice_encode(mask)
    assert heap.size < mask.size // 20
    offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
    info = dict(shape=mask.shape, bitpix=8, bytepix=1, bzero=0.0, bscale=1.0, blocksize=32)
    got = bbr.rice_decode(torch.from_numpy(heap).cuda(), offs, lens, info)
    assert torch.equal(got.cpu(), torch.from_numpy(mask))
    del got
    raw = synth.make_raw('BG3', 4001)[0]
    heap, lens = bbr.rice_encode(raw)
    assert heap.size < raw.size * 2 * 0.45
    offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
    info = dict(shape=raw.shape, bitpix=16, bytepix=2, bzero=32768.0, bscale=1.0, blocksize=32)
    got = bbr.rice_decode(torch.from_numpy(heap).cuda(), offs, lens, info)
    assert got.dtype == torch.uint16 and np.array_equal(got.cpu().numpy(), raw)


@pytest.mark.gpu
def test_mask_init_reads_an_fpacked_bad_pixel_mask(tmp_path, small_bb, monkeypatch):
    """The reference's bad-pixel masks are .fits.fz (Settings/set_blackbox.py:187-193): mask_init finds
    the fpacked file through already_exists and decodes it on the GPU."""
    from blackbox_b200 import reduce as bbr, set_bb
    from oracle import reduce as R, rice
    small_bb(96, 132)
    rng = np.random.default_rng(12)
    data = rng.normal(300, 10, (192, 1056)).astype(np.float32)
    data[50:54, 100:104] = 3e5
    bpm = np.zeros(data.shape, np.uint8)
    bpm[rng.random(data.shape) < 0.003] = 1
    bpm[:4] = 32; bpm[-4:] = 32
    rice.write_fz_u8(str(tmp_path / 'BG3_bpm_q_0p2.fits.fz'), bpm)
    monkeypatch.setattr(set_bb, 'bad_pixel_mask', {'BG3': str(tmp_path / 'BG3_bpm_0p2.fits.fz')})
    bbr._bpm_registry.clear()
    bbr.tel = 'BG3'
    hdr = {'BIASM{}'.format(i + 1): 3200.0 for i in range(16)}
    mask_o, _ = R.mask_init(data.copy(), dict(hdr), bpm, 'object', tel='BG3')
    mask_g, _ = bbr.mask_init(data.copy(), dict(hdr), 'q', 'object')
    assert np.array_equal(mask_g, mask_o) and (mask_g & 32).any() and (mask_g & 1).any()


@pytest.mark.gpu
def test_run_host_with_fpacked_frames_in_and_rice_coded_mask_out(tmp_path, small_bb):
    """BatchReducer.run_host fed with .fits.fz raw frames (CompressedImage) and asked for the mask as
    the reference's fpacked product: image and (decoded) mask equal the plain path bit for bit; the
    mask file written from the device's bytes reads back through the oracle."""
    import torch
    from blackbox_b200 import fitsio, reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import BatchReducer
    from oracle import rice
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 43, shape)
    coeffs = synth.make_xtalk(44)[3]
    raws = []
    for seed in (42, 43, 44, 45, 46):
        raw = synth.make_raw(tel, seed, nstars=100, ncosmics=60)[0]
        raw[40:46, 2000:2006] = 65535
        raws.append(raw)
    batch = BatchReducer(tel, raws[0].shape, depth=3, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=3)
    plain_in = [torch.from_numpy(r.view(np.int16)).view(torch.uint16).pin_memory() for r in raws]
    imgs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in raws]
    masks = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in raws]
    want = batch.run_host(plain_in, imgs, masks, exptimes=[30.0 + k for k in range(len(raws))])
    packed = []
    for k, r in enumerate(raws):
        path = rice.write_fz(str(tmp_path / 'raw{}.fits.fz'.format(k)), r, {'EXPTIME': 30.0 + k})
        packed.append(fitsio.read_compressed(path, pinned=True))
    imgs2 = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in raws]
    nbytes = batch.mask_fz_bytes(1 << 20)
    masks2 = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in raws]
    for rep in range(2):
        got = batch.run_host(packed, imgs2, masks2, mask_fz=True,
                             exptimes=[float(p.header['EXPTIME'][0]) for p in packed])
        for k in range(len(raws)):
            assert torch.equal(imgs2[k], imgs[k]), (rep, k)
            heap, lens = got[k].mask_fz
            assert heap.size < shape[0] * shape[1] // 8
            for row in (0, 7, shape[0] // 2, shape[0] - 1):
                o = int(np.sum(lens[:row], dtype=np.int64))
                dec = rice.decode_tile(heap[o:o + lens[row]].tobytes(), shape[1], 1).astype(np.uint8)
                assert np.array_equal(dec, masks[k][row].numpy()), (rep, k, row)
            assert got[k].header['NCOSMICS'] == want[k].header['NCOSMICS']
            assert got[k].header_mask == want[k].header_mask
    heap, lens = got[2].mask_fz
    path = fitsio.write_compressed(str(tmp_path / 'f_mask.fits.fz'), heap, lens, shape, 8,
                                   {k: (v, '') for k, v in got[2].header_mask.items()})
    _, back = bbr.read_fits_image(path)
    assert torch.equal(back.cpu(), masks[2])
    # a corrupt frame is reported
    bad = fitsio.read_compressed(str(tmp_path / 'raw0.fits.fz'), pinned=True)
    bad.lengths[5] = 3
    with pytest.raises(ValueError):
        batch.run_host([bad], imgs2, masks2, mask_fz=True)


@pytest.mark.parametrize('bytepix', [1, 2, 4])
def test_c_copy_of_the_codec_equals_the_python_statement(bytepix):
    """oracle/csrc/bbo.c holds the same coder in C (whole frames in seconds): same bytes, same pixels."""
    from oracle import rice
    for nx in (1, 32, 33, 700):
        for r in _int_rows(nx + 3 * bytepix, nx, bytepix):
            slow = rice.encode_tile(r, bytepix, fast=False)
            assert rice.encode_tile(r, bytepix, fast=True) == slow
            assert np.array_equal(rice.decode_tile(slow, nx, bytepix, fast=True),
                                  rice.decode_tile(slow, nx, bytepix, fast=False))
    with pytest.raises(ValueError):
        rice.decode_tile(rice.encode_tile(np.arange(300) * 7, bytepix)[:-20], 300, bytepix, fast=True)


# ---------------------------------------------------------------------------------------------
# `fpack -q 16` of the reduced image (blackbox.py:826-836)
# ---------------------------------------------------------------------------------------------
def _float_rows(seed, nx, nrows=24):
    """Rows with different noise levels and the special cases of the noise estimate: constant rows
    (not quantised), runs of equal values (d2 / d3 skip their entries), a negative background, a
    step, a row shorter on noise than on signal."""
    rng = np.random.default_rng(seed)
    img = np.empty((nrows, nx), dtype=np.float32)
    for r in range(nrows):
        img[r] = rng.normal(rng.uniform(-50, 900), rng.uniform(0.5, 60), nx)
    img[1] = 7.25                                         # constant: no noise, stored losslessly
    img[2, : nx // 2] = 0.0                               # half the row constant
    img[3] = np.round(img[3])                             # integers: equal neighbours now and then
    img[4] = np.round(rng.normal(0, 0.6, nx))             # mostly runs of equal values
    img[5] = np.where(np.arange(nx) < nx // 3, 100.0, 5000.0) + rng.normal(0, 3, nx)
    img[6] = 0.0
    img[6, 1] = 3.0                                       # differs only where no difference looks: no noise either
    img[8] = 0.0
    img[8, nx // 2] = 3.0                                 # five differences, three of them non-zero
    img[7] = rng.normal(-400, 9, nx)                      # negative zero point
    return img


def test_fpack_quantisation_statement(tmp_path):
    """The oracle's restatement of fits_quantize_float: ZSCALE = noise / q with the noise of a
    Gaussian row recovered to a few per cent by all three estimators, the zero point a whole
    number of steps, errors below half a step, unquantisable rows losslessly in the fall-back
    column -- and the file reads back through blackbox_b200.fitsio."""
    from blackbox_b200 import fitsio
    from oracle import rice
    rng = np.random.default_rng(12)
    row = rng.normal(300, 12, 10560).astype(np.float32)
    lo, hi, n2, n3, n5 = rice.fn_noise5_row(row)
    assert lo == row.min() and hi == row.max()
    for n in (n2, n3, n5):
        assert abs(n / 12.0 - 1.0) < 0.05
    q, scale, zero = rice.fpack_quantize_row(row, 3, 16.0, 77)
    assert scale == min(n2, n3, n5) / 16.0 and abs(zero / scale - round(zero / scale)) < 1e-9
    assert np.abs(rice.unquantize_tile(q, 3, scale, zero, 1, 77) - row).max() <= 0.5 * scale * (1 + 1e-6)
    assert rice.fpack_quantize_row(np.full(100, 3.5, np.float32), 0)[0] is None
    assert rice.fpack_quantize_row(np.array([1, 2, np.inf] * 10, np.float32), 0)[0] is None
    # differences of rows shorter than 9 pixels do not exist: no noise, not quantised
    assert rice.fpack_quantize_row(np.arange(8, dtype=np.float32), 0)[0] is None
    # d2's median is taken over count(d3) entries of a zero-filled array
    v = np.zeros(40, np.float32)
    v[10:30] = rng.normal(0, 1, 20)
    assert rice.fn_noise5_row(v)[2] <= 1.0483579 * np.median(np.abs(v[12:30] - v[10:28])) + 1e-6
    img = _float_rows(5, 333)
    path, back = rice.write_fz_f32(str(tmp_path / 'red.fits.fz'), img, {'S-BKG': 12.5}, zdither0=4321)
    ci = fitsio.read_compressed(path)
    assert sorted(ci.fallback) == [1, 6] and np.array_equal(ci.fallback[1], img[1]) and ci.info['zdither0'] == 4321
    assert (ci.lengths[[1, 6]] == 0).all() and (np.delete(ci.lengths, [1, 6]) > 0).all()
    heap = np.asarray(ci.heap)
    for r in range(img.shape[0]):
        if r in ci.fallback:
            continue
        qrow = rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(), img.shape[1], 4)
        qrow = np.asarray(qrow).astype(np.int64).astype(np.int32)
        assert np.array_equal(rice.unquantize_tile(qrow, r, ci.zscale[r], ci.zzero[r], 1, 4321), back[r])
        assert np.abs(back[r] - img[r]).max() <= 0.5 * ci.zscale[r] * (1 + 1e-6) + 1e-4


def test_write_compressed_float_layout_on_the_host(tmp_path):
    """fitsio.write_compressed(zbitpix=-32): Rice tiles + ZSCALE / ZZERO columns + gzipped rows,
    read back by read_compressed and decoded by the oracle."""
    from blackbox_b200 import fitsio
    from oracle import rice
    img = _float_rows(9, 200, nrows=10)
    tiles, zs, zz, lossless = [], [], [], {}
    for r in range(img.shape[0]):
        q, s, z = rice.fpack_quantize_row(img[r], r, 16.0, 55)
        tiles.append(b'' if q is None else rice.encode_tile(q, 4))
        zs.append(s)
        zz.append(z)
        if q is None:
            lossless[r] = img[r]
    with pytest.raises(fitsio.FitsError):
        fitsio.write_compressed(str(tmp_path / 'bad.fits.fz'), np.frombuffer(b''.join(tiles), np.uint8),
                                [len(t) for t in tiles], img.shape, -32, zscale=zs, zzero=zz, zdither0=55)
    path = fitsio.write_compressed(str(tmp_path / 'red.fits.fz'), np.frombuffer(b''.join(tiles), np.uint8),
                                   [len(t) for t in tiles], img.shape, -32, {'AIRMASS': 1.25}, zscale=zs, zzero=zz,
                                   zdither0=55, lossless_rows=lossless)
    ci = fitsio.read_compressed(path)
    assert ci.info['bitpix'] == -32 and ci.info['quantize'] == 'SUBTRACTIVE_DITHER_1' and ci.info['zdither0'] == 55
    assert ci.header['AIRMASS'][0] == 1.25 and sorted(ci.fallback) == sorted(lossless)
    heap = np.asarray(ci.heap)
    for r in range(img.shape[0]):
        if r in lossless:
            assert np.array_equal(ci.fallback[r], img[r])
            continue
        qrow = rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(), img.shape[1], 4)
        qrow = np.asarray(qrow).astype(np.int64).astype(np.int32)
        assert np.abs(rice.unquantize_tile(qrow, r, ci.zscale[r], ci.zzero[r], 1, 55) - img[r]).max() <= 0.5 * zs[r] * (1 + 1e-6) + 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('nx,zdither0', [(333, 1), (1000, 9999), (10560, 4242), (9, 17), (8, 3)])
def test_gpu_fpack_f32_equals_the_oracle(tmp_path, nx, zdither0):
    """bbx_fpack_f32: ZSCALE / ZZERO bit for bit what the oracle's fits_quantize_float restatement
    gives, the heap byte for byte the Rice code of its integers; unquantisable rows (constant, a
    non-finite pixel, fewer than 9 pixels) are left to the lossless column; the file that comes out
    reads back (on the GPU) to within half a quantisation step."""
    import torch
    from blackbox_b200 import fitsio, reduce as bbr
    from oracle import rice
    img = _float_rows(31 + nx, nx, nrows=24 if nx < 5000 else 12)
    if nx > 20:
        img[9, 5] = np.inf
        img[10, nx - 1] = np.nan
    packed = bbr.fpack_f32(torch.from_numpy(img).cuda(), 16.0, zdither0)
    offs = np.concatenate(([0], np.cumsum(packed['lengths'])[:-1]))
    for r in range(img.shape[0]):
        q, s, z = rice.fpack_quantize_row(img[r], r, 16.0, zdither0)
        assert packed['zscale'][r] == s and packed['zzero'][r] == z, r
        if q is None:
            assert packed['lengths'][r] == 0 and np.array_equal(packed['lossless_rows'][r], img[r], equal_nan=True)
            continue
        assert packed['heap'][offs[r]:offs[r] + packed['lengths'][r]].tobytes() == rice.encode_tile(q, 4), r
    assert len(packed['lossless_rows']) >= (2 if nx >= 9 else img.shape[0])
    path = fitsio.write_compressed(str(tmp_path / 'red.fits.fz'), shape=img.shape, zbitpix=-32,
                                   header={'RDNOISE': 9.5}, **packed)
    hdr, back = bbr.read_fits_image(path)
    back = back.cpu().numpy()
    assert hdr['RDNOISE'] == 9.5
    for r in range(img.shape[0]):
        if r in packed['lossless_rows']:
            assert np.array_equal(back[r], img[r], equal_nan=True)
        else:
            assert np.abs(back[r] - img[r]).max() <= 0.5 * packed['zscale'][r] * (1 + 1e-6) + 1e-4


@pytest.mark.gpu
def test_run_host_writes_the_image_as_fpack_q16(tmp_path, small_bb):
    """BatchReducer.run_host(img_fz=True): what leaves the device is the reference's disk product
    (`fpack -q 16 -D -Y`, blackbox.py:836).  Its rows equal the oracle's quantisation of the float32
    image of the plain path bit for bit, the file reads back to within half a step of a sixteenth of
    the row noise, and the guessed copy size is topped up when it was short."""
    import torch
    from blackbox_b200 import fitsio, reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import BatchReducer
    from oracle import rice
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 53, shape)
    coeffs = synth.make_xtalk(54)[3]
    raws = [synth.make_raw(tel, seed, nstars=100, ncosmics=60)[0] for seed in (52, 53, 54, 55, 56, 57)]
    batch = BatchReducer(tel, raws[0].shape, depth=3, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=3)
    plain_in = [torch.from_numpy(r.view(np.int16)).view(torch.uint16).pin_memory() for r in raws]
    imgs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in raws]
    masks = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in raws]
    want = batch.run_host(plain_in, imgs, masks)
    nbytes = batch.img_fz_bytes(2.0)
    fz = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in raws]
    seeds = [9998 + k for k in range(len(raws))]
    for rep in range(2):
        if rep == 1:
            batch._fz_img_guess = batch.img_fz_bytes(0) + 4096          # far too short: every frame is topped up
        got = batch.run_host(plain_in, fz, masks, img_fz=True, zdither0=seeds)
        assert batch.d2h_bytes - sum(m.numel() for m in masks) < sum(i.numel() * 4 for i in imgs) // 2
        for k in range(len(raws)):
            p = got[k].img_fz
            assert p['zdither0'] == 1 + (seeds[k] - 1) % 10000 and got[k].header == want[k].header
            offs = np.concatenate(([0], np.cumsum(p['lengths'], dtype=np.int64)[:-1]))
            ref = imgs[k].numpy()
            for row in (0, 5, shape[0] // 2, shape[0] - 1):
                q, s, z = rice.fpack_quantize_row(ref[row], row, 16.0, p['zdither0'])
                assert p['zscale'][row] == s and p['zzero'][row] == z, (rep, k, row)
                assert p['heap'][offs[row]:offs[row] + p['lengths'][row]].tobytes() == rice.encode_tile(q, 4), (rep, k, row)
    p = got[3].img_fz
    path = fitsio.write_compressed(str(tmp_path / 'red.fits.fz'), shape=shape, zbitpix=-32,
                                   header={k: (v, '') for k, v in got[3].header.items()}, **p)
    hdr, back = bbr.read_fits_image(path)
    err = np.abs(back.cpu().numpy() - imgs[3].numpy())
    assert (err <= 0.5 * p['zscale'][:, None] * (1 + 1e-6) + 1e-3).all()
    assert os.path.getsize(path) < imgs[3].numel() * 4 // 3
    assert hdr['NCOSMICS'] == want[3].header['NCOSMICS']


@pytest.mark.gpu
def test_reduce_night_tool_with_fpacked_files_on_both_sides(small_bb, tmp_path):
    """tools/reduce_night.py --fpack on a directory of .fits.fz raw frames with fpacked masters:
    _red.fits.fz / _mask.fits.fz come out; the mask equals the oracle's chain, the image is the
    oracle's image to within half a quantisation step (a sixteenth of the row noise), the header
    keywords of the reduction steps are in the files and the raw frame's BZERO is not."""
    import importlib.util
    from blackbox_b200 import fitsio, reduce as bbr, set_bb, synth
    from oracle import reduce as R, rice
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    raw_dir, out_dir = tmp_path / 'raw', tmp_path / 'red'
    raw_dir.mkdir()
    raws = [synth.make_raw(tel, 4800 + k, nstars=100, ncosmics=60)[0] for k in range(5)]
    mbias, mflat, bpm = synth.make_masters(tel, 4800, shape)
    victim, source, corr, coeffs = synth.make_xtalk(4802)
    for k, r in enumerate(raws):
        rice.write_fz(str(raw_dir / 'BG3_2026_{:02d}.fits.fz'.format(k)), r, {'EXPTIME': 45.0 + k, 'FILTER': 'q'})
    fitsio.write_primary(str(tmp_path / 'mbias.fits'), mbias)
    fitsio.write_primary(str(tmp_path / 'mflat.fits'), mflat)
    rice.write_fz_u8(str(tmp_path / 'bpm.fits.fz'), bpm)
    synth.write_xtalk_file(str(tmp_path / 'xtalk.txt'), victim, source, corr)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('reduce_night', os.path.join(here, 'tools', 'reduce_night.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    n = tool.main([str(raw_dir), str(out_dir), '--tel', tel, '--mbias', str(tmp_path / 'mbias.fits'),
                   '--mflat', str(tmp_path / 'mflat.fits'), '--bpm', str(tmp_path / 'bpm.fits.fz'),
                   '--xtalk', str(tmp_path / 'xtalk.txt'), '--niter', '2', '--fpack', '--chunk', '2'])
    assert n == 5
    for k, r in enumerate(raws):
        data_o, mask_o, hdr_o, _ = R.reduce_frame(r, tel, mbias, mflat, bpm, coeffs, niter=2)
        red = str(out_dir / 'BG3_2026_{:02d}_red.fits.fz'.format(k))
        ci = fitsio.read_compressed(red)
        assert ci.info['bitpix'] == -32 and ci.info['zdither0'] == 1 + k and ci.info['bzero'] == 0.0
        h, img = bbr.read_fits_image(red)
        _, m = bbr.read_fits_image(str(out_dir / 'BG3_2026_{:02d}_mask.fits.fz'.format(k)))
        assert np.mean(m.cpu().numpy() != mask_o) <= 1e-5
        err = np.abs(img.cpu().numpy() - data_o)
        assert np.mean(err <= 0.5 * ci.zscale[:, None] * (1 + 1e-6) + 1e-3) > 0.999
        assert h['FILTER'] == 'q' and h['REDFILE'].endswith('_red') and h['EXPTIME'] == 45.0 + k
        assert h['BIASMEAN'] == pytest.approx(hdr_o['BIASMEAN'], rel=1e-9) and h['NOBJ-SAT'] == hdr_o['NOBJ-SAT']


def test_fpack_output_layout_is_what_the_host_side_parses():
    """The size / offset entry points of bbx_fpack_f32 are host code (no GPU): the table columns sit
    where reduce.FpackEncoder.parse looks for them, the always-fitting output size is columns +
    one fixed-stride row per tile, and the scratch is the Rice encoder's for BYTEPIX 4."""
    from blackbox_b200._lib import query
    for H, W in ((1, 1), (7, 33), (240, 10560), (10560, 10560)):
        off = query('bbx_fpack_f32_heap_offset', H)
        assert off == 16 + (4 * H + 15) // 16 * 16 + 2 * ((8 * H + 15) // 16 * 16) and off % 16 == 0
        nblocks = (W + 31) // 32
        stride = (4 + nblocks * 129 + 8 + 15) // 16 * 16
        assert query('bbx_fpack_f32_out_bytes', H, W) == off + H * stride
        assert query('bbx_fpack_f32_work_bytes', H, W) == query('bbx_rice_encode_work_bytes', H, W, 4) >= H * stride
    assert query('bbx_fpack_f32_out_bytes', 0, 5) == 0 and query('bbx_fpack_f32_heap_offset', 0) == 0
