"""Tile-compressed (.fits.fz, RICE_1) raw frames: the oracle codec against hand-derived vectors,
the host parser of blackbox_b200.fitsio, and the GPU decoder (bbx_rice_decode16)."""
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
