"""fitsio: header parsing / writing and the data unit as stored (CPU); decode / encode on the GPU."""
import numpy as np
import pytest


def _cards(*cards):
    text = ''.join(c.ljust(80) for c in cards) + 'END'.ljust(80)
    text += ' ' * (-len(text) % 2880)
    return text.encode('ascii')


def test_read_primary_parses_cards_and_maps_the_data_unit(tmp_path):
    from blackbox_b200 import fitsio
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 65536, size=(6, 10), dtype=np.uint16)
    stored = (counts.astype(np.int32) - 32768).astype('>i2')
    hdr = _cards('SIMPLE  =                    T / conforms to FITS standard',
                 'BITPIX  =                   16 / array data type',
                 'NAXIS   =                    2', 'NAXIS1  =                   10', 'NAXIS2  =                    6',
                 'BSCALE  =                    1', 'BZERO   =                32768',
                 "OBJECT  = 'O''Brien field'     / target", 'EXPTIME =                 60.5 / [s]',
                 'COMMENT raw frame', "FILTER  = 'q       '")
    path = tmp_path / 'raw.fits'
    payload = stored.tobytes()
    path.write_bytes(hdr + payload + b'\0' * (-len(payload) % 2880))
    header, data, info = fitsio.read_primary(str(path))
    assert info == dict(bitpix=16, shape=(6, 10), bzero=32768.0, bscale=1.0, offset=2880)
    assert header['OBJECT'] == ("O'Brien field", 'target')
    assert header['EXPTIME'] == (60.5, '[s]') and header['FILTER'][0] == 'q'
    assert header['COMMENT'][0] == ['raw frame']
    assert data.dtype == np.dtype('>i2') and np.array_equal(data, stored)
    assert np.array_equal(fitsio.to_native(data, info), counts)


def test_write_primary_round_trip(tmp_path):
    from blackbox_b200 import fitsio
    rng = np.random.default_rng(2)
    img = rng.standard_normal((7, 12)).astype(np.float32)
    mask = rng.integers(0, 128, size=(7, 12), dtype=np.uint8)
    raw = rng.integers(0, 65536, size=(7, 12), dtype=np.uint16)
    hdr = {'GAIN1': (2.614, '[e-/ADU] gain applied to channel 1'), 'NOBJ-SAT': 12, 'XTALK-P': True,
           'REDFILE': 'BG3_20260101_red', 'BIASMEAN': 3208.809706141008, 'COMMENT': ['a', 'b']}
    for name, arr in (('img', img), ('mask', mask), ('raw', raw)):
        path = str(tmp_path / (name + '.fits'))
        fitsio.write_primary(path, arr, hdr)
        assert (tmp_path / (name + '.fits')).stat().st_size % 2880 == 0
        h, data, info = fitsio.read_primary(path)
        assert np.array_equal(fitsio.to_native(data, info) if name != 'mask' else np.asarray(data), arr)
        assert h['GAIN1'] == hdr['GAIN1'] and h['NOBJ-SAT'][0] == 12 and h['XTALK-P'][0] is True
        assert h['REDFILE'][0] == 'BG3_20260101_red' and h['BIASMEAN'][0] == hdr['BIASMEAN']
        assert h['COMMENT'][0] == ['a', 'b']
    # big-endian bytes handed in directly (what reduce.fits_encode produces on the GPU)
    path = str(tmp_path / 'be.fits')
    fitsio.write_primary(path, img.astype('>f4').view(np.uint8).reshape(-1), hdr, be_bytes=True, shape=img.shape, bitpix=-32)
    _, data, info = fitsio.read_primary(path)
    assert np.array_equal(fitsio.to_native(data, info), img)


def test_read_primary_rejects_what_it_cannot_handle(tmp_path):
    from blackbox_b200 import fitsio
    bad = tmp_path / 'cube.fits'
    bad.write_bytes(_cards('SIMPLE  =                    T', 'BITPIX  =                    8', 'NAXIS   =                    0'))
    with pytest.raises(fitsio.FitsError):
        fitsio.read_primary(str(bad))
    trunc = tmp_path / 'trunc.fits'
    trunc.write_bytes(_cards('SIMPLE  =                    T', 'BITPIX  =                  -32', 'NAXIS   =                    2',
                             'NAXIS1  =                  100', 'NAXIS2  =                  100') + b'\0' * 2880)
    with pytest.raises(fitsio.FitsError):
        fitsio.read_primary(str(trunc))


@pytest.mark.gpu
def test_fits_decode_encode_on_the_gpu(tmp_path):
    """raw uint16 (BZERO 32768) and float32 data units: GPU decode of the bytes on disk equals the
    host decode; GPU encode writes a file that reads back to the same array."""
    import torch
    from blackbox_b200 import fitsio, reduce as bbr
    rng = np.random.default_rng(3)
    for shape in ((53, 60), (64, 128), (1, 7)):
        raw = rng.integers(0, 65536, size=shape, dtype=np.uint16)
        raw[0, 0], raw[-1, -1] = 0, 65535
        img = rng.standard_normal(shape).astype(np.float32)
        p_raw, p_img = str(tmp_path / 'r.fits'), str(tmp_path / 'i.fits')
        fitsio.write_primary(p_raw, raw, {'EXPTIME': 60.0})
        fitsio.write_primary(p_img, img)
        for path, want in ((p_raw, raw), (p_img, img)):
            _, buf, info = fitsio.read_primary(path, pinned=True)
            got = bbr.fits_decode(buf, info)
            assert tuple(got.shape) == shape
            if want.dtype == np.uint16:
                assert got.dtype == torch.uint16
                assert np.array_equal(got.view(torch.int16).cpu().numpy().view(np.uint16), want)
            else:
                assert np.array_equal(got.cpu().numpy(), want)
            be, bitpix = bbr.fits_encode(got)
            out = str(tmp_path / 'o.fits')
            fitsio.write_primary(out, be.cpu(), be_bytes=True, shape=shape, bitpix=bitpix,
                                 bzero=32768 if want.dtype == np.uint16 else None)
            _, data, info2 = fitsio.read_primary(out)
            assert np.array_equal(fitsio.to_native(data, info2), want)
