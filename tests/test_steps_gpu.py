"""GPU parity of the single reduction steps against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rng_img(rng, shape, level=300.0):
    return (level + 20 * rng.standard_normal(shape)).astype(np.float32)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('k', [3, 5, 7])
def test_medfilt_bit_exact(k):
    import torch
    from blackbox_b200 import _lib, reduce as bbr
    from oracle import clib
    rng = np.random.default_rng(k)
    img = _rng_img(rng, (97, 203))
    img[10:14, 50:60] = 5000.0
    img[rng.random(img.shape) < 0.02] = 0.0          # ties
    ref = {3: clib.medfilt3, 5: clib.medfilt5, 7: clib.medfilt7}[k](img)
    t = torch.from_numpy(img).cuda()
    out = torch.empty_like(t)
    _lib.call('bbx_medfilt', bbr._ptr(t), bbr._ptr(out), img.shape[0], img.shape[1], k, bbr._stream())
    assert np.array_equal(out.cpu().numpy(), ref)


def test_medfilt_tiny_images():
    import torch
    from blackbox_b200 import _lib, reduce as bbr
    from oracle import clib
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (2, 9), (5, 5), (6, 4), (7, 7), (8, 33)]:
        img = _rng_img(rng, shape)
        t = torch.from_numpy(img).cuda()
        for k, ref in ((3, clib.medfilt3), (5, clib.medfilt5), (7, clib.medfilt7)):
            out = torch.empty_like(t)
            _lib.call('bbx_medfilt', bbr._ptr(t), bbr._ptr(out), shape[0], shape[1], k, bbr._stream())
            assert np.array_equal(out.cpu().numpy(), ref(img)), (shape, k)


def test_laplace_plus_bit_exact():
    import torch
    from blackbox_b200 import _lib, reduce as bbr
    from oracle import clib, lacosmic
    rng = np.random.default_rng(3)
    for shape in [(1, 1), (1, 7), (9, 1), (64, 131)]:
        img = _rng_img(rng, shape)
        ref_c = clib.rebin(np.maximum(clib.laplace(clib.subsample(img)), 0))
        assert np.array_equal(ref_c, lacosmic.laplace_plus(img))
        t = torch.from_numpy(img).cuda()
        out = torch.empty_like(t)
        _lib.call('bbx_laplace_plus', bbr._ptr(t), bbr._ptr(out), shape[0], shape[1], bbr._stream())
        assert np.array_equal(out.cpu().numpy(), ref_c), shape


def test_masked_lower_median():
    import torch
    from blackbox_b200 import _lib, reduce as bbr
    from oracle import clib
    rng = np.random.default_rng(4)
    for n, frac in [(1, 0.0), (2, 0.0), (1001, 0.3), (250000, 0.1), (64, 1.0)]:
        img = _rng_img(rng, (n,))
        img[::7] = -img[::7]
        mask = (rng.random(n) < frac).astype(np.uint8)
        t, m = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
        work = torch.empty(_lib.query('bbx_select_work_bytes'), dtype=torch.uint8, device='cuda')
        out = torch.zeros(1, dtype=torch.float32, device='cuda')
        _lib.call('bbx_masked_lower_median', bbr._ptr(t), bbr._ptr(m), n, bbr._ptr(work), bbr._ptr(out), bbr._stream())
        good = img[mask == 0]
        want = clib.lower_median(good) if good.size else np.float32(0)
        assert out.item() == want, (n, frac)


# ------------------------------------------------------------------------------------------
def _lacosmic_case(seed, shape=(160, 232), ncr=60, masked_frac=0.01):
    rng = np.random.default_rng(seed)
    img = (300 + np.sqrt(300) * rng.standard_normal(shape)).astype(np.float32)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    for _ in range(25):                                     # stars
        y0, x0, f = rng.uniform(0, shape[0]), rng.uniform(0, shape[1]), rng.uniform(2e3, 2e5)
        img += (f / (2 * np.pi * 1.5 ** 2) * np.exp(-((yy - y0) ** 2 + (xx - x0) ** 2) / (2 * 1.5 ** 2))).astype(np.float32)
    for _ in range(ncr):                                    # cosmic rays, also on the borders
        y, x = rng.integers(0, shape[0]), rng.integers(0, shape[1])
        for t in range(rng.integers(1, 8)):
            yy_, xx_ = min(y + t // 2, shape[0] - 1), min(x + t, shape[1] - 1)
            img[yy_, xx_] += rng.uniform(500, 30000)
    mask = rng.random(shape) < masked_frac
    mask[40:60, 100:104] = True
    return img, mask


@pytest.mark.parametrize('mode', ['lazy', 'dense'])
@pytest.mark.parametrize('seed,niter,sigclip', [(1, 4, 15), (2, 3, 20), (3, 1, 4.5), (4, 4, 6)])
def test_detect_cosmics_bit_exact(seed, niter, sigclip, mode):
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic
    img, mask = _lacosmic_case(seed)
    kw = dict(sigclip=sigclip, sigfrac=0.01 if seed != 4 else 0.3, objlim=3, niter=niter, readnoise=8.5,
              gain=1.0, satlevel=np.inf, cleantype='medmask', sepmed=False)
    info_o, info_g = {}, {}
    cr_o, clean_o = lacosmic.detect_cosmics(img, inmask=mask, info=info_o, **kw)
    cr_g, clean_g = bbr.detect_cosmics(img, inmask=mask, info=info_g,
                                       mode=bbr.LAC_LAZY if mode == 'lazy' else bbr.LAC_DENSE, **kw)
    assert cr_o.sum() > 50
    assert np.array_equal(cr_g, cr_o)
    assert np.array_equal(clean_g.view(np.uint32), clean_o.view(np.uint32))
    assert info_g['iterations'] == info_o['iterations']
    assert np.array_equal(info_g['ncr_per_iter'], info_o['ncr_per_iter'])


@pytest.mark.parametrize('sigclip,sigfrac,niter', [(5.0, 0.01, 5), (12.0, 0.3, 4)])
def test_detect_cosmics_many_hits_lazy_equals_dense_equals_oracle(sigclip, sigfrac, niter):
    """A busy 1500 x 2000 frame (thousands of overlapping cosmic-ray tracks, low thresholds, more
    iterations than flag stamps): the work lists of the lazy path are long, neighbourhoods of
    different candidates overlap heavily and the growth steps meet every pixel from several
    sides -- still the dense twin's and the oracle's bits, run after run."""
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic
    img, mask = _lacosmic_case(11, shape=(1500, 2000), ncr=6000, masked_frac=0.02)
    kw = dict(sigclip=sigclip, sigfrac=sigfrac, objlim=2, niter=niter, readnoise=8.5, gain=1.0,
              satlevel=np.inf, cleantype='medmask', sepmed=False)
    info_o = {}
    cr_o, clean_o = lacosmic.detect_cosmics(img, inmask=mask, info=info_o, **kw)
    assert cr_o.sum() > 10000
    for mode in (bbr.LAC_LAZY, bbr.LAC_LAZY, bbr.LAC_DENSE, None):
        info_g = {}
        cr_g, clean_g = bbr.detect_cosmics(img, inmask=mask, info=info_g, mode=mode, **kw)
        assert np.array_equal(info_g['ncr_per_iter'], info_o['ncr_per_iter']), mode
        assert np.array_equal(cr_g, cr_o), mode
        assert np.array_equal(clean_g.view(np.uint32), clean_o.view(np.uint32)), mode


def test_detect_cosmics_background_level():
    """A fat cosmic-ray blob leaves interior pixels without usable neighbours: they get the
    global background level (lower median of all unmasked input pixels).  The lazy path finds
    it from a sampled bracket + one histogram bin per float32 key inside the bracket (mode 0),
    or from the dense radix select run up front (mode 2)."""
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic
    rng = np.random.default_rng(21)
    img = (300 + 17 * rng.standard_normal((96, 120))).astype(np.float32)
    img[40:53, 50:63] += rng.uniform(20000, 60000, (13, 13)).astype(np.float32)
    img[10, 10] += 9000.0
    mask = np.zeros(img.shape, bool)
    mask[:, :7] = True
    kw = dict(sigclip=15, sigfrac=0.01, objlim=3, niter=4, readnoise=8.5, gain=1.0,
              satlevel=np.inf, cleantype='medmask', sepmed=False)
    for m in (None, mask):
        info_o, info_g = {}, {}
        cr_o, clean_o = lacosmic.detect_cosmics(img, inmask=m, info=info_o, **kw)
        assert (clean_o[cr_o] == info_o['background']).sum() > 3     # the case is exercised
        for mode in (bbr.LAC_LAZY, bbr.LAC_LAZY_BG, bbr.LAC_DENSE, None):
            cr_g, clean_g = bbr.detect_cosmics(img, inmask=m, info=info_g, mode=mode, **kw)
            assert np.array_equal(cr_g, cr_o) and np.array_equal(clean_g, clean_o)
            assert info_g['lazy_status'] == 0
            assert info_g['iterations'] == info_o['iterations']


def test_detect_cosmics_background_bracket_miss():
    """Two pixel populations of equal size (sky ~300, nebula ~30000): the median sits at the
    boundary, the sampled bracket spans far more float32 keys than the histogram has bins, so
    mode 0 must report that it needs the level (never a wrong value) and the automatic mode
    must repeat with mode 2."""
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic
    rng = np.random.default_rng(22)
    img = (300 + 17 * rng.standard_normal((96, 120))).astype(np.float32)
    img[:, 60:] += 30000
    img[:, 60:] += (170 * rng.standard_normal((96, 60))).astype(np.float32)
    img[40:53, 20:33] += rng.uniform(20000, 60000, (13, 13)).astype(np.float32)
    mask = None
    kw = dict(sigclip=15, sigfrac=0.01, objlim=3, niter=4, readnoise=8.5, gain=1.0,
              satlevel=np.inf, cleantype='medmask', sepmed=False)
    info_o, info_g = {}, {}
    cr_o, clean_o = lacosmic.detect_cosmics(img, inmask=mask, info=info_o, **kw)
    assert (clean_o[cr_o] == info_o['background']).sum() > 3
    with pytest.raises(RuntimeError):
        bbr.detect_cosmics(img, inmask=mask, mode=bbr.LAC_LAZY, **kw)
    cr_g, clean_g = bbr.detect_cosmics(img, inmask=mask, info=info_g, **kw)
    assert info_g['lazy_status'] == bbr.LAC_STATUS_NEED_BG
    assert np.array_equal(cr_g, cr_o) and np.array_equal(clean_g, clean_o)


def test_detect_cosmics_no_mask_and_early_stop():
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic
    rng = np.random.default_rng(9)
    img = (100 + rng.standard_normal((64, 80))).astype(np.float32)     # nothing to find
    kw = dict(sigclip=15, sigfrac=0.01, objlim=3, niter=4, readnoise=5.0, gain=1.0,
              satlevel=np.inf, cleantype='medmask', sepmed=False)
    info_o, info_g = {}, {}
    cr_o, clean_o = lacosmic.detect_cosmics(img, info=info_o, **kw)
    cr_g, clean_g = bbr.detect_cosmics(img, info=info_g, **kw)
    assert not cr_o.any() and not cr_g.any()
    assert info_g['iterations'] == info_o['iterations'] == 1
    assert np.array_equal(clean_g, clean_o)


def test_detect_cosmics_unsupported_modes_raise():
    from blackbox_b200 import reduce as bbr
    img = np.zeros((16, 16), np.float32)
    with pytest.raises(NotImplementedError):
        bbr.detect_cosmics(img)                                    # astroscrappy defaults
    with pytest.raises(NotImplementedError):
        bbr.detect_cosmics(img, sepmed=False, cleantype='medmask', satlevel=5e4)


# ------------------------------------------------------------------------------------------
def _mask_case(seed, shape=(2 * 96, 8 * 132)):
    rng = np.random.default_rng(seed)
    data = _rng_img(rng, shape)
    sat = 2.5e5
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    # rings (holes), a diagonal-gap ring (leaks), blobs, a bleed column, border contact
    for (cy, cx, r0, r1) in [(30, 100, 5, 8), (140, 700, 9, 11), (60, 400, 3, 4)]:
        rr = np.hypot(yy - cy, xx - cx)
        data[(rr >= r0) & (rr <= r1)] = sat
    data[100:108, 300] = sat; data[100, 300:308] = sat; data[108, 300:309] = sat; data[100:108, 308] = sat
    data[101, 309] = sat
    data[20:25, 900:905] = sat
    data[0:3, 50:53] = sat                      # touches the image border
    data[shape[0] - 1, 200:203] = sat
    data[10:90, 555] = sat                      # bleed trail
    data[50, 10] = np.nan
    data[51, 11] = np.inf
    # two blobs separated by a one-pixel gap: closing joins them
    data[160:164, 100:104] = sat; data[160:164, 105:109] = sat
    bpm = np.zeros(shape, np.uint8)
    bpm[rng.random(shape) < 0.002] = 1
    bpm[:4, :] = 32; bpm[-4:, :] = 32; bpm[:, :4] = 32; bpm[:, -4:] = 32
    return data, bpm


@pytest.mark.parametrize('sparse', [True, False])
@pytest.mark.parametrize('tel', ['ML1', 'BG3'])
def test_mask_init_bit_exact(tel, sparse, small_bb, monkeypatch):
    import torch
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    monkeypatch.setattr(bbr, 'MASK_MORPH_SPARSE', sparse)
    small_bb(96, 132)
    data, bpm = _mask_case(5)
    bpm[120:123, 500] = 4            # a bad-pixel mask that already carries saturated bits
    bpm[130, 510:512] = 8            # ... and saturated-connected ones: part of the closing
    bpm[130, 513] = 8                # one-pixel gap at column 512: closed, filled if unmasked
    hdr = {'BIASM{}'.format(i + 1): 6500.0 + i for i in range(16)}
    hdr_o, hdr_g = dict(hdr), dict(hdr)
    d_o = data.copy()
    mask_o, hm_o = R.mask_init(d_o, hdr_o, bpm, 'object', tel=tel)
    bbr.tel = tel
    d_g = data.copy()
    mask_g, hm_g = bbr.mask_init(d_g, hdr_g, 'q', 'object', bpm=bpm)
    assert np.array_equal(mask_g, mask_o), np.argwhere(mask_g != mask_o)[:10]
    assert np.array_equal(d_g, d_o)
    assert hdr_g['NOBJ-SAT'] == hdr_o['NOBJ-SAT'] > 5
    assert hm_g['SATURATE'] == hm_o['SATURATE']
    for i in range(16):
        assert hdr_g['SATLEV{}'.format(i + 1)] == hdr_o['SATLEV{}'.format(i + 1)]
    # torch in, torch out, nothing else changes
    d_t = torch.from_numpy(data.copy()).cuda()
    mask_t, _ = bbr.mask_init(d_t, dict(hdr), 'q', 'object', bpm=torch.from_numpy(bpm).cuda())
    assert np.array_equal(mask_t.cpu().numpy(), mask_o)
    assert np.array_equal(d_t.cpu().numpy(), d_o)


def test_mask_init_no_saturation_and_no_bpm(small_bb):
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    small_bb(96, 132)
    rng = np.random.default_rng(1)
    data = _rng_img(rng, (192, 1056))
    hdr = {'BIASM{}'.format(i + 1): 6500.0 for i in range(16)}
    bbr.tel = 'ML1'
    mask_o, _ = R.mask_init(data.copy(), dict(hdr), None, 'object', tel='ML1')
    mask_g, _ = bbr.mask_init(data.copy(), dict(hdr), 'q', 'object')
    assert not mask_o.any() and np.array_equal(mask_g, mask_o)


def test_mask_init_reads_the_bad_pixel_mask_like_the_reference(small_bb, tmp_path, monkeypatch, caplog):
    """The reference-style call mask_init(data, header, filt, imgtype) -- no mask passed in -- reads
    set_bb.bad_pixel_mask with 'bpm' -> 'bpm_<filt>' from disk (blackbox.py:4386-4398); a missing
    file is a warning and a mask of zeros, never silent."""
    import logging
    from blackbox_b200 import fitsio, reduce as bbr, set_bb
    from oracle import reduce as R
    small_bb(96, 132)
    data, bpm = _mask_case(9)
    hdr = {'BIASM{}'.format(i + 1): 6500.0 + i for i in range(16)}
    monkeypatch.setattr(set_bb, 'bad_pixel_mask', {'ML1': str(tmp_path / 'ML1_bpm_0p2.fits')})
    bbr._bpm_registry.clear()
    bbr.tel = 'ML1'
    fitsio.write_primary(str(tmp_path / 'ML1_bpm_q_0p2.fits'), bpm)
    mask_o, _ = R.mask_init(data.copy(), dict(hdr), bpm, 'object', tel='ML1')
    mask_g, _ = bbr.mask_init(data.copy(), dict(hdr), 'q', 'object')
    assert np.array_equal(mask_g, mask_o) and (mask_g & 32).any()
    # a second call is served from the device cache; a rewritten file is read again
    mask_g2, _ = bbr.mask_init(data.copy(), dict(hdr), 'q', 'object')
    assert np.array_equal(mask_g2, mask_o)
    # another filter: no file -> warning + zeros (as the reference)
    with caplog.at_level(logging.WARNING, logger='blackbox_b200.reduce'):
        mask_u, _ = bbr.mask_init(data.copy(), dict(hdr), 'u', 'object')
    assert any('does not exist' in r.getMessage() for r in caplog.records)
    mask_z, _ = R.mask_init(data.copy(), dict(hdr), None, 'object', tel='ML1')
    assert np.array_equal(mask_u, mask_z) and not (mask_u & 32).any()


def test_in_place_steps_refuse_tensors_they_would_have_to_copy(small_bb):
    """xtalk_corr / gain_corr / mask_init / fill_edge_pixels work in place: a CUDA tensor that is not
    contiguous float32 would be silently copied and the result lost -- they raise instead."""
    import torch
    from blackbox_b200 import reduce as bbr
    small_bb(24, 132)
    bbr.tel = 'ML1'
    big = torch.zeros((48, 2 * 1056), dtype=torch.float32, device='cuda')
    view = big[:, :1056]                                           # a non-contiguous view
    with pytest.raises(TypeError):
        bbr.xtalk_corr(view, np.zeros((16, 16)))
    with pytest.raises(TypeError):
        bbr.gain_corr(view, {}, tel='ML1')
    with pytest.raises(TypeError):
        bbr.xtalk_corr(torch.zeros((48, 1056), dtype=torch.float64, device='cuda'), np.zeros((16, 16)))
    with pytest.raises(TypeError):
        bbr.os_corr(torch.zeros((48 + 40, 12000), dtype=torch.int16, device='cuda'), {}, 'object', tel='ML1')


def test_mask_header_counts(small_bb):
    from blackbox_b200 import reduce as bbr
    rng = np.random.default_rng(2)
    mask = rng.integers(0, 128, (50, 77)).astype(np.uint8)
    hm = {}
    bbr.tel = 'ML1'
    bbr.mask_header(mask, hm)
    for name, short in [('bad', 'BP'), ('edge', 'EP'), ('saturated', 'SP'), ('saturated-connected', 'SCP'),
                        ('satellite trail', 'STP'), ('cosmic ray', 'CRP')]:
        from blackbox_b200 import set_bb
        v = set_bb.mask_value[name]
        assert hm['M-{}NUM'.format(short)] == int(np.sum(mask & v == v))


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('with_mask', [True, False])
def test_xtalk_parity(with_mask, small_bb, tmp_path):
    from blackbox_b200 import reduce as bbr, synth
    from oracle import reduce as R
    small_bb(64, 132)
    rng = np.random.default_rng(6)
    shape = (128, 1056)
    data = _rng_img(rng, shape, level=50.0)
    data[rng.random(shape) < 0.01] += 60000.0
    mask = None
    if with_mask:
        mask = rng.choice(np.array([0, 0, 0, 1, 2, 32, 64, 33], np.uint8), size=shape)
    v, s, c, coeffs = synth.make_xtalk(11, amp=3e-4)
    path = tmp_path / 'xtalk.txt'
    synth.write_xtalk_file(str(path), v, s, c)
    assert np.array_equal(bbr.read_crosstalk_file(str(path)), coeffs)
    d_o = data.copy()
    R.xtalk_corr(d_o, coeffs, None if mask is None else mask.copy(), tel='ML1')
    bbr.tel = 'ML1'
    d_g = data.copy()
    m_g = None if mask is None else mask.copy()
    bbr.xtalk_corr(d_g, str(path), m_g)
    if mask is not None:
        assert np.array_equal(m_g, mask)
    assert np.mean(d_g == d_o) > 0.9999
    np.testing.assert_allclose(d_g, d_o, rtol=1e-6, atol=0)
    assert np.abs(d_o - data).max() > 1.0          # the correction did something


@pytest.mark.parametrize('shape', [(128, 1056), (2 * 77, 8 * 1320), (2 * 40, 8 * 660), (2 * 33, 8 * 100), (2 * 9, 8 * 30)])
def test_xtalk_kernels_agree_and_count_the_mask(shape):
    """Every crosstalk kernel -- the tile kernel bbx_xtalk runs, the TMA-staged persistent one (boxes
    of 128 positions, clipped at the channel edge), the generic ones -- gives
    the same bits as the oracle class allows (> 99.99 % identical, 1e-6), and identical bits among
    themselves; the per-bit mask counts taken on the way equal numpy's (blackbox.py:4601-4620)."""
    import ctypes as C
    import torch
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200._lib import call
    from oracle import reduce as R
    H, W = shape
    saved = (set_bb.ysize_chan, set_bb.xsize_chan)
    set_bb.ysize_chan, set_bb.xsize_chan = H // 2, W // 8
    try:
        rng = np.random.default_rng(H + W)
        data = _rng_img(rng, shape, level=50.0)
        data[rng.random(shape) < 0.01] += 60000.0
        data[rng.random(shape) < 0.01] = -3.0
        mask = rng.choice(np.array([0, 0, 0, 1, 2, 4, 8, 32, 64, 33, 12], np.uint8), size=shape)
        coeffs = np.ascontiguousarray(synth.make_xtalk(11, amp=3e-4)[3], dtype=np.float64)
        d_o = data.copy()
        R.xtalk_corr(d_o, coeffs, mask.copy(), tel='BG3')
        bits = bbr._bits('BG3')
        m_t = torch.from_numpy(mask).cuda()
        outs = {}
        for variant in (0, 3, 5, 4, 2, 1):
            img = torch.from_numpy(data.copy()).cuda()
            counts = torch.full((136,), -1, dtype=torch.int64, device='cuda')
            call('bbx_xtalk_counts', bbr._ptr(img), bbr._ptr(m_t), H, W, H // 2, W // 8,
                 coeffs.ctypes.data_as(C.c_void_p), C.byref(bits), variant, bbr._ptr(counts), bbr._stream())
            outs[variant] = img.cpu().numpy()
            assert counts[:8].cpu().tolist() == [int(((mask >> b) & 1).sum()) for b in range(8)], variant
        assert torch.equal(m_t.cpu(), torch.from_numpy(mask))
        for variant, got in outs.items():
            assert np.array_equal(got, outs[1]), variant
        assert np.mean(outs[0] == d_o) > 0.9999
        np.testing.assert_allclose(outs[0], d_o, rtol=1e-6, atol=0)
    finally:
        set_bb.ysize_chan, set_bb.xsize_chan = saved


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 3, 15, 20, 33, 50, 64])
def test_stack_median_bit_exact(n):
    from blackbox_b200 import reduce as bbr
    rng = np.random.default_rng(n)
    frames = [_rng_img(rng, (33, 129)) for _ in range(n)]
    frames[0][3, 5] = -1e30
    if n > 2:
        frames[1][7, 7] = np.nan
        frames[2][8, 8] = np.inf
    want = np.median(np.stack(frames), axis=0)
    got, _ = bbr.master_combine(frames, 'bias')
    assert got.dtype == np.float32
    assert np.array_equal(got, want, equal_nan=True)


def test_master_flat_combine(small_bb):
    from blackbox_b200 import reduce as bbr, set_bb
    from oracle import reduce as R
    small_bb(96, 132)
    set_bb_sec = dict(set_bb.flat_norm_sec)
    set_bb.flat_norm_sec['ML1'] = (slice(20, 90), slice(100, 500))
    try:
        rng = np.random.default_rng(8)
        shape = (192, 1056)
        frames = [(_rng_img(rng, shape, level=20000.0 * rng.uniform(0.8, 1.2))) for _ in range(15)]
        frames[3][100, 100] = -5.0
        for f in frames:
            f[50, 60] = -1.0                        # non-positive median -> 1
        bpm = np.zeros(shape, np.uint8)
        bpm[:5, :] = 32
        bpm[40, 40] = 33                            # not == edge: untouched
        medsec = [None] * 15
        medsec[4] = 19876.5
        want, sc_o = R.master_median(frames, 'flat', medsec=medsec, bpm=bpm, tel='ML1')
        got, sc_g = bbr.master_combine(frames, 'flat', medsec=medsec, bpm=bpm, tel='ML1')
        assert np.array_equal(np.float32(sc_g), np.float32(sc_o))
        assert np.array_equal(got, want)
        assert got[50, 60] == 1.0 and (got[:5] == 1.0).all()
    finally:
        set_bb.flat_norm_sec.update(set_bb_sec)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('as_tensor', [False, True])
def test_cosmics_corr_parity(as_tensor, small_bb):
    import torch
    from blackbox_b200 import reduce as bbr, set_bb
    from oracle import reduce as R
    small_bb(80, 116)
    img, mask = _lacosmic_case(11, shape=(160, 928))
    data_mask = np.where(mask, 1, 0).astype(np.uint8)
    data_mask[:3, :] |= 32
    hdr_o = {'RDNOISE': 8.5, 'EXPTIME': 30.0}
    hdr_g = dict(hdr_o)
    hm_o, hm_g = {}, {}
    d_o, m_o = R.cosmics_corr(img.copy(), hdr_o, data_mask.copy(), hm_o, tel='ML1')
    bbr.tel = 'ML1'
    if as_tensor:
        d_in, m_in = torch.from_numpy(img.copy()).cuda(), torch.from_numpy(data_mask.copy()).cuda()
        d_g, m_g = bbr.cosmics_corr(d_in, hdr_g, m_in, hm_g)
        assert m_g.data_ptr() == m_in.data_ptr()
        d_g, m_g = d_g.cpu().numpy(), m_g.cpu().numpy()
    else:
        d_g, m_g = bbr.cosmics_corr(img.copy(), hdr_g, data_mask.copy(), hm_g)
    assert (m_o & set_bb.mask_value['cosmic ray']).sum() > 0
    assert np.array_equal(m_g, m_o) and np.array_equal(d_g, d_o)
    assert hdr_g['NCOSMICS'] == hdr_o['NCOSMICS'] == hm_g['NCOSMICS'] > 0


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('ysc,xsc', [(33, 41), (64, 40), (7, 4)])
def test_channel_medians_and_edge_fill_bit_exact(ysc, xsc, small_bb):
    """np.median per channel (odd and even pixel counts, ties, negative values, infinities) and
    the edge-pixel fill of blackbox.py:1958-1974 against the oracle."""
    import torch
    from blackbox_b200 import reduce as bbr, set_bb
    from oracle import reduce as R
    small_bb(ysc, xsc)
    rng = np.random.default_rng(ysc * 100 + xsc)
    H, W = 2 * ysc, 8 * xsc
    data = (300 + 30 * rng.standard_normal((H, W))).astype(np.float32)
    data[:ysc, :xsc] = np.round(data[:ysc, :xsc])               # channel 1: many ties
    data[:ysc, xsc:2 * xsc] -= 310                              # channel 2: both signs
    data[ysc:, :xsc][::3, ::2] = np.inf                         # channel 9: infinities
    data[ysc:, xsc:2 * xsc] = 5.0                               # channel 10: constant
    mask = np.zeros((H, W), dtype=np.uint8)
    e = set_bb.mask_value['edge']
    mask[:2, :] = e
    mask[:, -3:] = e | set_bb.mask_value['bad']
    mask[rng.random((H, W)) < 0.01] |= set_bb.mask_value['cosmic ray']
    want = data.copy()
    meds_o = R.fill_edge_pixels(want, mask, tel='BG3')
    bbr.tel = 'BG3'
    meds = bbr.channel_medians(data).cpu().numpy()
    assert np.array_equal(meds.view(np.uint32), meds_o.view(np.uint32))
    got = torch.from_numpy(data.copy()).cuda()
    bbr.fill_edge_pixels(got, torch.from_numpy(mask).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    data_np = data.copy()
    bbr.fill_edge_pixels(data_np, mask)                         # numpy in, mutated in place
    assert np.array_equal(data_np, want)
    # a NaN makes its channel's median NaN (np.median), the others are unaffected
    data[ysc + 1, 3 * xsc + 1] = np.nan
    meds2 = bbr.channel_medians(data).cpu().numpy()
    assert np.isnan(meds2[8 + 3]) and np.array_equal(np.delete(meds2, 11), np.delete(meds, 11))
    # np.nanmedian semantics (get_flatstats): NaNs do not count; an all-NaN channel gives NaN
    data[:ysc, 5 * xsc:6 * xsc][::2] = np.nan
    data[ysc:, 7 * xsc:8 * xsc] = np.nan
    meds3 = bbr.channel_medians(data, ignore_nan=True).cpu().numpy()
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            want3 = np.array([np.nanmedian(data[(i // 8) * ysc:(i // 8 + 1) * ysc, (i % 8) * xsc:(i % 8 + 1) * xsc])
                              for i in range(16)], dtype=np.float32)
    assert np.isnan(meds3[15]) and np.array_equal(meds3, want3, equal_nan=True)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,sigma,maxiters', [(1, 3.0, 5), (2, 3.0, 5), (5, 2.0, 5), (20, 3.0, 5), (20, 2.5, 1),
                                               (33, 3.0, 5), (64, 3.0, 3)])
def test_clipped_stack_median_bit_exact(n, sigma, maxiters):
    """The optional sigma-clipped combine against astropy's sigma_clip + np.ma.median as the
    oracle restates them: outliers, ties, constant stacks, NaN / inf inputs, fully rejected pixels."""
    import torch
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    rng = np.random.default_rng(100 + n)
    shape = (37, 53)
    frames = [(1000 + 10 * rng.standard_normal(shape)).astype(np.float32) for _ in range(n)]
    for k in range(0, n, 4):                                     # outliers (cosmic rays in single frames)
        hit = rng.random(shape) < 0.05
        frames[k][hit] += rng.uniform(100, 5000, hit.sum()).astype(np.float32)
    for f in frames:
        f[0, :10] = 7.0                                          # a constant stack (std = 0)
        f[1, :10] = np.round(f[1, :10])                          # ties
    frames[0][2, 0] = np.nan
    frames[n // 2][2, 1] = np.inf
    frames[n - 1][2, 2] = -np.inf
    for f in frames:
        f[3, 0] = np.nan                                         # nothing survives
    want = R.master_median_clipped(frames, sigma=sigma, maxiters=maxiters)
    got, _ = bbr.master_combine([torch.from_numpy(f).cuda() for f in frames], 'bias', clip_sigma=sigma,
                                clip_maxiters=maxiters)
    got = got.cpu().numpy()
    assert np.isnan(want[3, 0]) and np.isnan(got[3, 0])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or np.array_equal(got, want, equal_nan=True)
    # the plain median stays the default and differs where outliers matter
    plain, _ = bbr.master_combine([torch.from_numpy(f).cuda() for f in frames], 'bias')
    assert np.array_equal(plain.cpu().numpy(), np.median(np.stack(frames), axis=0), equal_nan=True)


# ------------------------------------------------------------------------------------------
def _nonlin_splines(seed):
    from scipy import interpolate
    rng = np.random.default_rng(seed)
    splines = []
    for i in range(16):
        x = np.linspace(0, 60000, 120)
        y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
        k = (3, 3, 2, 5)[i % 4]
        splines.append(interpolate.UnivariateSpline(x, y, k=k, s=x.size * (2e-4) ** 2))
    return splines


def test_nonlin_corr_parity(small_bb):
    """nonlin_corr (blackbox.py:7392-7437) against the reference's own arithmetic on scipy spline
    objects: bit-exact, including the counts above 50000 (divided by 2, as the reference does) and
    negative / NaN pixels."""
    import pickle
    import torch
    from blackbox_b200 import reduce as bbr, set_bb
    from oracle import reduce as R
    small_bb(96, 200)
    rng = np.random.default_rng(9)
    H, W = 2 * 96, 8 * 200
    data = rng.uniform(-500, 140000, size=(H, W)).astype(np.float32)       # electrons; counts up to ~53000
    data[5, 7], data[100, 900] = np.nan, 0.0
    splines = _nonlin_splines(4)
    want = R.nonlin_corr(data.copy(), splines, tel='BG3')
    bbr.tel = 'BG3'
    got = bbr.nonlin_corr(torch.from_numpy(data.copy()).cuda(), splines).cpu().numpy()
    assert (data / np.float32(set_bb.gain['BG3'][0]) > 50000).sum() > 100
    assert np.array_equal(got, want, equal_nan=True)
    got_np = bbr.nonlin_corr(data.copy(), splines)                         # numpy in, list in
    assert np.array_equal(got_np, want, equal_nan=True)


def test_nonlin_corr_reads_the_pickle(small_bb, tmp_path):
    import pickle
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    small_bb(16, 40)
    data = np.random.default_rng(1).uniform(0, 90000, size=(32, 320)).astype(np.float32)
    splines = _nonlin_splines(5)
    path = tmp_path / 'nonlin.pkl'
    path.write_bytes(pickle.dumps(splines))
    bbr.tel = 'ML1'
    got = bbr.nonlin_corr(data.copy(), str(path))
    assert np.array_equal(got, R.nonlin_corr(data.copy(), splines, tel='ML1'))


def test_drop_in_functions_refuse_mismatched_shapes(small_bb):
    import torch
    from blackbox_b200 import reduce as bbr
    small_bb(24, 40)
    img = np.zeros((48, 320), dtype=np.float32)
    bbr.tel = 'BG3'
    with pytest.raises(ValueError):
        bbr.xtalk_corr(img, np.zeros((16, 16)), data_mask=np.zeros((48, 300), np.uint8))
    with pytest.raises(ValueError):
        bbr.xtalk_corr(img, np.zeros((15, 16)))
    with pytest.raises(ValueError):
        bbr.master_combine([img, img[:-1]], 'bias')
    with pytest.raises(ValueError):
        bbr.master_combine([img, img], 'flat', medsec=[1.0])
    with pytest.raises(ValueError):
        bbr.detect_cosmics(img, inmask=np.zeros((48, 319), bool), satlevel=np.inf, sepmed=False, cleantype='medmask')
    with pytest.raises(ValueError):
        bbr.fill_edge_pixels(img, np.zeros((47, 320), np.uint8))
    with pytest.raises(ValueError):
        bbr.subtract_mbias(torch.zeros((48, 320), device='cuda'), torch.zeros((48, 319), device='cuda'))

