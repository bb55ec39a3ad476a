"""The CPU oracle checked against independent statements of the same algorithms: plain Python
loops for the sigma clipping, scipy.ndimage for the filters, a per-pair loop for the crosstalk,
and the numpy twin of the C LACosmic.  (The reference ships no tests or golden vectors and
astropy / astroscrappy are not installable here: see DESIGN.md, "Oracle".)"""
import numpy as np
import pytest
from scipy import ndimage

from oracle import clib, lacosmic, reduce as R, stats


# ------------------------------------------------------------------------------------------
# sigma clipping
# ------------------------------------------------------------------------------------------
def naive_clip(values, sigma, maxiters, cen):
    """astropy's published loop, one slice, pure Python; returns the final bounds."""
    buf = [float(v) for v in values]
    lo = hi = float('nan')
    it = 0
    while buf:
        mean = sum(buf) / len(buf)
        std = (sum((mean - v) ** 2 for v in buf) / len(buf)) ** 0.5
        c = mean if cen == 'mean' else float(np.median(buf))
        lo, hi = c - sigma * std, c + sigma * std
        kept = [v for v in buf if lo <= v <= hi]
        if len(kept) == len(buf):
            break
        buf = kept
        it += 1
        if it >= maxiters:
            break
    return lo, hi


@pytest.mark.parametrize('cen', ['mean', 'median'])
def test_sigma_clip_matches_naive_loop(cen):
    rng = np.random.default_rng(1)
    data = rng.normal(100, 5, (40, 37)).astype(np.float32)
    data[rng.random(data.shape) < 0.05] += 300
    data[3, :] = 0
    data[5, 2] = np.nan
    for axis in (0, 1, None):
        out = stats.sigma_clip(data, sigma=2.5, maxiters=5, cenfunc=cen, axis=axis, masked=True)
        m = np.ma.getmaskarray(out)
        slices = [data.ravel()] if axis is None else ([data[:, j] for j in range(37)] if axis == 0 else list(data))
        masks = [m.ravel()] if axis is None else ([m[:, j] for j in range(37)] if axis == 0 else list(m))
        for v, mk in zip(slices, masks):
            ok = np.isfinite(v)
            lo, hi = naive_clip(v[ok].astype(np.float64), 2.5, 5, cen)
            want = ~ok | (v < lo) | (v > hi)
            assert np.array_equal(mk, want)


def test_sigma_clipped_stats_values_and_mask_value():
    rng = np.random.default_rng(2)
    data = rng.normal(6500, 9, (50, 174)).astype(np.float32)
    data[:, 3] = 0.0                        # masked by mask_value=0
    data[7, 10:14] = 9000.0
    mean, med, std = stats.sigma_clipped_stats(data, axis=1, mask_value=0, cenfunc='mean')
    assert mean.dtype == np.float64 and mean.shape == (50,)
    for r in range(50):
        v = data[r].astype(np.float64)
        ok = v != 0
        lo, hi = naive_clip(v[ok], 3.0, 5, 'mean')
        keep = ok & (v >= lo) & (v <= hi)
        assert mean[r] == pytest.approx(v[keep].mean(), rel=1e-14)
        assert std[r] == pytest.approx(v[keep].std(), rel=1e-12)
        assert med[r] == np.median(v[keep])
    m1, _, s1 = stats.sigma_clipped_stats(data, mask_value=0, cenfunc='mean')
    v = data.ravel().astype(np.float64)
    lo, hi = naive_clip(v[v != 0], 3.0, 5, 'mean')
    keep = (v != 0) & (v >= lo) & (v <= hi)
    assert m1 == pytest.approx(v[keep].mean(), rel=1e-13) and s1 == pytest.approx(v[keep].std(), rel=1e-12)


def test_sigma_clip_edge_cases():
    z = np.zeros((4, 6), np.float32)
    mean, med, std = stats.sigma_clipped_stats(z, axis=1, mask_value=0)
    assert np.isnan(mean).all()
    c = np.full(9, 3.25, np.float32)
    mean, med, std = stats.sigma_clipped_stats(c)
    assert (mean, med, std) == (3.25, 3.25, 0.0)
    out = stats.sigma_clip(np.array([1.0, 1.0, 1.0, 100.0]), sigma=1.0, maxiters=None or 5, cenfunc='mean')
    assert np.ma.getmaskarray(out).tolist() == [False, False, False, True]


# ------------------------------------------------------------------------------------------
# LACosmic building blocks
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('k', [3, 5, 7])
def test_medfilt_interior_is_scipy_border_is_input(k):
    rng = np.random.default_rng(k)
    img = rng.normal(0, 1, (31, 45)).astype(np.float32)
    out = {3: clib.medfilt3, 5: clib.medfilt5, 7: clib.medfilt7}[k](img)
    r = k // 2
    ref = ndimage.median_filter(img, size=k)
    assert np.array_equal(out[r:-r, r:-r], ref[r:-r, r:-r])
    border = np.ones(img.shape, bool)
    border[r:-r, r:-r] = False
    assert np.array_equal(out[border], img[border])


def test_dilations_match_scipy():
    rng = np.random.default_rng(5)
    b = (rng.random((40, 33)) < 0.05)
    d3 = clib.dilate3(b.astype(np.uint8)).astype(bool)
    ref = ndimage.binary_dilation(b, structure=np.ones((3, 3), bool))
    assert np.array_equal(d3[1:-1, 1:-1], ref[1:-1, 1:-1])
    edge = np.ones(b.shape, bool)
    edge[1:-1, 1:-1] = False
    assert np.array_equal(d3[edge], b[edge])
    st = np.ones((5, 5), bool)
    st[0, 0] = st[0, 4] = st[4, 0] = st[4, 4] = False
    d5 = clib.dilate5(b.astype(np.uint8), 2).astype(bool)
    assert np.array_equal(d5, ndimage.binary_dilation(b, structure=st, iterations=2))


def test_laplace_plus_c_equals_numpy_twin():
    rng = np.random.default_rng(6)
    for shape in [(1, 1), (2, 3), (17, 29)]:
        img = rng.normal(50, 30, shape).astype(np.float32)
        c = clib.rebin(np.maximum(clib.laplace(clib.subsample(img)), 0))
        assert np.array_equal(c, lacosmic.laplace_plus(img))
    # Laplacian of a constant image: only the zero-padded frame responds
    flat = np.full((6, 7), 10.0, np.float32)
    lp = lacosmic.laplace_plus(flat)
    assert (lp[1:-1, 1:-1] == 0).all() and (lp[0, :] > 0).all()


def test_lower_median_is_a_k_minus_1_over_2():
    rng = np.random.default_rng(7)
    for n in (1, 2, 3, 10, 11, 1000):
        a = rng.normal(0, 1, n).astype(np.float32)
        assert clib.lower_median(a) == np.sort(a)[(n - 1) // 2]


@pytest.mark.parametrize('seed', [1, 2])
def test_detect_cosmics_c_equals_numpy_twin(seed):
    rng = np.random.default_rng(seed)
    shape = (70, 90)
    img = (300 + 17 * rng.standard_normal(shape)).astype(np.float32)
    for _ in range(25):
        y, x = rng.integers(0, shape[0]), rng.integers(0, shape[1])
        img[y, x:x + rng.integers(1, 5)] += rng.uniform(800, 30000)
    img[30:37, 40:47] += 40000.0                       # fat blob -> background level is used
    mask = rng.random(shape) < 0.02
    kw = dict(sigclip=15, sigfrac=0.01, objlim=3, readnoise=8.5, niter=4)
    info_c, info_n = {}, {}
    cr_c, clean_c = lacosmic.detect_cosmics(img, inmask=mask, gain=1.0, satlevel=np.inf, sepmed=False,
                                            cleantype='medmask', info=info_c, **kw)
    cr_n, clean_n = lacosmic.detect_cosmics_numpy(img, inmask=mask, info=info_n, **kw)
    assert cr_c.sum() > 30
    assert np.array_equal(cr_c, cr_n)
    assert np.array_equal(clean_c.view(np.uint32), clean_n.view(np.uint32))
    assert info_c['background'] == info_n['background']
    assert (clean_c == info_c['background']).any()
    assert not np.array_equal(clean_c, img) and np.array_equal(clean_c[~cr_c], img[~cr_c])


def test_detect_cosmics_refuses_other_modes():
    with pytest.raises(NotImplementedError):
        lacosmic.detect_cosmics(np.zeros((8, 8), np.float32))


# ------------------------------------------------------------------------------------------
# reduction steps
# ------------------------------------------------------------------------------------------
def test_xtalk_equals_per_pair_loop(small_bb):
    """The matmul formulation against the reference's older per-pair statement of the same
    correction (victim -= coeff * source, y-mirrored across CCD halves; blackbox.py:7263-7388),
    with every source taken from the uncorrected image."""
    from blackbox_b200 import synth
    from blackbox_b200.geometry import define_sections
    small_bb(24, 32)
    rng = np.random.default_rng(3)
    shape = (48, 256)
    data = rng.normal(40, 30, shape).astype(np.float32)
    mask = rng.choice(np.array([0, 0, 1, 2, 32], np.uint8), size=shape)
    _, _, _, coeffs = synth.make_xtalk(5)
    got = data.copy()
    R.xtalk_corr(got, coeffs, mask, tel='ML1')
    chan = define_sections(shape, tel='ML1')[0]
    src_ok = (data > 0) & (mask & 1 == 0) & (mask & 2 == 0)
    want = data.astype(np.float64).copy()
    for v in range(16):
        corr = np.zeros((24, 32))
        for s in range(16):
            src = (data[chan[s]] * src_ok[chan[s]]).astype(np.float64)
            if s // 8 != v // 8:
                src = np.flipud(src)
            corr += coeffs[s, v] * src
        want[chan[v]] -= corr * (mask[chan[v]] & 32 == 0)
    np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-6)
    assert np.mean(got == want.astype(np.float32)) > 0.999


def test_gain_corr_is_a_float32_multiply(small_bb):
    small_bb(24, 32)
    from blackbox_b200 import set_bb
    rng = np.random.default_rng(4)
    raw = rng.integers(0, 65535, (2 * 44, 8 * 212)).astype(np.float32)
    data = raw.copy()
    hdr = {}
    R.gain_corr(data, hdr, tel='BG3')
    g = np.float32(set_bb.gain['BG3'][9])
    assert np.array_equal(data[44:, 212:424], raw[44:, 212:424] * g)
    assert hdr['GAIN10'] == set_bb.gain['BG3'][9]


def test_mask_init_semantics(small_bb):
    small_bb(40, 48)
    shape = (80, 384)
    data = np.full(shape, 100.0, np.float32)
    sat = 1e6
    data[10:13, 10:13] = sat                      # blob in channel 0
    data[11, 11] = 100.0                          # hole -> filled only if mask == 0 there
    data[70, 200] = sat                           # single pixel, top half (channel 12)
    data[5, 5] = np.nan
    bpm = np.zeros(shape, np.uint8)
    bpm[20, 20] = 1
    hdr = {'BIASM{}'.format(i + 1): 0.0 for i in range(16)}
    mask, hm = R.mask_init(data, hdr, bpm, 'object', tel='ML1')
    mv = {'bad': 1, 'sat': 4, 'satcon': 8, 'xtalk': 64}
    assert data[5, 5] == 0 and mask[5, 5] == mv['bad'] and mask[20, 20] == 1
    assert (mask[10:13, 10:13] & mv['sat']).sum() == 8 * mv['sat']
    assert mask[9, 9] & mv['satcon'] and not mask[9, 9] & mv['sat']
    # crosstalk victims: same position in the other bottom channels, mirrored in the top ones
    assert mask[10, 48 + 10] & mv['xtalk'] and mask[40 + (39 - 10), 10] & mv['xtalk']
    assert not mask[10, 10] & mv['xtalk']         # a channel is not its own victim
    # the hole pixel (11, 11) is a crosstalk victim of nobody here -> mask == 0 -> filled
    assert mask[11, 11] == mv['satcon']
    assert hdr['NOBJ-SAT'] == 2 and hm['NOBJ-SAT'] == 2
    assert hm['SATURATE'] == pytest.approx(np.mean(hdr['SATURATE']))


def test_master_median_flat_rules(small_bb):
    small_bb(40, 48)
    rng = np.random.default_rng(8)
    frames = [rng.normal(1000 * (i + 1), 5, (80, 384)).astype(np.float32) for i in range(4)]
    bpm = np.zeros((80, 384), np.uint8)
    bpm[0, :] = 32
    out, scales = R.master_median(frames, 'flat', medsec=[1000.0, 2000.0, 3000.0, 4000.0], bpm=bpm)
    assert (out[0] == 1).all() and abs(np.median(out[1:]) - 1) < 0.01
    cube = np.stack([f / np.float32(s) for f, s in zip(frames, scales)])
    assert np.array_equal(out[1:], np.median(cube, axis=0)[1:])


def test_running_median_and_spline_match_product_host_code():
    from blackbox_b200 import hostfit
    rng = np.random.default_rng(9)
    for n in (0, 2, 3, 4, 5, 6, 9, 180):
        y = rng.normal(0, 1, n).astype(np.float32)
        assert np.array_equal(hostfit.running_median3(y), R.running_median3(y) if n else y)
    mean = (5 * np.exp(-np.arange(1320) / 40.0) + rng.normal(0, 0.3, 1320)).astype(np.float32)
    std = np.abs(rng.normal(1, 0.1, 1320)).astype(np.float32)
    n = np.full(1320, 10)
    n[7] = 1
    d = {}
    R.hos_fit(mean, std, n, np.zeros(1320, bool), 'BG3', 0, diag=d)
    np.testing.assert_array_equal(hostfit.hos_spline(mean, std, n), d['spline'])


def test_os_corr_recovers_injected_levels(small_bb):
    from blackbox_b200 import synth
    small_bb(120)
    raw, truth = synth.make_raw('ML1', 11, nstars=50, ncosmics=5)
    data = raw.astype(np.float32)
    hdr = {}
    R.gain_corr(data, hdr, tel='ML1')
    diag = {}
    out = R.os_corr(data, hdr, 'object', tel='ML1', diag=diag)
    assert out.shape == (240, 10560) and out.dtype == np.float32
    from blackbox_b200 import set_bb
    for i in range(16):
        assert hdr['BIASM{}'.format(i + 1)] == pytest.approx(truth['bias_adu'][i] * set_bb.gain['ML1'][i], abs=25)
        assert 6 < hdr['RDN{}'.format(i + 1)] < 11
        assert hdr['VFITOK{}'.format(i + 1)] is True
    # the sky level (150 ADU) survives, the bias level (3050 ADU) does not
    assert 250 < np.median(out) < 400


def test_sigma_clip_mean_matches_scipy_sigmaclip():
    """An independent published implementation of the same loop: scipy.stats.sigmaclip (mean and
    population std of the survivors, keep lo <= x <= hi, until nothing is rejected) is
    astropy's sigma_clip with cenfunc='mean' and no iteration limit.  The restatement must keep
    the same survivors and end on the same bounds."""
    from scipy import stats as sstats
    from oracle.stats import sigma_clip
    rng = np.random.default_rng(8)
    for n, nout in [(50, 3), (1000, 40), (5000, 0), (7, 1)]:
        x = rng.normal(100.0, 5.0, n)
        x[:nout] += rng.uniform(30, 500, nout)
        for sig in (2.0, 3.0):
            kept, lo, hi = sstats.sigmaclip(x, sig, sig)
            ours = sigma_clip(x, sigma=sig, maxiters=None, cenfunc='mean', masked=True)
            assert np.array_equal(np.sort(ours.compressed()), np.sort(kept)), (n, nout, sig)
            surv = ours.compressed()
            assert lo == pytest.approx(surv.mean() - sig * surv.std(), rel=1e-12)
            assert hi == pytest.approx(surv.mean() + sig * surv.std(), rel=1e-12)


def test_recalled_choices_are_switchable_and_default_to_what_is_implemented():
    """Every recalled detail of the restated third-party algorithms sits behind a named switch of the
    C oracle (value 0 = implemented by the oracle and by the CUDA kernels).  Flipping one changes the
    result where it should, resetting restores it; the committed matrix
    (tests/golden/oracle_choice_matrix.json, tools/oracle_choice_matrix.py) quantifies each on the
    seeded frames."""
    import json
    import os
    from oracle import clib, lacosmic as L
    assert clib.lib().bbo_num_choices() == len(clib.CHOICES)
    clib.reset_choices()
    rng = np.random.default_rng(5)
    img = rng.normal(100, 3, (40, 40)).astype(np.float32)
    img[20, 20] += 4000.0                                  # a one-pixel cosmic ray
    mask = np.zeros(img.shape, bool)
    mask[19, 19] = True                                    # 25 - 1 (itself) - 1 (masked) = 23 usable... make it even:
    mask[21, 21] = True                                    # 22 usable neighbours: lower != upper median
    kw = dict(sigclip=15, sigfrac=0.01, objlim=3, gain=1.0, readnoise=5.0, satlevel=np.inf, niter=1,
              sepmed=False, cleantype='medmask')
    cr0, clean0 = L.detect_cosmics(img, mask, **kw)
    assert cr0[20, 20]
    try:
        clib.set_choice('CLEAN_MEDIAN', 1)
        cr1, clean1 = L.detect_cosmics(img, mask, **kw)
        assert np.array_equal(cr1, cr0) and clean1[20, 20] > clean0[20, 20]
        clib.set_choice('CLEAN_MEDIAN', 0)
        clib.set_choice('SIGCLIP_CMP', 1)
        cr2, clean2 = L.detect_cosmics(img, mask, **kw)
        assert np.array_equal(cr2, cr0) and np.array_equal(clean2, clean0)
    finally:
        clib.reset_choices()
    cr3, clean3 = L.detect_cosmics(img, mask, **kw)
    assert np.array_equal(cr3, cr0) and np.array_equal(clean3, clean0)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oracle_choice_matrix.json')
    rows = json.load(open(path))
    assert {r['switch'] for r in rows} == {n for n, _ in clib.CHOICES}
    moved = {r['switch'] for r in rows if any(f['cr_flags_changed'] or f['image_pixels_changed'] for f in r['frames'])}
    # on the seeded frames only these recalled details change anything at all
    assert moved == {'CLEAN_MEDIAN', 'LAPLACE_EDGE', 'MEDFILT_FRAME'}
    worst = max(f['cr_flags_changed'] / f['npix'] for r in rows for f in r['frames'] if 'no BPM' not in f['frame'])
    assert worst == 0.0                                   # with the bad-pixel mask's edge frame: no mask pixel moves
