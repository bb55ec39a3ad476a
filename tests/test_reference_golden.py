"""The oracle (and the GPU path) against vectors produced by EXECUTING THE REFERENCE'S OWN CODE
(tests/golden/reference_golden.json, made by tests/golden/make_reference_golden.py from
/root/reference/blackbox.py with stub modules for its absent dependencies; astropy's sigma clipping
and astroscrappy are the oracle's restatements in that run, everything else -- define_sections,
gain_corr, os_corr, mask_init, cosmics_corr, xtalk_corr, nonlin_corr -- is the reference verbatim).
The fixtures travel; /root/reference is not needed to run these tests."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'reference_golden.json')))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def spots(a, n=8):
    flat = np.asarray(a).ravel()
    return [float(flat[i]) for i in np.linspace(0, flat.size - 1, n).astype(int)]


def _inputs(g):
    from blackbox_b200 import set_bb, synth
    tel, seed, ysc = g['tel'], g['seed'], g['ysize_chan']
    if g.get('xbin', 1) == 2:
        raw, _ = synth.make_raw(tel, seed, ysize_chan=ysc // 2, xsize_chan=660, os_rows=10, os_cols=90,
                                nstars=300, ncosmics=60)
        assert digest(raw) == g['raw_sha256']
        return raw, None, None, None, None
    raw, _ = synth.make_raw(tel, seed, nstars=400, ncosmics=150)
    if tel != 'ML1':
        raw[ysc - 50:ysc, 300:304] = 65535
        raw[ysc - 900:ysc - 880, 1500 * 2 + 20:1500 * 2 + 24] = 65535
    raw[40:48, 2000:2008] = 65535
    if g.get('variant') == 'hos':
        synth.add_hos_contamination(raw, ysc)
    if g.get('variant') == 'rings':
        synth.add_saturated_rings(raw)
    assert digest(raw) == g['raw_sha256']
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
    coeffs = synth.make_xtalk(seed + 2)[3]
    return raw, mbias, mflat, bpm, coeffs


def test_define_sections_equals_the_reference():
    from blackbox_b200.geometry import define_sections
    for key, want in GOLD['sections'].items():
        shape, b = key.split('_bin')
        H, W = (int(v) for v in shape.split('x'))
        got = define_sections((H, W), xbin=int(b), ybin=int(b), tel='BG3')
        got = [[[[s.start, s.stop] for s in pair] for pair in tup] for tup in got]
        assert got == want, key


@pytest.mark.parametrize('idx', range(len(GOLD['frames'])))
def test_oracle_equals_the_reference_step_by_step(idx, small_bb):
    """gain_corr, os_corr, mask_init, mask_header bit for bit (and, on the small MeerLICHT frame,
    cosmics_corr and xtalk_corr too: the big BlackGEM frame's LACosmic takes minutes on the CPU and
    is left to the GPU test below).  Case 2 is a FULL-SIZE BG2 frame (the only size at which the
    reference's saturated-column windows of BG2 exist: channel-9 split fit, blackbox.py:6716-6760),
    without cosmics_corr; case 3 a 2x2-binned frame through gain_corr + os_corr."""
    from blackbox_b200 import set_bb, synth
    from oracle import reduce as R
    g = GOLD['frames'][idx]
    tel = g['tel']
    small_bb(g['ysize_chan'], lim=dict(set_bb.hos_sat_ypix_lim))
    raw, mbias, mflat, bpm, coeffs = _inputs(g)
    header = {'EXPTIME': 60.0}
    data = np.array(raw, dtype=np.float32)
    R.gain_corr(data, header, tel=tel)
    assert digest(data) == g['gain_sha256']
    xbin = g.get('xbin', 1)
    data = R.os_corr(data, header, 'object', xbin=xbin, ybin=xbin, tel=tel)
    assert spots(data) == pytest.approx(g['os_spots'], rel=1e-6)
    assert digest(data) == g['os_sha256']
    for key, want in g['os_header'].items():
        assert header[key] == want, key
    if xbin == 2:
        return
    if set_bb.get_par(set_bb.subtract_mbias, tel):
        data -= mbias
    if g.get('variant') == 'rings':
        synth.add_nonfinite(data, bpm)
    data_mask, header_mask = R.mask_init(data, header, bpm, 'object', tel=tel)
    assert {str(b): int(((data_mask & b) != 0).sum()) for b in (1, 4, 8, 32, 64)} == g['mask_counts']
    assert digest(data_mask) == g['mask_init_sha256']
    for key, want in g['mask_header'].items():
        if key != 'NCOSMICS':
            assert float(header_mask[key]) == want, key
    if g['ysize_chan'] > 400 and g['cosmics']:
        return
    data /= mflat
    if g['cosmics']:
        data, data_mask = R.cosmics_corr(data, header, data_mask, header_mask, tel=tel)
        assert digest(data_mask) == g['cosmics_mask_sha256'] and digest(data) == g['cosmics_sha256']
        assert header['NCOSMICS'] == g['NCOSMICS']
    hm2 = {}
    R.mask_header(data_mask, hm2, tel=tel)
    assert {k: int(v) for k, v in hm2.items() if k.endswith('NUM')} == g['mask_header_counts']
    R.xtalk_corr(data, coeffs, data_mask, tel=tel)
    assert digest(data) == g['xtalk_sha256']


def test_oracle_nonlin_equals_the_reference(small_bb):
    from scipy import interpolate
    from oracle import reduce as R
    g = GOLD['nonlin'][0]
    small_bb(48, 60)
    rng = np.random.default_rng(g['seed'])
    splines = []
    for i in range(16):
        x = np.linspace(0, 60000, 80)
        y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
        splines.append(interpolate.UnivariateSpline(x, y, k=3, s=x.size * 4e-8))
    data = rng.uniform(-500, 140000, size=(96, 480)).astype(np.float32)
    assert digest(data) == g['input_sha256']
    assert digest(R.nonlin_corr(data.copy(), splines, tel='BG3')) == g['output_sha256']


def _oracle_chain(g, raw, mbias, mflat, bpm, coeffs):
    """The oracle through the case's steps on the case's inputs, holding on to the frame after every
    step.  Each stage the fixture has a digest for is checked against it first, so the arrays the
    GPU is compared with below ARE the reference's own outputs (the fixture was made by executing
    blackbox.py), just no longer squeezed through a hash."""
    from blackbox_b200 import set_bb, synth
    from oracle import reduce as R
    tel, xbin = g['tel'], g.get('xbin', 1)
    out = {}
    header = {'EXPTIME': 60.0}
    data = np.array(raw, dtype=np.float32)
    R.gain_corr(data, header, tel=tel)
    data = R.os_corr(data, header, 'object', xbin=xbin, ybin=xbin, tel=tel)
    assert digest(data) == g['os_sha256']
    out['os'] = data.copy()
    out['header'] = header
    if xbin == 2:
        return out
    if set_bb.get_par(set_bb.subtract_mbias, tel):
        data -= mbias
    if g.get('variant') == 'rings':
        synth.add_nonfinite(data, bpm)
    out['pre_mask'] = data.copy()
    data_mask, header_mask = R.mask_init(data, header, bpm, 'object', tel=tel)
    assert digest(data_mask) == g['mask_init_sha256']
    out['mask_init'] = data_mask.copy()
    data /= mflat
    if g['cosmics']:
        data, data_mask = R.cosmics_corr(data, header, data_mask, header_mask, tel=tel)
        assert digest(data_mask) == g['cosmics_mask_sha256'] and digest(data) == g['cosmics_sha256']
        out['cosmics'] = data.copy()
    R.xtalk_corr(data, coeffs, data_mask, tel=tel)
    assert digest(data) == g['xtalk_sha256']
    out['final'], out['mask'], out['header_mask'] = data, data_mask, header_mask
    return out


def _same_frame(got, want, scale, what):
    """The float class of the parity contract over the WHOLE frame: every pixel within 1e-5
    relative (+ 1e-5 of the level that was subtracted), and at least 99.9 % of them identical."""
    from conftest import float_class_ok
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, what
    assert float_class_ok(got, want, scale=scale).all(), what
    assert np.mean(got == want) >= 0.999, (what, float(np.mean(got == want)))


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD['frames'])))
def test_gpu_chain_against_the_reference(idx, small_bb):
    """Every reference-made case on the GPU: masks bit for bit, header values, and the FULL image
    against the oracle's frame for the same inputs (which is first shown to be the reference's own
    output, digest by digest).  Whole-chain cases go through FramePipeline; the cases that stop
    short of the whole chain, or that inject non-finite pixels half way (the saturated-rings case:
    fill_sat_holes and the non-finite rule of mask_init), through the step functions."""
    from blackbox_b200 import set_bb
    from blackbox_b200.pipeline import FramePipeline
    g = GOLD['frames'][idx]
    tel = g['tel']
    small_bb(g['ysize_chan'], lim=dict(set_bb.hos_sat_ypix_lim))
    raw, mbias, mflat, bpm, coeffs = _inputs(g)
    want = _oracle_chain(g, raw, mbias, mflat, bpm, coeffs)
    if g.get('xbin', 1) == 2 or not g['cosmics'] or g.get('variant') == 'rings':
        return _gpu_steps(g, want, raw, mbias, mflat, bpm, coeffs)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, exptime=60.0)
    res = pipe.reduce(raw)
    assert digest(res.mask.cpu().numpy()) == g['cosmics_mask_sha256']
    assert {k: int(v) for k, v in res.header_mask.items() if k.endswith('NUM')} == g['mask_header_counts']
    assert res.header['NCOSMICS'] == g['NCOSMICS']
    for key in ('BIASMEAN', 'RDNOISE'):
        assert res.header[key] == pytest.approx(g['os_header'][key], rel=1e-9)
    for key, val in g['os_header'].items():
        if isinstance(val, bool):
            assert res.header[key] == val, key
        elif isinstance(val, float):
            # monomial coefficients come out of a different (orthogonal-polynomial) solver: 1e-5;
            # levels and noise values: 1e-8
            coef = key.startswith('BIAS') and 'A' in key[4:]
            assert res.header[key] == pytest.approx(val, rel=1e-5 if coef else 1e-8, abs=1e-9 if coef else 1e-12), key
    assert res.header['SATURATE'] == pytest.approx(g['mask_header']['SATURATE'], rel=1e-9)
    assert res.header['NOBJ-SAT'] == g['mask_header']['NOBJ-SAT']
    _same_frame(res.img.cpu().numpy(), want['final'], g['os_header']['BIASMEAN'], 'final image')


def _gpu_steps(g, want, raw, mbias, mflat, bpm, coeffs):
    """The cases that do not run as one pipeline, through the drop-in step functions in
    blackbox_reduce's order, every intermediate frame against the oracle's."""
    import torch
    from blackbox_b200 import reduce as bbr, set_bb, synth
    tel, xbin = g['tel'], g.get('xbin', 1)
    bbr.tel = tel
    scale = g['os_header']['BIASMEAN']
    header = {'EXPTIME': 60.0}
    data = bbr.os_corr(raw, header, 'object', xbin=xbin, ybin=xbin, tel=tel)
    for key in ('BIASMEAN', 'RDNOISE'):
        assert header[key] == pytest.approx(g['os_header'][key], rel=1e-9)
    _same_frame(data, want['os'], scale, 'os_corr')
    if xbin == 2:
        return
    if set_bb.get_par(set_bb.subtract_mbias, tel):
        data = data - mbias
    if g.get('variant') == 'rings':
        synth.add_nonfinite(data, bpm)
    dev = torch.from_numpy(data).cuda()
    mask, header_mask = bbr.mask_init(dev, header, 'q', 'object', bpm=bpm)
    assert digest(mask.cpu().numpy()) == g['mask_init_sha256']
    assert torch.isfinite(dev).all()                              # non-finite pixels zeroed in place
    assert float(header_mask['SATURATE']) == pytest.approx(g['mask_header']['SATURATE'], rel=1e-9)
    assert header_mask['NOBJ-SAT'] == g['mask_header']['NOBJ-SAT']
    if not g['cosmics']:
        hm2 = {}
        bbr.mask_header(mask, hm2)
        assert {k: int(v) for k, v in hm2.items() if k.endswith('NUM')} == g['mask_header_counts']
    dev /= torch.from_numpy(mflat).cuda()
    if g['cosmics']:
        dev, mask = bbr.cosmics_corr(dev, header, mask, header_mask)
        assert digest(mask.cpu().numpy()) == g['cosmics_mask_sha256']
        assert header['NCOSMICS'] == g['NCOSMICS']
        _same_frame(dev.cpu().numpy(), want['cosmics'], scale, 'cosmics_corr')
        hm2 = {}
        bbr.mask_header(mask, hm2)
        assert {k: int(v) for k, v in hm2.items() if k.endswith('NUM')} == g['mask_header_counts']
    bbr.xtalk_corr(dev, coeffs, mask)
    _same_frame(dev.cpu().numpy(), want['final'], scale, 'final image')


@pytest.mark.gpu
def test_gpu_nonlin_against_the_reference(small_bb):
    from scipy import interpolate
    from blackbox_b200 import reduce as bbr
    g = GOLD['nonlin'][0]
    small_bb(48, 60)
    rng = np.random.default_rng(g['seed'])
    splines = []
    for i in range(16):
        x = np.linspace(0, 60000, 80)
        y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
        splines.append(interpolate.UnivariateSpline(x, y, k=3, s=x.size * 4e-8))
    data = rng.uniform(-500, 140000, size=(96, 480)).astype(np.float32)
    bbr.tel = 'BG3'
    assert digest(bbr.nonlin_corr(data.copy(), splines)) == g['output_sha256']


@pytest.mark.skipif(not os.path.isfile('/root/reference/blackbox.py'), reason='the reference tree is not on this box')
def test_generator_still_imports_and_runs_the_reference():
    """Where the reference is present (the build container), the generator's stub modules still let
    blackbox.py import, and its define_sections / gain_corr give what the fixture file holds --
    i.e. the committed vectors can be regenerated.  In a subprocess: the stubs go into sys.modules."""
    import subprocess
    import sys
    code = (
        "import sys, json, numpy as np\n"
        "sys.path.insert(0, {gold!r}); sys.path.insert(0, {root!r})\n"
        "import make_reference_golden as g\n"
        "bb = g.load_reference()\n"
        "secs = g.sections_as_lists(bb.define_sections((10600, 12000), xbin=1, ybin=1, tel='BG3'))\n"
        "data = np.full((440, 12000), 2.0, dtype='float32'); hdr = g.Header()\n"
        "bb.gain_corr(data, hdr, tel='ML1')\n"
        "print(json.dumps({{'version': bb.__version__, 'secs': secs, 'gain1': hdr['GAIN1'], 'px': float(data[0, 0])}}))\n"
    ).format(gold=os.path.join(HERE, 'golden'), root=os.path.dirname(HERE))
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out['version'] == GOLD['reference_version']
    assert out['secs'] == GOLD['sections']['10600x12000_bin1']
    assert out['gain1'] == 2.112 and out['px'] == float(np.float32(2.0) * np.float32(2.112))
