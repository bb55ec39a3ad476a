"""Golden fixtures (tests/golden/golden.json, made by tests/golden/make_golden.py): the oracle
must keep reproducing them (CPU), and the CUDA path must hit them too (GPU)."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'golden.json')))


def _maker():
    spec = importlib.util.spec_from_file_location('make_golden', os.path.join(HERE, 'golden', 'make_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('idx', range(len(GOLD['chain'])))
def test_oracle_reproduces_chain_golden(idx, small_bb):
    g = GOLD['chain'][idx]
    small_bb(g['ysize_chan'])
    got = _maker().chain_case(g['tel'], g['seed'], g['ysize_chan'], g['niter'])
    assert got['raw_sha256'] == g['raw_sha256']
    assert got['mask_sha256'] == g['mask_sha256']
    assert got['mask_counts'] == g['mask_counts']
    assert got['image_spots'] == pytest.approx(g['image_spots'], rel=1e-6)
    for key in ('BIASMEAN', 'RDNOISE', 'NCOSMICS'):
        assert got[key] == pytest.approx(g[key], rel=1e-9)
    assert got['NOBJ-SAT'] == g['NOBJ-SAT']


@pytest.mark.parametrize('idx', range(len(GOLD['lacosmic'])))
def test_oracle_reproduces_lacosmic_golden(idx):
    g = GOLD['lacosmic'][idx]
    got = _maker().lacosmic_case(g['seed'])
    assert got == g


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD['chain'])))
def test_gpu_chain_hits_golden(idx, small_bb):
    from blackbox_b200 import set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    mk = _maker()
    g = GOLD['chain'][idx]
    small_bb(g['ysize_chan'])
    raw, _ = synth.make_raw(g['tel'], g['seed'], nstars=250, ncosmics=80)
    assert mk.digest(raw) == g['raw_sha256']
    shape = (2 * g['ysize_chan'], 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(g['tel'], g['seed'] + 1, shape)
    coeffs = synth.make_xtalk(g['seed'] + 2)[3]
    res = FramePipeline(g['tel'], raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs,
                        niter=g['niter']).reduce(raw)
    mask = res.mask.cpu().numpy()
    img = res.img.cpu().numpy()
    assert mk.digest(mask) == g['mask_sha256']            # masks: bit-exact
    assert mk.spots(img) == pytest.approx(g['image_spots'], rel=1e-5)
    assert res.header['BIASMEAN'] == pytest.approx(g['BIASMEAN'], rel=1e-9)
    assert res.header['RDNOISE'] == pytest.approx(g['RDNOISE'], rel=1e-9)
    assert res.header['NOBJ-SAT'] == g['NOBJ-SAT']
    assert res.header['NCOSMICS'] == pytest.approx(g['NCOSMICS'])


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD['lacosmic'])))
def test_gpu_lacosmic_hits_golden(idx):
    from blackbox_b200 import reduce as bbr
    mk = _maker()
    g = GOLD['lacosmic'][idx]
    rng = np.random.default_rng(g['seed'])
    img = (300 + 17 * rng.standard_normal((96, 128))).astype(np.float32)
    for _ in range(30):
        y, x = rng.integers(0, 96), rng.integers(0, 120)
        img[y, x:x + rng.integers(1, 6)] += rng.uniform(800, 30000)
    assert mk.digest(img) == g['input_sha256']
    crmask, clean = bbr.detect_cosmics(img, sigclip=15, sigfrac=0.01, objlim=3, niter=4, readnoise=8.5,
                                       gain=1.0, satlevel=np.inf, cleantype='medmask', sepmed=False)
    assert mk.digest(crmask.astype(np.uint8)) == g['crmask_sha256']
    assert mk.digest(clean) == g['clean_sha256']


@pytest.mark.parametrize('idx', range(len(GOLD['extras'])))
def test_oracle_reproduces_extras_golden(idx):
    g = GOLD['extras'][idx]
    got = _maker().extras_case(g['seed'])
    assert got == g


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(GOLD['extras'])))
def test_gpu_extras_hit_golden(idx, small_bb):
    """Master combine (plain / sigma-clipped), nonlin_corr and the edge-pixel fill on the GPU
    reproduce the digests of the committed fixtures (same seeded inputs as make_golden.py)."""
    import torch
    from scipy import interpolate
    from blackbox_b200 import reduce as bbr
    mk = _maker()
    g = GOLD['extras'][idx]
    rng = np.random.default_rng(g['seed'])
    small_bb(24, 40)
    shape = (48, 320)
    frames = [(1000 + 10 * rng.standard_normal(shape)).astype(np.float32) for _ in range(20)]
    for k in (0, 7, 13):
        hit = rng.random(shape) < 0.03
        frames[k][hit] += rng.uniform(100, 5000, hit.sum()).astype(np.float32)
    dev = [torch.from_numpy(f).cuda() for f in frames]
    plain, _ = bbr.master_combine(dev, 'bias')
    clipped, _ = bbr.master_combine(dev, 'bias', clip_sigma=3.0, clip_maxiters=5)
    assert mk.digest(plain.cpu().numpy()) == g['median20_sha256']
    assert mk.digest(clipped.cpu().numpy()) == g['clipped20_sha256']
    splines = []
    for i in range(16):
        x = np.linspace(0, 60000, 80)
        y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
        splines.append(interpolate.UnivariateSpline(x, y, k=3, s=x.size * 4e-8))
    data = rng.uniform(-500, 140000, size=shape).astype(np.float32)
    bbr.tel = 'BG3'
    nl = bbr.nonlin_corr(data.copy(), splines)
    assert mk.digest(nl) == g['nonlin_sha256']
    mask = np.zeros(shape, dtype=np.uint8)
    mask[:2] = 32
    mask[:, -3:] = 33
    filled = data.copy()
    meds = bbr.fill_edge_pixels(filled, mask)
    assert mk.digest(filled) == g['edge_fill_sha256']
    assert [float(m) for m in meds.cpu().numpy()] == g['channel_medians']
