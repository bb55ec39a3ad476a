"""GPU parity on 2x2-binned frames (BASELINE.json config 5: raw 5300x6000, data 2640x660 per
channel, 5 horizontal-overscan rows and 87 vertical-overscan columns in use;
blackbox.py:6345-6399 with xbin = ybin = 2) and the master bias of 50 frames."""
import numpy as np
import pytest

from conftest import float_class_ok

pytestmark = pytest.mark.gpu


def _binned_setup(small_bb, ysc_unbinned):
    """Unbinned settings (os_corr divides them by xbin / ybin itself, blackbox.py:6345-6354);
    the BlackGEM saturated-column windows are not scaled by the reference (SURVEY.md 8d cfg 5),
    so they are chosen to fall inside the binned channel."""
    q = ysc_unbinned // 8
    return small_bb(ysc_unbinned, 1320, lim={'BG2': (2 * q, 4 * q), 'BG3': (q, 2 * q), 'BG4': (q, 2 * q)})


def _binned_raw(tel, seed, ysc_unbinned):
    from blackbox_b200 import synth
    raw, _ = synth.make_raw(tel, seed, ysize_chan=ysc_unbinned // 2, xsize_chan=660, os_rows=10,
                            os_cols=90, nstars=300, ncosmics=60)
    return raw


@pytest.mark.parametrize('tel', ['ML1', 'BG3'])
def test_binned_os_corr_parity(tel, small_bb):
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    ysc = 400
    _binned_setup(small_bb, ysc)
    raw = _binned_raw(tel, 5001, ysc)
    assert raw.shape == (2 * (ysc // 2 + 10), 8 * (660 + 90))
    if tel != 'ML1':
        raw[150:200, 300:303] = 65535                 # saturated columns next to the overscan
    hdr_o, hdr_g = {}, {}
    data_o = raw.astype(np.float32)
    R.gain_corr(data_o, hdr_o, tel=tel)
    diag = {}
    out_o = R.os_corr(data_o.copy(), hdr_o, 'object', xbin=2, ybin=2, tel=tel, diag=diag)
    out_g, st = bbr.os_corr(raw, hdr_g, 'object', xbin=2, ybin=2, tel=tel, return_state=True)
    assert out_g.shape == out_o.shape == (ysc, 8 * 660)
    for i in range(16):
        ch = diag['chans'][i]
        np.testing.assert_allclose(st.mean_vos[i].cpu().numpy(), ch['mean_vos_col'], rtol=1e-13, atol=0)
        assert np.array_equal(st.hos_n[i].cpu().numpy(), ch['nvalues'])
        np.testing.assert_allclose(st.oscan[i].cpu().numpy(), ch['oscan'], rtol=0, atol=2e-5)
    for key in ('BIASMEAN', 'RDNOISE'):
        assert hdr_g[key] == pytest.approx(hdr_o[key], rel=1e-9)
    assert float_class_ok(out_g, out_o, scale=hdr_o['BIASMEAN']).all()
    assert np.mean(out_g == out_o) > 0.999


def test_binned_chain_with_masks(small_bb):
    """Config 5: binned BG3 frame through overscan + master bias + mask_init (bad-pixel mask,
    saturation, crosstalk victims, holes) + flat + LACosmic + crosstalk; the steps after os_corr
    see a reduced frame whose channel size is the binned one (set_bb.ysize_chan / xsize_chan =
    binned sizes, as the reference needs for xtalk_corr, blackbox.py:7214-7216)."""
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    tel, ysc = 'BG3', 400
    _binned_setup(small_bb, ysc)
    raw = _binned_raw(tel, 5002, ysc)
    raw[60:66, 1000:1006] = 65535                     # saturated blob
    shape = (ysc, 8 * 660)
    mbias, mflat, bpm = synth.make_masters(tel, 77, shape)
    coeffs = synth.make_xtalk(5)[3]
    # oracle, step by step
    hdr_o = {'EXPTIME': 60.0}
    d = raw.astype(np.float32)
    R.gain_corr(d, hdr_o, tel=tel)
    d = R.os_corr(d, hdr_o, 'object', xbin=2, ybin=2, tel=tel)
    set_bb.ysize_chan, set_bb.xsize_chan = ysc // 2, 660
    d -= mbias
    mask_o, hm_o = R.mask_init(d, hdr_o, bpm, 'object', tel=tel)
    d /= mflat
    d, mask_o = R.cosmics_corr(d, hdr_o, mask_o, hm_o, tel=tel, niter=3)
    R.xtalk_corr(d, coeffs, mask_o, tel=tel)
    # product: one pipeline call
    set_bb.ysize_chan, set_bb.xsize_chan = ysc, 1320
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=3,
                         xbin=2, ybin=2, exptime=60.0)
    res = pipe.reduce(raw)
    img, mask = res.img.cpu().numpy(), res.mask.cpu().numpy()
    assert img.shape == shape
    assert np.mean(mask != mask_o) <= 1e-5
    mv = set_bb.mask_value
    for name in ('saturated', 'saturated-connected', 'crosstalk', 'bad', 'edge'):
        assert np.array_equal(mask & mv[name], mask_o & mv[name]), name
    same_cr = (mask & mv['cosmic ray']) == (mask_o & mv['cosmic ray'])
    assert float_class_ok(img, d, scale=hdr_o['BIASMEAN'])[same_cr].all()
    assert np.mean(img == d) > 0.999
    assert res.header['NOBJ-SAT'] == hdr_o['NOBJ-SAT']


def test_master_bias_of_50_binned_frames():
    """Config 5's master bias: 50 frames of 5280x5280 (5.6 GB resident).  np.median on a
    1/64 row sample (bit-exact), and on the full frame the defining property of the median of
    an even count: exactly 25 frames <= lo and 25 frames >= hi with master == (lo + hi) / 2."""
    import torch
    from blackbox_b200 import reduce as bbr
    n, H, W = 50, 5280, 5280
    gen = torch.Generator(device='cuda')
    gen.manual_seed(5050)
    frames = [torch.randn((H, W), generator=gen, device='cuda', dtype=torch.float32) * 8.0 for _ in range(n)]
    frames[7][100:200] = frames[3][100:200]           # ties between frames
    master, _ = bbr.master_combine(frames, 'bias')
    rows = torch.arange(0, H, 64, device='cuda')
    sample = np.stack([f[rows].cpu().numpy() for f in frames])
    assert np.array_equal(master[rows].cpu().numpy(), np.median(sample, axis=0))
    le = torch.zeros((H, W), dtype=torch.int16, device='cuda')
    ge = torch.zeros((H, W), dtype=torch.int16, device='cuda')
    for f in frames:
        le += (f <= master).to(torch.int16)
        ge += (f >= master).to(torch.int16)
    assert int(le.min()) >= n // 2 and int(ge.min()) >= n // 2
    # row-stripe sharding (8 stripes of 660 rows) gives the same bits
    from blackbox_b200 import distributed as D
    for g in (0, 3, 7):
        r0, r1 = D.stripe_bounds(H, g, 8)
        assert (r0, r1) == (660 * g, 660 * (g + 1))
        part, _ = bbr.master_combine([f[r0:r1] for f in frames], 'bias')
        assert torch.equal(part, master[r0:r1])
