"""GPU parity: os_corr (kernels K1, F1, K2a-c, F2, K3) against the CPU oracle."""
import numpy as np
import pytest

from conftest import float_class_ok

pytestmark = pytest.mark.gpu


def _run(tel, small_bb, seed=1001, as_f32=False):
    import torch
    from blackbox_b200 import reduce as bbr, synth
    from oracle import reduce as R
    small_bb(200)
    raw, _ = synth.make_raw(tel, seed, nstars=400, ncosmics=50)
    if tel != 'ML1':
        # a few heavily saturated columns next to the horizontal overscan
        raw[150:200, 300:304] = 65535
        raw[250:300, 1500 + 40:1500 + 43] = 65535
    # oracle
    hdr_o = {}
    data_o = raw.astype(np.float32)
    R.gain_corr(data_o, hdr_o, tel=tel)
    diag = {}
    out_o = R.os_corr(data_o.copy(), hdr_o, 'object', tel=tel, diag=diag)
    # GPU
    hdr_g = {}
    if as_f32:
        data_g = torch.from_numpy(raw.astype(np.float32)).cuda()
        bbr.gain_corr(data_g, hdr_g, tel=tel)
        assert np.array_equal(data_g.cpu().numpy(), data_o)
        out_g, st = bbr.os_corr(data_g, hdr_g, 'object', tel=tel, return_state=True)
    else:
        out_g, st = bbr.os_corr(raw, hdr_g, 'object', tel=tel, return_state=True)
        out_g = torch.from_numpy(out_g)
    return raw, out_o, hdr_o, diag, out_g.cpu().numpy(), hdr_g, st


@pytest.mark.parametrize('tel', ['ML1', 'BG3', 'BG2'])
@pytest.mark.parametrize('as_f32', [False, True])
def test_os_corr_parity(tel, as_f32, small_bb):
    raw, out_o, hdr_o, diag, out_g, hdr_g, st = _run(tel, small_bb, as_f32=as_f32)
    ch = diag['chans']
    mean_vos = st.mean_vos.cpu().numpy()
    fit = st.vos_fit.cpu().numpy()
    for i in range(16):
        np.testing.assert_allclose(mean_vos[i], ch[i]['mean_vos_col'], rtol=1e-13, atol=0)
        np.testing.assert_allclose(fit[i], ch[i]['fit_vos_col'], rtol=1e-10, atol=0)
        assert bool(st.vfit_ok[i].item()) == ch[i]['polyfit_ok']
        np.testing.assert_allclose(st.vos_coef[i, :4].cpu().numpy(), ch[i]['p'][::-1], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(st.dlevel[i].item(), ch[i]['dlevel'], rtol=0, atol=1e-6)
        # column statistics: counts are selections (exact); means/stds may differ in the last
        # float32 bit where the fit vector differs in its last float64 bits
        n_g = st.hos_n[i].cpu().numpy()
        assert np.array_equal(n_g, ch[i]['nvalues'])
        if ch[i]['satcol'] is not None:
            assert np.array_equal(st.satcol[i].cpu().numpy().astype(bool), ch[i]['satcol'])
        v = n_g > 0
        np.testing.assert_allclose(st.hos_mean[i].cpu().numpy()[v], ch[i]['mean_hos'][v], rtol=0, atol=2e-5)
        v = n_g > 1
        np.testing.assert_allclose(st.hos_std[i].cpu().numpy()[v], ch[i]['std_hos'][v], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st.oscan[i].cpu().numpy(), ch[i]['oscan'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(st.std_vos.cpu().numpy(), diag['std_vos'], rtol=1e-9)
    np.testing.assert_allclose(st.biasm.cpu().numpy(), diag['mean_vos'], rtol=1e-11)
    for key in ('BIASMEAN', 'RDNOISE'):
        assert hdr_g[key] == pytest.approx(hdr_o[key], rel=1e-9)
    for i in range(16):
        assert hdr_g['VFITOK{}'.format(i + 1)] == hdr_o['VFITOK{}'.format(i + 1)]
    assert out_g.shape == out_o.shape and out_g.dtype == np.float32
    ok = float_class_ok(out_g, out_o, scale=hdr_o['BIASMEAN'])
    assert ok.all(), 'max abs diff {}'.format(np.abs(out_g - out_o).max())
    exact = np.mean(out_g == out_o)
    assert exact > 0.999, 'bit-exact fraction {}'.format(exact)


def test_reduce_apply_bit_exact(small_bb):
    """K3 alone is bit-exact when it is handed the oracle's own fit vectors."""
    import torch
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200.geometry import Geometry
    from oracle import reduce as R
    tel = 'BG3'
    small_bb(200)
    raw, _ = synth.make_raw(tel, 77, nstars=300, ncosmics=20)
    shape = (400, 10560)
    mbias, mflat, bpm = synth.make_masters(tel, 5, shape)
    mflat[10, 10] = 0.0             # division by zero must behave as numpy's
    hdr = {}
    data = raw.astype(np.float32)
    R.gain_corr(data, hdr, tel=tel)
    diag = {}
    out_o = R.os_corr(data, hdr, 'object', tel=tel, diag=diag)
    out_o -= mbias
    out_o[5, 7] = np.inf            # exercised below through the GPU path as well
    mask_o, _ = R.mask_init(out_o, hdr, bpm, 'object', tel=tel, diag=(d2 := {}))
    with np.errstate(divide='ignore', invalid='ignore'):
        out_o /= mflat
    # GPU with the oracle's vectors
    geom = Geometry.from_raw_shape(raw.shape, tel=tel)
    raw_t = bbr._to_dev(raw)
    st = bbr.OverscanState(geom, raw_t.device)
    st.vos_fit.copy_(torch.from_numpy(np.stack([c['fit_vos_col'] for c in diag['chans']])))
    st.oscan.copy_(torch.from_numpy(np.stack([c['oscan'] for c in diag['chans']])))
    sat = (np.array(set_bb.satlevel[tel]) * np.array(set_bb.gain[tel]) - diag['mean_vos'])
    st.satlevel.copy_(torch.from_numpy(sat))
    mb = mbias.copy()
    # reproduce the injected inf: make the bias subtraction overflow to +inf at (5, 7)
    mb[5, 7] = -np.inf
    img, mask = bbr.apply_enqueue(raw_t, geom, tel, st=st, gain=set_bb.gain[tel],
                                  mbias=bbr._to_dev(mb), mflat=bbr._to_dev(mflat),
                                  bpm=bbr._to_dev(bpm), want_mask=True)
    img = img.cpu().numpy()
    mask = mask.cpu().numpy()
    assert np.array_equal(img.view(np.uint32), out_o.view(np.uint32))
    mv = set_bb.mask_value
    seed_o = np.array(bpm, copy=True)
    seed_o[5, 7] |= mv['bad'] if bpm[5, 7] == 0 else 0
    seed_o[d2['mask_sat']] |= mv['saturated'] | 0x80     # 0x80: internal saturation marker
    assert np.array_equal(mask, seed_o)
