"""GPU checks at BASELINE.json's full size (10600 x 12000 raw -> 10560 x 10560 reduced).

The CPU oracle needs minutes per full frame for most steps, so the full-size checks are
  * the oracle itself where it is affordable (LACosmic in C with OpenMP; crosstalk and the fused
    per-pixel pass on row samples, which are exact sub-problems of the full frame), and
  * size-independent properties: the lazy / sparse kernels against their dense twins bit for bit,
    run-to-run determinism, rank properties of the stack median.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TEL = 'BG3'


@pytest.fixture(scope='module')
def full():
    import torch
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    assert (set_bb.ysize_chan, set_bb.xsize_chan) == (5280, 1320)
    raw, _ = synth.make_raw(TEL, 4001)
    shape = (10560, 10560)
    mbias, mflat, bpm = synth.make_masters(TEL, 9, shape)
    coeffs = synth.make_xtalk(3)[3]
    raw_t = bbr._to_dev(raw)
    # the chain up to (not including) LACosmic and crosstalk: their common input
    pre = FramePipeline(TEL, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=None, niter=0)
    res = pre.reduce(raw_t)
    d = dict(raw=raw, raw_t=raw_t, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs,
             img_pre=res.img.clone(), mask_pre=res.mask.clone(), hdr_pre=dict(res.header), st=pre.st,
             geom=pre.geom)
    yield d
    del d
    torch.cuda.empty_cache()


def test_fullsize_fused_pass_rows_against_numpy(full):
    """bbx_reduce_apply on the full frame: sampled rows restated with numpy in the reference's
    order and roundings (gain f32 multiply, blackbox.py:7460; vertical fit and overscan vector
    subtracted in float64 and rounded to float32, :6551, :6844; master bias :1679; non-finite
    -> 0, :4417; master flat :1825) from the device's own fit vectors: bit-exact."""
    from blackbox_b200 import set_bb
    g, st = full['geom'], full['st']
    gain = np.array(set_bb.get_par(set_bb.gain, TEL), dtype=np.float32)
    fit = st.vos_fit.cpu().numpy()
    oscan = st.oscan.cpu().numpy()
    raw, mbias, mflat = full['raw'], full['mbias'], full['mflat']
    img = full['img_pre']
    ysc, xsc, dy, dx = g.ysize_chan, g.xsize_chan, g.dy, g.dx
    rng = np.random.default_rng(5)
    for r in range(2):
        for ly in sorted(rng.choice(ysc, 24, replace=False).tolist() + [0, ysc - 1]):
            raw_row = g.data_y0[r] + ly
            trow = raw_row - r * dy
            want = np.empty(8 * xsc, dtype=np.float32)
            for c in range(8):
                i = 8 * r + c
                v = raw[raw_row, c * dx:c * dx + xsc].astype(np.float32) * gain[i]
                v = (v.astype(np.float64) - fit[i, trow]).astype(np.float32)
                v = (v.astype(np.float64) - oscan[i]).astype(np.float32)
                want[c * xsc:(c + 1) * xsc] = v
            rr = r * ysc + ly
            want -= mbias[rr]
            want[~np.isfinite(want)] = 0
            want /= mflat[rr]
            got = img[rr].cpu().numpy()
            assert np.array_equal(got, want, equal_nan=True), (r, ly)


def test_fullsize_xtalk_rows_against_oracle(full, small_bb):
    """A band of k tile rows of the bottom channels plus the mirrored band of the top channels
    is a closed sub-problem of xtalk_corr (blackbox.py:7138-7258): the oracle on that 2k-row
    frame must reproduce the full-frame kernel's rows."""
    import torch
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    img = full['img_pre'].clone()
    mask = full['mask_pre']
    bbr.xtalk_enqueue(img, mask, full['coeffs'], TEL)
    H, k = img.shape[0], 16
    for ly0 in (0, 1777, 5280 - k):
        sub = torch.cat([full['img_pre'][ly0:ly0 + k], full['img_pre'][H - ly0 - k:H - ly0]]).cpu().numpy()
        msub = torch.cat([mask[ly0:ly0 + k], mask[H - ly0 - k:H - ly0]]).cpu().numpy()
        small_bb(k)
        R.xtalk_corr(sub, full['coeffs'], msub, tel=TEL)
        got = torch.cat([img[ly0:ly0 + k], img[H - ly0 - k:H - ly0]]).cpu().numpy()
        assert np.allclose(got, sub, rtol=1e-6, atol=1e-6)
        assert np.mean(got == sub) > 0.999


def test_fullsize_lacosmic_against_oracle(full):
    """BASELINE.json config 3: detect_cosmics, 4 iterations, on the full 10560^2 reduced frame
    (reference settings otherwise: sigclip 20, sigfrac 0.01, objlim 3; set_blackbox.py:211-218)
    against the C oracle: mask and cleaned image bit for bit."""
    from blackbox_b200 import reduce as bbr
    from oracle import lacosmic as L
    rn = float(full['hdr_pre']['RDNOISE'])
    inmask = (full['mask_pre'] != 0)
    info_g, info_o = {}, {}
    kw = dict(sigclip=20, sigfrac=0.01, objlim=3, gain=1.0, readnoise=rn, satlevel=np.inf, niter=4,
              sepmed=False, cleantype='medmask')
    cr_g, clean_g = bbr.detect_cosmics(full['img_pre'], inmask=inmask, info=info_g, **kw)
    cr_o, clean_o = L.detect_cosmics(full['img_pre'].cpu().numpy(), inmask.cpu().numpy(), info=info_o, **kw)
    assert info_g['iterations'] == info_o['iterations']
    assert list(info_g['ncr_per_iter']) == list(info_o['ncr_per_iter'])
    assert cr_o.sum() > 1000
    assert np.array_equal(cr_g.cpu().numpy(), cr_o)
    assert np.array_equal(clean_g.cpu().numpy(), clean_o, equal_nan=True)


def test_fullsize_channel_medians_and_edge_fill(full):
    """np.median of each 5280 x 1320 channel (6 969 600 pixels) and the edge-pixel fill at the
    end of blackbox_reduce (blackbox.py:1958-1974) on the full frame against the oracle."""
    from blackbox_b200 import reduce as bbr
    from oracle import reduce as R
    want = full['img_pre'].cpu().numpy()
    mask = full['mask_pre'].cpu().numpy()
    meds_o = R.fill_edge_pixels(want, mask, tel=TEL)
    bbr.tel = TEL
    got = full['img_pre'].clone()
    meds = bbr.fill_edge_pixels(got, full['mask_pre'])
    assert np.array_equal(meds.cpu().numpy().view(np.uint32), meds_o.view(np.uint32))
    assert np.array_equal(got.cpu().numpy(), want)
    assert (mask & 32).sum() > 100000


def test_fullsize_chain_lazy_equals_dense_and_is_deterministic(full):
    """Whole chain, 4 LACosmic iterations: the lazy LACosmic / sparse morphology path and the
    dense kernels (every intermediate image materialised) give identical bits; a second run of
    the same frame through the same pipeline object reproduces them."""
    import torch
    from blackbox_b200 import reduce as bbr
    from blackbox_b200.pipeline import FramePipeline
    pipe = FramePipeline(TEL, full['raw'].shape, mbias=full['mbias'], mflat=full['mflat'], bpm=full['bpm'],
                         coeffs=full['coeffs'], niter=4)
    raw_t = full['raw_t']
    res = pipe.reduce(raw_t)
    assert not res.redo
    img1, mask1, hdr1 = res.img.clone(), res.mask.clone(), dict(res.header)
    assert int((mask1 & 0x80).sum()) == 0
    assert hdr1['LAC-NIT'] >= 2 and hdr1['NCOSMICS'] > 0 and hdr1['NOBJ-SAT'] > 0
    res2 = pipe.reduce(raw_t)
    assert torch.equal(res2.img, img1) and torch.equal(res2.mask, mask1)
    assert res2.header['NCOSMICS'] == hdr1['NCOSMICS']
    # dense twins
    img3 = torch.empty_like(img1)
    mask3 = torch.empty_like(mask1)
    pipe._overscan(raw_t)
    pipe._rest(raw_t, img3, mask3, dense_morph=True, lac_mode=bbr.LAC_DENSE)
    torch.cuda.synchronize()
    assert int(pipe.lwork.info[2].item()) == 0
    assert torch.equal(mask3, mask1)
    assert torch.equal(img3, img1)
    assert int(pipe.ncosmic.item()) / pipe.exptime == hdr1['NCOSMICS']
    assert int(pipe.mwork.nobj.item()) == hdr1['NOBJ-SAT']


def test_fullsize_master_flat_of_20(full):
    """Config 2 at full size: median of 20 normalised flats (8.9 GB resident).  np.median on a
    row sample (bit-exact); on the full frame the rank property of the median of an even
    count and invariance under a permutation of the frames."""
    import torch
    from blackbox_b200 import reduce as bbr
    n, H, W = 20, 10560, 10560
    gen = torch.Generator(device='cuda')
    gen.manual_seed(2020)
    resp = torch.from_numpy(full['mflat']).cuda()
    frames, medsec = [], []
    for i in range(n):
        level = 20000.0 * (1.0 + 0.2 * (i / (n - 1) - 0.5))
        f = resp * level
        f += torch.randn((H, W), generator=gen, device='cuda', dtype=torch.float32) * (level ** 0.5)
        frames.append(f)
        medsec.append(level)
    master, scales = bbr.master_combine(frames, 'flat', medsec=medsec, bpm=full['bpm'], tel=TEL)
    assert scales == [float(m) for m in medsec]
    rows = torch.arange(7, H, 160, device='cuda')
    cube = np.stack([(f[rows].cpu().numpy() / np.float32(m)) for f, m in zip(frames, medsec)])
    want = np.median(cube, axis=0)
    bpm_rows = full['bpm'][rows.cpu().numpy()]
    want[(bpm_rows == 32) | (want <= 0)] = 1
    assert np.array_equal(master[rows].cpu().numpy(), want)
    perm = [frames[i] for i in (np.random.default_rng(1).permutation(n))]
    pmed = [medsec[i] for i in (np.random.default_rng(1).permutation(n))]
    master2, _ = bbr.master_combine(perm, 'flat', medsec=pmed, bpm=full['bpm'], tel=TEL)
    assert torch.equal(master, master2)
    del perm, master2
    # rank property away from the pixels the flat rule overwrites
    free = (torch.from_numpy(full['bpm']).cuda() != 32) & (master != 1)
    le = torch.zeros((H, W), dtype=torch.int16, device='cuda')
    ge = torch.zeros((H, W), dtype=torch.int16, device='cuda')
    for f, m in zip(frames, medsec):
        v = f / torch.tensor(m, dtype=torch.float32, device='cuda')     # a true float32 division
        le += (v <= master).to(torch.int16)
        ge += (v >= master).to(torch.int16)
    assert int(le[free].min()) >= n // 2 and int(ge[free].min()) >= n // 2


def test_fullsize_config1_meerlicht_chain_against_oracle():
    """BASELINE.json config 1 at full size: one MeerLICHT 10600 x 12000 raw frame through gain +
    overscan + mask_init + master flat + crosstalk (ML1 subtracts no master bias,
    set_blackbox.py:37) against the oracle's whole-frame result."""
    import torch
    from conftest import float_class_ok
    from blackbox_b200 import set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    tel = 'ML1'
    raw, _ = synth.make_raw(tel, 1001)
    shape = (10560, 10560)
    mbias, mflat, bpm = synth.make_masters(tel, 1002, shape)
    coeffs = synth.make_xtalk(1003)[3]
    data_o, mask_o, hdr_o, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs,
                                              steps=('gain', 'os', 'bias', 'mask', 'flat', 'xtalk'))
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=0)
    res = pipe.reduce(raw)
    mask = res.mask.cpu().numpy()
    assert np.array_equal(mask, mask_o)
    assert hdr_o['NOBJ-SAT'] > 10 and res.header['NOBJ-SAT'] == hdr_o['NOBJ-SAT']
    for key in ('BIASMEAN', 'RDNOISE', 'SATURATE'):
        assert res.header[key] == pytest.approx(hdr_o[key], rel=1e-9)
    img = res.img.cpu().numpy()
    del res, pipe
    torch.cuda.empty_cache()
    assert float_class_ok(img, data_o, scale=hdr_o['BIASMEAN']).all()
    assert np.mean(img == data_o) > 0.999


def test_fullsize_config4_blackgem_chain_against_oracle(full):
    """The bench's own frame: one full-size BlackGEM (BG3, seed 4001) 10600 x 12000 raw frame
    through the WHOLE chain -- gain, overscan, master bias, mask_init, master flat, LACosmic with
    4 iterations, crosstalk (BASELINE.json config 4) -- against the oracle's whole-frame result:
    mask bit for bit, header counts, image in the float class with > 99.9 % identical pixels."""
    import torch
    from conftest import float_class_ok
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    data_o, mask_o, hdr_o, hm_o = R.reduce_frame(full['raw'], TEL, full['mbias'], full['mflat'], full['bpm'],
                                                 full['coeffs'], exptime=45.0, niter=4)
    R.mask_header(mask_o, hm_o, tel=TEL)
    pipe = FramePipeline(TEL, full['raw'].shape, mbias=full['mbias'], mflat=full['mflat'], bpm=full['bpm'],
                         coeffs=full['coeffs'], niter=4)
    pipe.enqueue(full['raw_t'], exptime=45.0)
    res = pipe.finish()
    assert not res.redo
    mask = res.mask.cpu().numpy()
    assert np.array_equal(mask, mask_o)
    assert hdr_o['NOBJ-SAT'] > 10 and res.header['NOBJ-SAT'] == hdr_o['NOBJ-SAT']
    assert hdr_o['NCOSMICS'] > 1 and res.header['NCOSMICS'] == hdr_o['NCOSMICS']
    assert ({k: int(v) for k, v in res.header_mask.items() if k.endswith('NUM')}
            == {k: int(v) for k, v in hm_o.items() if k.endswith('NUM')})
    for key in ('BIASMEAN', 'RDNOISE', 'SATURATE'):
        assert res.header[key] == pytest.approx(hdr_o[key], rel=1e-9)
    for i in range(16):
        for key in ('BIASM{}'.format(i + 1), 'RDN{}'.format(i + 1)):
            assert res.header[key] == pytest.approx(hdr_o[key], rel=1e-8), key
    img = res.img.cpu().numpy()
    del res, pipe
    torch.cuda.empty_cache()
    assert float_class_ok(img, data_o, scale=hdr_o['BIASMEAN']).all()
    assert np.mean(img == data_o) > 0.999


def test_fullsize_batch_reducer_equals_sequential(full):
    """Full-size frames through BatchReducer (4 pipelines in flight, overscan stage two frames
    ahead on high-priority streams, CUDA-graph replay) and through run_host: the bits of the
    one-by-one pipeline, pass after pass."""
    import torch
    from blackbox_b200.pipeline import BatchReducer, FramePipeline
    kw = dict(mbias=full['mbias'], mflat=full['mflat'], bpm=full['bpm'], coeffs=full['coeffs'], niter=4)
    base = full['raw_t']
    gen = torch.Generator(device='cuda')
    gen.manual_seed(77)
    raws = [base]
    for _ in range(2):                                   # two more noise realisations of the frame
        r = (base.view(torch.int16).to(torch.int32) & 0xffff) + torch.randint(-3, 4, base.shape, device='cuda',
                                                                             generator=gen, dtype=torch.int32)
        raws.append(r.clamp_(0, 65535).to(torch.int16).view(torch.uint16).contiguous())
    single = FramePipeline(TEL, full['raw'].shape, **kw)
    want = []
    for r in raws:
        res = single.reduce(r)
        want.append((res.img.clone(), res.mask.clone()))
    del single
    order = [0, 1, 2, 1, 0, 2, 2]
    frames = [raws[i] for i in order]
    batch = BatchReducer(TEL, full['raw'].shape, depth=4, ahead=2, use_graphs=True, **kw)
    imgs = [torch.empty_like(want[0][0]) for _ in order]
    masks = [torch.empty_like(want[0][1]) for _ in order]
    for rep in range(3):
        results = batch.run(frames, imgs, masks)
        torch.cuda.synchronize()
        for k, i in enumerate(order):
            assert not results[k].redo
            assert torch.equal(imgs[k], want[i][0]), (rep, k)
            assert torch.equal(masks[k], want[i][1]), (rep, k)
    assert sum(p.graph_replays for p in batch.pipes) > 0
    del imgs, masks
    host_raws = [r.cpu().pin_memory() for r in raws]
    host_imgs = [torch.empty(want[0][0].shape, dtype=torch.float32).pin_memory() for _ in raws]
    host_masks = [torch.empty(want[0][1].shape, dtype=torch.uint8).pin_memory() for _ in raws]
    for rep in range(2):
        batch.run_host(host_raws, host_imgs, host_masks)
        for i in range(len(raws)):
            assert torch.equal(host_imgs[i], want[i][0].cpu()), (rep, i)
            assert torch.equal(host_masks[i], want[i][1].cpu()), (rep, i)
