"""GPU parity of the whole chain (FramePipeline) against oracle.reduce_frame."""
import numpy as np
import pytest

from conftest import float_class_ok

pytestmark = pytest.mark.gpu


def _inputs(tel, seed, ysc):
    from blackbox_b200 import synth
    raw, truth = synth.make_raw(tel, seed, nstars=400, ncosmics=150)
    shape = (2 * ysc, 10560)
    mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
    _, _, _, coeffs = synth.make_xtalk(seed + 2)
    return raw, mbias, mflat, bpm, coeffs


@pytest.mark.parametrize('tel,niter', [('ML1', 3), ('BG3', 4), ('BG2', 1)])
def test_full_chain_parity(tel, niter, small_bb):
    from blackbox_b200.pipeline import FramePipeline
    from blackbox_b200 import set_bb
    from oracle import reduce as R
    ysc = 160
    small_bb(ysc)
    raw, mbias, mflat, bpm, coeffs = _inputs(tel, 4001, ysc)
    raw[40:48, 2000:2008] = 65535                     # a saturated blob with neighbours
    data_o, mask_o, hdr_o, hm_o = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs,
                                                 exptime=60.0, niter=niter)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs,
                         niter=niter, exptime=60.0)
    res = pipe.reduce(raw)
    img, mask = res.img.cpu().numpy(), res.mask.cpu().numpy()
    mv = set_bb.mask_value
    # masks: bit-exact except documented threshold ties (<= 1e-5 of the pixels)
    mism = np.mean(mask != mask_o)
    assert mism <= 1e-5, 'mask mismatch fraction {}'.format(mism)
    for name in ('saturated', 'saturated-connected', 'crosstalk', 'bad', 'edge'):
        assert np.array_equal(mask & mv[name], mask_o & mv[name]), name
    assert (mask & 0x80).sum() == 0
    # float output
    same_cr = (mask & mv['cosmic ray']) == (mask_o & mv['cosmic ray'])
    ok = float_class_ok(img, data_o, scale=hdr_o['BIASMEAN'])
    assert ok[same_cr].all(), np.abs(img - data_o)[same_cr].max()
    assert np.mean(img == data_o) > 0.999
    for key in ('BIASMEAN', 'RDNOISE', 'SATURATE'):
        assert res.header[key] == pytest.approx(hdr_o[key], rel=1e-9)
    assert res.header['NOBJ-SAT'] == hdr_o['NOBJ-SAT']
    if mism == 0:
        assert res.header['NCOSMICS'] == pytest.approx(hdr_o['NCOSMICS'])
        # mask_header keywords (blackbox.py:4601-4620): pixel counts per mask type
        for name, short in (('bad', 'BP'), ('edge', 'EP'), ('saturated', 'SP'), ('saturated-connected', 'SCP'),
                            ('cosmic ray', 'CRP')):
            assert res.header_mask['M-{}NUM'.format(short)] == int(((mask_o & mv[name]) != 0).sum()), name
            assert res.header_mask['M-{}VAL'.format(short)] == mv[name]
    # a second frame through the same pipeline object (buffers are reused)
    raw2, *_ = _inputs(tel, 4002, ysc)
    data_o2, mask_o2, _, _ = R.reduce_frame(raw2, tel, mbias, mflat, bpm, coeffs, niter=niter)
    res2 = pipe.reduce(raw2)
    assert np.mean(res2.mask.cpu().numpy() != mask_o2) <= 1e-5
    assert np.mean(res2.img.cpu().numpy() == data_o2) > 0.999


def test_chain_spline_redo(small_bb):
    """Columns < 150 with saturated pixels above them (BlackGEM) need the host spline: the
    pipeline reads the device flags after the overscan stage and patches those columns."""
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    tel, ysc = 'BG3', 160
    small_bb(ysc)
    raw, mbias, mflat, bpm, coeffs = _inputs(tel, 4100, ysc)
    raw[100:160, 1500 * 2 + 20:1500 * 2 + 24] = 65535       # channel 3, columns 20..23, next to hos
    diag = {}
    data_o, mask_o, hdr_o, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs, niter=2, diag=diag)
    assert diag['chans'][2]['spline_needed'].sum() >= 4
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=2)
    res = pipe.reduce(raw)
    assert res.spline_columns >= 4 and not res.redo
    assert np.mean(res.mask.cpu().numpy() != mask_o) <= 1e-5
    img = res.img.cpu().numpy()
    assert float_class_ok(img, data_o, scale=hdr_o['BIASMEAN']).mean() > 0.99999
    assert np.mean(img == data_o) > 0.999


@pytest.mark.parametrize('depth,graphs,prio', [(2, False, False), (3, False, True), (3, True, True)])
def test_batch_reducer_equals_single_pipeline(depth, graphs, prio, small_bb):
    """Several frames in flight (stage A run ahead on a high-priority stream, stages replayed as
    CUDA graphs) give the bits of the one-by-one pipeline; so does the host-buffer entry."""
    import torch
    from blackbox_b200 import reduce as bbr
    from blackbox_b200.pipeline import BatchReducer, FramePipeline
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw0, mbias, mflat, bpm, coeffs = _inputs(tel, 4200, ysc)
    raws = [raw0] + [_inputs(tel, 4200 + i, ysc)[0] for i in (1, 2, 3, 4)]
    raws[2][100:120, 1500 * 2 + 20:1500 * 2 + 24] = 65535     # frame 2 needs the host spline
    kw = dict(mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=2)
    single = FramePipeline(tel, raw0.shape, **kw)
    want = []
    for r in raws:
        res = single.reduce(r)
        want.append((res.img.cpu().numpy().copy(), res.mask.cpu().numpy().copy(), res.header['NCOSMICS'],
                     res.spline_columns))
    assert want[2][3] >= 4
    batch = BatchReducer(tel, raw0.shape, depth=depth, split_priority=prio, use_graphs=graphs, **kw)
    raws_t = [bbr._to_dev(r) for r in raws]
    imgs = [torch.empty((2 * ysc, 10560), dtype=torch.float32, device='cuda') for _ in raws]
    masks = [torch.empty((2 * ysc, 10560), dtype=torch.uint8, device='cuda') for _ in raws]
    for rep in range(3):                      # graphs are captured on the second pass and replayed on the third
        for t in imgs + masks:
            t.zero_()
        results = batch.run(raws_t, imgs, masks, fill_header=True)
        torch.cuda.synchronize()
        for k, (img, mask, nc, ncols) in enumerate(want):
            assert np.array_equal(imgs[k].cpu().numpy(), img, equal_nan=True), (rep, k)
            assert np.array_equal(masks[k].cpu().numpy(), mask), (rep, k)
            assert results[k].header['NCOSMICS'] == nc
            assert results[k].spline_columns == ncols
    if graphs:
        assert sum(p.graph_replays for p in batch.pipes) > 0
    # host buffers in, host buffers out
    host_raws = [torch.from_numpy(r.view(np.int16)).view(torch.uint16).pin_memory() for r in raws]
    host_imgs = [torch.empty((2 * ysc, 10560), dtype=torch.float32).pin_memory() for _ in raws]
    host_masks = [torch.empty((2 * ysc, 10560), dtype=torch.uint8).pin_memory() for _ in raws]
    for rep in range(3):
        results = batch.run_host(host_raws, host_imgs, host_masks, fill_header=True)
        for k, (img, mask, nc, _) in enumerate(want):
            assert np.array_equal(host_imgs[k].numpy(), img, equal_nan=True), (rep, k)
            assert np.array_equal(host_masks[k].numpy(), mask), (rep, k)
            assert results[k].header['NCOSMICS'] == nc


def test_run_host_with_fits_data_units(small_bb, tmp_path):
    """Raw FITS file -> pinned big-endian bytes -> run_host(fits=True) -> big-endian float32 bytes
    -> reduced FITS file: the file reads back to the image the plain pipeline produces."""
    import torch
    from blackbox_b200 import fitsio
    from blackbox_b200.pipeline import BatchReducer, FramePipeline
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw0, mbias, mflat, bpm, coeffs = _inputs(tel, 4300, ysc)
    raws = [raw0, _inputs(tel, 4301, ysc)[0], _inputs(tel, 4302, ysc)[0]]
    kw = dict(mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=2)
    single = FramePipeline(tel, raw0.shape, **kw)
    want = []
    for r in raws:
        res = single.reduce(r)
        want.append((res.img.cpu().numpy().copy(), res.mask.cpu().numpy().copy()))
    host_raws = []
    for k, r in enumerate(raws):
        path = str(tmp_path / 'raw{}.fits'.format(k))
        fitsio.write_primary(path, r, {'EXPTIME': 60.0})
        _, buf, info = fitsio.read_primary(path, pinned=True)
        assert info['bitpix'] == 16 and info['bzero'] == 32768.0
        host_raws.append(buf.view(torch.uint16).view(info['shape']))
    shape = (2 * ysc, 10560)
    host_imgs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in raws]
    host_masks = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in raws]
    batch = BatchReducer(tel, raw0.shape, depth=3, **kw)
    for rep in range(2):
        batch.run_host(host_raws, host_imgs, host_masks, fits=True)
        for k, (img, mask) in enumerate(want):
            out = str(tmp_path / 'red{}.fits'.format(k))
            fitsio.write_primary(out, host_imgs[k].view(torch.uint8).reshape(-1), {'REDFILE': 'x'}, be_bytes=True,
                                 shape=shape, bitpix=-32)
            _, data, info = fitsio.read_primary(out)
            assert np.array_equal(fitsio.to_native(data, info), img, equal_nan=True), (rep, k)
            assert np.array_equal(host_masks[k].numpy(), mask), (rep, k)


def test_pipeline_edge_fill_option(small_bb):
    """fill_edge=True appends the edge-pixel fill of blackbox.py:1958-1974 to the chain."""
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw, mbias, mflat, bpm, coeffs = _inputs(tel, 4400, ysc)
    data_o, mask_o, _, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs, niter=2)
    meds_o = R.fill_edge_pixels(data_o, mask_o, tel=tel)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=2, fill_edge=True)
    res = pipe.reduce(raw)
    img, mask = res.img.cpu().numpy(), res.mask.cpu().numpy()
    assert np.array_equal(mask, mask_o)
    edge = (mask & 32) != 0
    assert edge.sum() > 1000
    np.testing.assert_allclose(pipe.chan_med.cpu().numpy(), meds_o, rtol=1e-5)
    assert np.mean(img == data_o) > 0.999
    assert np.allclose(img[edge], data_o[edge], rtol=1e-5)


def _bimodal_raw(tel, seed, ysc):
    """A frame whose median sits between two populations (sky / bright nebula half) with a fat
    cosmic-ray blob: the lazy LACosmic needs the background level and its sampled bracket misses."""
    from blackbox_b200 import set_bb, synth
    sky, _ = synth.make_sky(tel, seed, ysc, set_bb.xsize_chan, nstars=50, ncosmics=40)
    half = int(sky.shape[1] * 0.505)              # the median rank sits just inside the dark half: the bracket straddles both
    rng = np.random.default_rng(seed)
    sky[:, half:] += 9000.0 + 60.0 * rng.standard_normal((sky.shape[0], sky.shape[1] - half)).astype(np.float32)
    sky[70:83, 300:313] += rng.uniform(8000, 20000, (13, 13)).astype(np.float32)
    raw, _ = synth.make_raw(tel, seed, sky=sky)
    return raw


def test_chain_redo_when_the_lazy_lacosmic_needs_the_background(small_bb):
    """FramePipeline.finish repeats the stage-B chain with the background level computed up
    front when the lazy path reports that it needs it; the result is the oracle's.  The same
    frame in the middle of a BatchReducer run (graphs on) is redone without disturbing its
    neighbours."""
    import torch
    from blackbox_b200 import reduce as bbr
    from blackbox_b200.pipeline import BatchReducer, FramePipeline
    from oracle import reduce as R
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw0, mbias, mflat, bpm, coeffs = _inputs(tel, 4500, ysc)
    bad = _bimodal_raw(tel, 4501, ysc)
    kw = dict(mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=3)
    data_o, mask_o, hdr_o, _ = R.reduce_frame(bad, tel, mbias, mflat, bpm, coeffs, niter=3)
    pipe = FramePipeline(tel, bad.shape, **kw)
    res = pipe.reduce(bad)
    assert res.redo
    assert np.mean(res.mask.cpu().numpy() != mask_o) <= 1e-5
    assert np.mean(res.img.cpu().numpy() == data_o) > 0.999
    want_bad = (res.img.cpu().numpy().copy(), res.mask.cpu().numpy().copy())
    res0 = pipe.reduce(raw0)
    assert not res0.redo
    want0 = (res0.img.cpu().numpy().copy(), res0.mask.cpu().numpy().copy())
    raws = [bbr._to_dev(r) for r in (raw0, bad, raw0, raw0, bad, raw0)]
    batch = BatchReducer(tel, raw0.shape, depth=4, use_graphs=True, **kw)
    imgs = [torch.empty_like(res.img) for _ in raws]
    masks = [torch.empty_like(res.mask) for _ in raws]
    for rep in range(3):
        results = batch.run(raws, imgs, masks)
        torch.cuda.synchronize()
        for k, r in enumerate(results):
            want = want_bad if k in (1, 4) else want0
            assert r.redo == (k in (1, 4)), (rep, k)
            assert np.array_equal(imgs[k].cpu().numpy(), want[0], equal_nan=True), (rep, k)
            assert np.array_equal(masks[k].cpu().numpy(), want[1]), (rep, k)


def test_reduce_night_tool_files_in_files_out(small_bb, tmp_path):
    """tools/reduce_night.py: raw FITS files in, _red.fits / _mask.fits out, equal to the
    oracle's chain; header keywords of the reduction steps end up in the files."""
    import importlib.util
    import os
    from blackbox_b200 import fitsio, synth
    from oracle import reduce as R
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw_dir, out_dir = tmp_path / 'raw', tmp_path / 'red'
    raw_dir.mkdir()
    raws = [_inputs(tel, 4600 + k, ysc)[0] for k in range(3)]
    _, mbias, mflat, bpm, coeffs = _inputs(tel, 4600, ysc)
    for k, r in enumerate(raws):
        fitsio.write_primary(str(raw_dir / 'BG3_2026_{:02d}.fits'.format(k)), r, {'EXPTIME': 60.0, 'FILTER': 'q'})
    fitsio.write_primary(str(tmp_path / 'mbias.fits'), mbias)
    fitsio.write_primary(str(tmp_path / 'mflat.fits'), mflat)
    fitsio.write_primary(str(tmp_path / 'bpm.fits'), bpm)
    victim, source, corr, _ = synth.make_xtalk(4602)
    synth.write_xtalk_file(str(tmp_path / 'xtalk.txt'), victim, source, corr)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('reduce_night', os.path.join(here, 'tools', 'reduce_night.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    n = tool.main([str(raw_dir), str(out_dir), '--tel', tel, '--mbias', str(tmp_path / 'mbias.fits'),
                   '--mflat', str(tmp_path / 'mflat.fits'), '--bpm', str(tmp_path / 'bpm.fits'),
                   '--xtalk', str(tmp_path / 'xtalk.txt'), '--niter', '2'])
    assert n == 3
    for k, r in enumerate(raws):
        data_o, mask_o, hdr_o, _ = R.reduce_frame(r, tel, mbias, mflat, bpm, coeffs, niter=2)
        h, d, info = fitsio.read_primary(str(out_dir / 'BG3_2026_{:02d}_red.fits'.format(k)))
        img = fitsio.to_native(d, info)
        _, m, _ = fitsio.read_primary(str(out_dir / 'BG3_2026_{:02d}_mask.fits'.format(k)))
        assert np.mean(np.asarray(m) != mask_o) <= 1e-5
        assert np.mean(img == data_o) > 0.999
        assert h['FILTER'][0] == 'q' and h['REDFILE'][0].endswith('_red')
        assert h['BIASMEAN'][0] == pytest.approx(hdr_o['BIASMEAN'], rel=1e-9)
        assert h['NOBJ-SAT'][0] == hdr_o['NOBJ-SAT']


def test_pipeline_refuses_mismatched_buffers(small_bb):
    """The C ABI takes bare pointers; the host layer must reject wrong shapes / dtypes / devices
    before anything is launched."""
    import torch
    from blackbox_b200 import reduce as bbr
    from blackbox_b200.pipeline import FramePipeline
    tel, ysc = 'BG3', 120
    small_bb(ysc)
    raw, mbias, mflat, bpm, coeffs = _inputs(tel, 4700, ysc)
    with pytest.raises(ValueError):
        FramePipeline(tel, raw.shape, mbias=mbias[:-1], mflat=mflat, bpm=bpm, coeffs=coeffs)
    with pytest.raises(ValueError):
        FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs[:8])
    with pytest.raises(ValueError):
        FramePipeline(tel, (raw.shape[0] + 1, raw.shape[1]), mbias=mbias, mflat=mflat, bpm=bpm)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=1)
    raw_t = bbr._to_dev(raw)
    with pytest.raises(ValueError):
        pipe.enqueue(raw_t[:-2])
    with pytest.raises(TypeError):
        pipe.enqueue(raw_t.to(torch.int32))
    with pytest.raises(ValueError):
        pipe.enqueue(raw_t, torch.empty((2 * ysc, 10560), dtype=torch.float64, device='cuda'), None)
    with pytest.raises(ValueError):
        pipe.enqueue(raw_t, torch.empty((2 * ysc, 10560), dtype=torch.float32, device='cuda').t().contiguous().t(), None)
    res = pipe.reduce(raw)                      # and it still works afterwards
    assert res.img.shape == (2 * ysc, 10560)


@pytest.mark.parametrize('tel', ['BG3', 'ML1'])
def test_fused_scan_equals_separate_passes_with_rings_next_to_cosmic_rays(tel, small_bb):
    """The dense Laplacian scan fused into the per-pixel pass (bbx_reduce_apply_scan) sees the image
    before the mask morphology has run: saturated rings whose holes get filled, crosstalk victims and
    saturated-connected pixels are masked AFTER the background statistics were taken, and the
    morphology has to take them out again.  Frame: rings (closed, broken, with an island) with fat
    cosmic-ray blobs right next to and inside them, so that the cleaning needs the background level
    (blob interiors have no usable neighbour).  The fused pipeline must give the bits of the
    pipeline with separate passes and of the oracle."""
    import torch
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    from oracle import reduce as R
    ysc = 200
    small_bb(ysc)
    raw, _ = synth.make_raw(tel, 7300, nstars=200, ncosmics=80)
    synth.add_saturated_rings(raw)
    for (cy, cx) in ((60, 700), (120, 4000), (150, 9100), (300, 2500), (90, 6000), (200, 11000)):
        raw[cy - 4:cy + 5, cx + 16:cx + 25] = np.minimum(raw[cy - 4:cy + 5, cx + 16:cx + 25].astype(np.int64) + 9000, 30000)
        raw[cy - 2:cy + 3, cx - 2:cx + 3] = np.minimum(raw[cy - 2:cy + 3, cx - 2:cx + 3].astype(np.int64) + 12000, 30000)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 7301, shape)
    coeffs = synth.make_xtalk(7302)[3]
    kw = dict(mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=4)
    data_o, mask_o, hdr_o, hm_o = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs, niter=4)
    cr = (mask_o & 2) != 0
    fused = FramePipeline(tel, raw.shape, fuse_scan=True, **kw)
    plain = FramePipeline(tel, raw.shape, fuse_scan=False, stats_in_apply=False, **kw)
    stats = FramePipeline(tel, raw.shape, fuse_scan=False, stats_in_apply=True, **kw)
    rf, rp, rs = fused.reduce(raw), plain.reduce(raw), stats.reduce(raw)
    assert rf.redo == rp.redo == rs.redo
    assert torch.equal(rf.mask, rp.mask) and torch.equal(rf.img, rp.img)
    assert torch.equal(rs.mask, rp.mask) and torch.equal(rs.img, rp.img)
    assert rf.header == rp.header and rf.header_mask == rp.header_mask
    assert rs.header == rp.header and rs.header_mask == rp.header_mask
    rs2 = stats.reduce(raw)
    assert torch.equal(rs2.mask, rp.mask) and torch.equal(rs2.img, rp.img)
    assert np.array_equal(rf.mask.cpu().numpy(), mask_o)
    img = rf.img.cpu().numpy()
    assert np.mean(img == data_o) > 0.999
    assert np.array_equal(img[cr], data_o[cr])                   # the cleaned pixels, background level included
    assert (mask_o & 8).sum() > 50 and cr.sum() > 300
    # a second frame through the same objects (list / statistics state is per frame)
    rf2 = fused.reduce(raw)
    assert torch.equal(rf2.mask, rf.mask) and torch.equal(rf2.img, rf.img)
