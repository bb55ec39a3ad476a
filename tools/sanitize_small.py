#!/usr/bin/env python
"""A small tour of every C-ABI entry point on tiny inputs (odd sizes, both telescopes, binned
frames, dense and lazy twins, graphs, host buffers).  Runs in a few seconds; where
compute-sanitizer is allowed (it is closed on the pool this was developed on) run it as

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth  # noqa: E402
from blackbox_b200.pipeline import BatchReducer, FramePipeline  # noqa: E402


def main():
    from scipy import interpolate
    ysc = 64
    set_bb.ysize_chan = ysc
    set_bb.hos_sat_ypix_lim = {'BG2': (32, 64), 'BG3': (16, 32), 'BG4': (16, 32)}
    rng = np.random.default_rng(0)
    for tel in ('BG3', 'ML1'):
        R.tel = tel
        raw, _ = synth.make_raw(tel, 7, nstars=60, ncosmics=40)
        raw[20:26, 2000:2006] = 65535
        raw[40:64, 1500 * 2 + 20:1500 * 2 + 24] = 65535
        shape = (2 * ysc, 8 * set_bb.xsize_chan)
        mbias, mflat, bpm = synth.make_masters(tel, 8, shape)
        coeffs = synth.make_xtalk(9)[3]
        pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=3, fill_edge=True)
        res = pipe.reduce(raw)
        # dense twins
        img, mask = torch.empty_like(res.img), torch.empty_like(res.mask)
        raw_t = R._to_dev(raw)
        pipe._overscan(raw_t)
        pipe._rest(raw_t, img, mask, dense_morph=True, lac_mode=R.LAC_DENSE)
        pipe._rest(raw_t, img, mask, lac_mode=R.LAC_LAZY_BG)
        torch.cuda.synchronize()
        # drop-in functions one by one
        hdr = {}
        out = R.os_corr(raw, hdr, 'object', tel=tel)
        data = raw.astype(np.float32)
        R.gain_corr(data, {}, tel=tel)
        hdr['EXPTIME'] = 60.0
        m, hm = R.mask_init(out.copy(), hdr, 'q', 'object', bpm=bpm)
        R.MASK_MORPH_SPARSE = False
        R.mask_init(out.copy(), hdr, 'q', 'object', bpm=bpm)
        R.MASK_MORPH_SPARSE = True
        R.mask_header(m, hm)
        d2, m2 = R.cosmics_corr(out.copy(), hdr, m.copy(), hm)
        R.xtalk_corr(d2, coeffs, m2)
        R.xtalk_corr(d2[:, :], coeffs, None)
        R.fill_edge_pixels(d2, m2)
        R.channel_medians(d2, ignore_nan=True)
        R.subtract_mbias(d2, mbias)
        R.divide_mflat(d2, mflat)
        batch = BatchReducer(tel, raw.shape, depth=3, use_graphs=True, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=2)
        raws = [R._to_dev(raw) for _ in range(4)]
        imgs = [torch.empty_like(res.img) for _ in raws]
        masks = [torch.empty_like(res.mask) for _ in raws]
        for _ in range(3):
            batch.run(raws, imgs, masks)
        hr = [r.cpu().pin_memory() for r in raws]
        hi = [torch.empty(res.img.shape, dtype=torch.float32).pin_memory() for _ in raws]
        hm_ = [torch.empty(res.mask.shape, dtype=torch.uint8).pin_memory() for _ in raws]
        batch.run_host(hr, hi, hm_)
    # binned geometry
    set_bb.ysize_chan = 128
    set_bb.hos_sat_ypix_lim = {'BG2': (32, 64), 'BG3': (16, 32), 'BG4': (16, 32)}
    rawb, _ = synth.make_raw('BG3', 3, ysize_chan=64, xsize_chan=660, os_rows=10, os_cols=90, nstars=30, ncosmics=10)
    R.os_corr(rawb, {}, 'object', xbin=2, ybin=2, tel='BG3')
    # masters, odd sizes, non-multiple-of-4 widths
    for n in (1, 2, 5, 20, 33, 64):
        frames = [torch.from_numpy(rng.standard_normal((37, 53)).astype(np.float32) + 100).cuda() for _ in range(n)]
        R.master_combine(frames, 'bias')
        R.master_combine(frames, 'flat', medsec=[100.0] * n, bpm=np.zeros((37, 53), np.uint8), tel='BG3')
        R.master_combine(frames, 'bias', clip_sigma=3.0)
    small = rng.standard_normal((50, 72)).astype(np.float32) * 10 + 300
    small[10, 10:13] += 5000
    for mode in (R.LAC_LAZY_BG, R.LAC_DENSE, None):
        R.detect_cosmics(small, inmask=None, sigclip=5, sigfrac=0.3, objlim=2, niter=3, readnoise=5.0, satlevel=np.inf,
                         sepmed=False, cleantype='medmask', mode=mode)
    set_bb.ysize_chan, set_bb.xsize_chan = 33, 41
    odd = rng.standard_normal((66, 328)).astype(np.float32)
    R.tel = 'BG3'
    R.channel_medians(odd)
    R.fill_edge_pixels(odd, np.full((66, 328), 32, np.uint8))
    spl = [interpolate.UnivariateSpline(np.linspace(0, 60000, 50), 1e-3 * rng.standard_normal(50), k=3, s=1e-4) for _ in range(16)]
    R.nonlin_corr(np.abs(odd) * 40000, spl)
    be, bp = R.fits_encode(torch.from_numpy(odd).cuda())
    R.fits_decode(be, dict(bitpix=bp, shape=odd.shape, bzero=0.0, bscale=1.0))
    u16 = torch.from_numpy(rng.integers(0, 65536, (7, 9), dtype=np.uint16).view(np.int16)).cuda().view(torch.uint16)
    be, bp = R.fits_encode(u16)
    R.fits_decode(be, dict(bitpix=16, shape=(7, 9), bzero=32768.0, bscale=1.0))
    torch.cuda.synchronize()
    print('sanitize tour done')


if __name__ == '__main__':
    main()
