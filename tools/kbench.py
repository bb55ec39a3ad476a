#!/usr/bin/env python
"""Time single C-ABI calls of libbbx.so in isolation on full-size synthetic data (CUDA events,
L2 flushed between repetitions).  Development aid, not the benchmark.

    python tools/kbench.py [--reps 10] [--only vos_std,apply,...]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth  # noqa: E402
from blackbox_b200._lib import call  # noqa: E402
from blackbox_b200.pipeline import FramePipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    ap.add_argument('--tel', default='BG3')
    args = ap.parse_args()
    only = set(x for x in args.only.split(',') if x)
    tel = args.tel
    raw = synth.make_raw(tel, 4001)[0]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    coeffs = synth.make_xtalk(3)[3]
    raw_t = R._to_dev(raw)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=4)
    res = pipe.reduce(raw_t)
    img, mask = res.img, res.mask
    st, g = pipe.st, pipe.geom.as_struct()
    gain_h = R._harr([float(x) for x in pipe.gain], C.c_float)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    s = R._stream()
    H, W = img.shape

    def timeit(name, fn, nbytes=None, setup=None):
        if only and name not in only:
            return
        ts = []
        for _ in range(args.reps):
            if setup:
                setup()
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        extra = '' if nbytes is None else '  %7.1f GB/s (algorithmic)' % (nbytes / med / 1e6)
        print('%-22s median %8.4f ms  min %8.4f ms%s' % (name, med, ts[0], extra))

    npx = H * W
    timeit('vos_rowstats', lambda: call('bbx_vos_rowstats', R._ptr(raw_t), 0, C.byref(g), gain_h, 3.0, 5, R._ptr(st.mean_vos), s), 29.5e6)
    timeit('vos_std', lambda: call('bbx_vos_std', R._ptr(raw_t), 0, C.byref(g), gain_h, R._ptr(st.vos_fit), R._ptr(st.dlevel), R._ptr(st.std_vos), s), 29.5e6)
    timeit('overscan_all', lambda: R.overscan_enqueue(raw_t, pipe.geom, tel, gain=pipe.gain, state=st))
    oi, om = torch.empty_like(img), torch.empty_like(mask)

    def fresh_apply():
        R.apply_enqueue(raw_t, pipe.geom, tel, st=st, gain=pipe.gain, mbias=pipe.mbias, mflat=pipe.mflat,
                        bpm=pipe.bpm, want_mask=True, out_img=oi, out_mask=om, mwork=pipe.mwork)

    timeit('apply', lambda: R.apply_enqueue(raw_t, pipe.geom, tel, st=st, gain=pipe.gain, mbias=pipe.mbias, mflat=pipe.mflat,
                                            bpm=pipe.bpm, want_mask=True, out_img=oi, out_mask=om, mwork=pipe.mwork), 1815.6e6)
    timeit('mask_morph', lambda: R.mask_morph_enqueue(om, tel, pipe.mwork),
           setup=lambda: R.apply_enqueue(raw_t, pipe.geom, tel, st=st, gain=pipe.gain, mbias=pipe.mbias, mflat=pipe.mflat,
                                         bpm=pipe.bpm, want_mask=True, out_img=oi, out_mask=om, mwork=pipe.mwork))
    fresh_apply()
    R.mask_morph_enqueue(om, tel, pipe.mwork)
    x_img = img.clone()
    timeit('xtalk', lambda: R.xtalk_enqueue(x_img, mask, pipe.coeffs, tel), 1003.6e6)
    crm = torch.empty_like(mask)
    work = pipe.lwork

    def lac_setup():
        x_img.copy_(oi)

    def lac_all():
        R.lacosmic_enqueue(x_img, om, crm, 20.0, 0.01, 3.0, 0.0, 4, work, readnoise_dev=pipe.means[1:])

    def lac_begin():
        call('bbx_lacosmic_begin', R._ptr(x_img), R._ptr(om), R._ptr(crm), H, W, 4, 0, R._ptr(work.buf), R._ptr(work.info), s)

    def lac_it0():
        call('bbx_lacosmic_iteration', R._ptr(x_img), R._ptr(om), R._ptr(crm), H, W, 20.0, float(np.float32(0.01)), 3.0, 0.0,
             R._ptr(pipe.means[1:]), 0, 0, R._ptr(work.buf), R._ptr(work.info), s)

    timeit('lacosmic_4it', lac_all, 4460.5e6, setup=lac_setup)
    timeit('lacosmic_begin', lac_begin, setup=lac_setup)
    timeit('lacosmic_it0', lac_it0, 1115.1e6, setup=lambda: (lac_setup(), lac_begin()))
    if not only or only & {'stack20_bias', 'stack20_flat', 'stack20_clipped'}:
        gen = torch.Generator(device='cuda')
        gen.manual_seed(1)
        frames = [torch.randn((H, W), generator=gen, device='cuda') * 100 + 20000 for _ in range(20)]
        medsec = [20000.0 + 10 * k for k in range(20)]
        out = torch.empty((H, W), dtype=torch.float32, device='cuda')
        nb = 21 * H * W * 4
        timeit('stack20_bias', lambda: R.master_combine(frames, 'bias', out=out), nb)
        timeit('stack20_flat', lambda: R.master_combine(frames, 'flat', medsec=medsec, bpm=pipe.bpm, tel=tel, out=out), nb + H * W)
        timeit('stack20_clipped', lambda: R.master_combine(frames, 'bias', out=out, clip_sigma=3.0), nb)
        del frames
    R.tel = tel
    e_img = img.clone()
    timeit('channel_medians', lambda: R.channel_medians(e_img), 3 * 446.1e6)
    timeit('fill_edge_pixels', lambda: R.fill_edge_pixels(e_img, mask), 3 * 446.1e6 + 111.5e6)
    print('info', work.info.cpu().numpy())


if __name__ == '__main__':
    main()
