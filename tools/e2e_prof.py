#!/usr/bin/env python
"""Where the host spends its time in BatchReducer.run_host (compressed products on both sides):
cProfile over one batch after a warm-up batch.  Development aid.

    python tools/e2e_prof.py [--batch 32] [--depth 8]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--depth', type=int, default=8)
    ap.add_argument('--ahead', type=int, default=2)
    args = ap.parse_args()
    from blackbox_b200 import fitsio, reduce as R, set_bb, synth
    from blackbox_b200.pipeline import BatchReducer
    tel = 'BG3'
    raws = [R._to_dev(synth.make_raw(tel, 4001 + k)[0]) for k in range(4)]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    batch = BatchReducer(tel, tuple(raws[0].shape), depth=args.depth, ahead=args.ahead, mbias=mbias, mflat=mflat, bpm=bpm,
                         coeffs=synth.make_xtalk(3)[3], niter=4, use_graphs=True)
    ring = []
    for r in raws:
        heap, lens = R.rice_encode(r)
        offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
        info = dict(shape=tuple(r.shape), bitpix=16, bytepix=2, bzero=32768.0, bscale=1.0, blocksize=32)
        ring.append(fitsio.CompressedImage({}, torch.from_numpy(heap).pin_memory(), offs, lens.astype(np.int32), info))
        ring[-1].descriptors()
    nout = args.depth
    host_img = [torch.empty(batch.img_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(nout)]
    host_mask = [torch.empty(batch.mask_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(nout)]
    host_raw = [ring[k % len(ring)] for k in range(args.batch)]
    for _ in range(2):
        batch.run_host(host_raw, host_img, host_mask, mask_fz=True, img_fz=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    batch.run_host(host_raw, host_img, host_mask, mask_fz=True, img_fz=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('plain: {:.3f} ms per frame ({:.1f} frames/s)'.format(dt / args.batch * 1e3, args.batch / dt))
    prof = cProfile.Profile()
    prof.enable()
    batch.run_host(host_raw, host_img, host_mask, mask_fz=True, img_fz=True)
    torch.cuda.synchronize()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats('cumulative').print_stats(45)


if __name__ == '__main__':
    main()
