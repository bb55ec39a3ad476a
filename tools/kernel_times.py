#!/usr/bin/env python
"""Warm per-kernel device times of the full chain (torch.profiler / CUPTI), averaged per frame.

    python tools/kernel_times.py [--frames 8] [--single]

Not a benchmark (profiler attached): use it to see where the frame time goes."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth  # noqa: E402
from blackbox_b200.pipeline import BatchReducer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--tel', default='BG3')
    ap.add_argument('--depth', type=int, default=4)
    ap.add_argument('--ahead', type=int, default=2)
    args = ap.parse_args()
    tel = args.tel
    raw = synth.make_raw(tel, 4001)[0]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    coeffs = synth.make_xtalk(3)[3]
    raws = [R._to_dev(raw) for _ in range(max(2, args.depth))]
    batch = BatchReducer(tel, raw.shape, depth=args.depth, ahead=args.ahead, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=4)
    imgs = [torch.empty(red, dtype=torch.float32, device='cuda') for _ in range(max(2, args.depth))]
    masks = [torch.empty(red, dtype=torch.uint8, device='cuda') for _ in range(max(2, args.depth))]
    frames = [raws[k % 2] for k in range(args.frames)]
    batch.run(frames, imgs, masks)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        batch.run(frames, imgs, masks)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    tot = 0.0
    for ev in prof.events():
        if ev.device_type.name != 'CUDA':
            continue
        name = ev.name.split('(')[0][:56]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time / 1e3 if hasattr(ev, 'device_time') else ev.cuda_time / 1e3
        tot += a[1] * 0
    tot = sum(v[1] for v in agg.values())
    print('%-58s %7s %11s %10s %6s' % ('kernel', 'n/frame', 'ms/frame', 'ms/launch', 'share'))
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('%-58s %7.1f %11.4f %10.4f %5.1f%%' % (k, n / args.frames, ms / args.frames, ms / n, 100 * ms / tot))
    print('sum of device times: %.3f ms/frame over %d frames' % (tot / args.frames, args.frames))


if __name__ == '__main__':
    main()
