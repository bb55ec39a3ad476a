#!/usr/bin/env python
"""Timeline of a BatchReducer run (torch.profiler / CUPTI chrome trace): how much of the wall time
the GPU is busy, how much of it with kernels of two frames overlapping, and where the gaps are.

    python tools/timeline.py [--frames 8] [--depth 2] [--out gpurun_out/trace.json]

Not a benchmark (profiler attached): a development aid for the stream / host pipelining."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth  # noqa: E402
from blackbox_b200.pipeline import BatchReducer  # noqa: E402


def analyse(path, frames):
    with open(path) as fh:
        tr = json.load(fh)
    ev = [e for e in tr['traceEvents'] if e.get('ph') == 'X']
    kern = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
    rt = [e for e in ev if e.get('cat') == 'cuda_runtime']
    if not kern:
        print('no device events in the trace')
        return
    t0 = min(e['ts'] for e in kern)
    t1 = max(e['ts'] + e['dur'] for e in kern)
    pts = []
    for e in kern:
        pts.append((e['ts'], 1))
        pts.append((e['ts'] + e['dur'], -1))
    pts.sort()
    busy = over = 0.0
    depth, last = 0, pts[0][0]
    gaps = []
    for t, d in pts:
        if depth >= 1:
            busy += t - last
        elif t - last > 5.0:
            gaps.append((last - t0, t - last))
        if depth >= 2:
            over += t - last
        depth += d
        last = t
    streams = sorted(set(e['args'].get('stream') for e in kern))
    kp = sorted([(e['ts'], 1) for e in kern if e.get('cat') == 'kernel'] +
                [(e['ts'] + e['dur'], -1) for e in kern if e.get('cat') == 'kernel'])
    kbusy, depth_k, last_k = 0.0, 0, kp[0][0]
    for t, d in kp:
        if depth_k >= 1:
            kbusy += t - last_k
        depth_k += d
        last_k = t
    print('kernels alone (copies left out): busy %.3f ms (%.1f%% of the wall time)' % (kbusy / 1e3, 100 * kbusy / (t1 - t0)))
    for cat, what in (('gpu_memcpy', 'copies'),):
        cp = [e for e in kern if e.get('cat') == cat]
        for kind in sorted(set(e['name'] for e in cp)):
            sel = [e for e in cp if e['name'] == kind]
            print('   %-28s %4d, %.3f ms per frame' % (kind, len(sel), sum(e['dur'] for e in sel) / 1e3 / frames))
    print('frames %d  wall %.3f ms (%.3f ms/frame)  busy %.3f ms (%.1f%%)  >=2 kernels resident %.3f ms (%.1f%%)' % (
        frames, (t1 - t0) / 1e3, (t1 - t0) / 1e3 / frames, busy / 1e3, 100 * busy / (t1 - t0), over / 1e3,
        100 * over / (t1 - t0)))
    print('sum of device durations %.3f ms/frame; streams %s' % (sum(e['dur'] for e in kern) / 1e3 / frames, streams))
    print('idle gaps > 5 us: %d, total %.3f ms; largest:' % (len(gaps), sum(g[1] for g in gaps) / 1e3))
    for at, g in sorted(gaps, key=lambda x: -x[1])[:8]:
        print('   at %.3f ms: %.1f us' % (at / 1e3, g))
    # per kernel: device time per frame, and the part of it during which nothing else was resident
    # ("alone": what the kernel really costs the frame rate)
    bounds = sorted(set([e['ts'] for e in kern] + [e['ts'] + e['dur'] for e in kern]))
    import bisect
    cover = [0] * (len(bounds) - 1)
    for e in kern:
        a, b = bisect.bisect_left(bounds, e['ts']), bisect.bisect_left(bounds, e['ts'] + e['dur'])
        for i in range(a, b):
            cover[i] += 1
    tot, alone = {}, {}
    for e in kern:
        name = e['name'].split('(')[0].split('<')[0][-44:]
        a, b = bisect.bisect_left(bounds, e['ts']), bisect.bisect_left(bounds, e['ts'] + e['dur'])
        tot[name] = tot.get(name, 0.0) + e['dur']
        alone[name] = alone.get(name, 0.0) + sum(bounds[i + 1] - bounds[i] for i in range(a, b) if cover[i] == 1)
    print('%-46s %10s %10s   (us per frame)' % ('kernel', 'device', 'alone'))
    for name in sorted(tot, key=lambda n: -alone[n])[:34]:
        print('%-46s %10.1f %10.1f' % (name, tot[name] / frames, alone[name] / frames))
    print('%-46s %10.1f %10.1f' % ('TOTAL', sum(tot.values()) / frames, sum(alone.values()) / frames))
    if rt:
        h0 = min(e['ts'] for e in rt)
        h1 = max(e['ts'] + e['dur'] for e in rt)
        sync = [e for e in rt if 'Synchronize' in e['name']]
        launch = [e for e in rt if 'Launch' in e['name']]
        print('host: runtime calls span %.3f ms, %d launches (%.3f ms inside launch calls), %d syncs (%.3f ms waiting)' % (
            (h1 - h0) / 1e3, len(launch), sum(e['dur'] for e in launch) / 1e3, len(sync), sum(e['dur'] for e in sync) / 1e3))


def main_host(args):
    import numpy as np
    from blackbox_b200 import fitsio
    tel = args.tel
    raws = [R._to_dev(synth.make_raw(tel, 4001 + k)[0]) for k in range(4)]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    batch = BatchReducer(tel, tuple(raws[0].shape), depth=args.depth, ahead=args.ahead, mbias=mbias, mflat=mflat, bpm=bpm,
                         coeffs=synth.make_xtalk(3)[3], niter=4, use_graphs=bool(args.graphs))
    ring = []
    for r in raws:
        heap, lens = R.rice_encode(r)
        offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
        info = dict(shape=tuple(r.shape), bitpix=16, bytepix=2, bzero=32768.0, bscale=1.0, blocksize=32)
        ring.append(fitsio.CompressedImage({}, torch.from_numpy(heap).pin_memory(), offs, lens.astype(np.int32), info))
        ring[-1].descriptors()
    host_img = [torch.empty(batch.img_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(args.depth)]
    host_mask = [torch.empty(batch.mask_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(args.depth)]
    host_raw = [ring[k % len(ring)] for k in range(args.frames)]
    for _ in range(2):
        batch.run_host(host_raw, host_img, host_mask, mask_fz=True, img_fz=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        batch.run_host(host_raw, host_img, host_mask, mask_fz=True, img_fz=True)
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    prof.export_chrome_trace(args.out)
    analyse(args.out, args.frames)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--tel', default='BG3')
    ap.add_argument('--depth', type=int, default=4)
    ap.add_argument('--ahead', type=int, default=2)
    ap.add_argument('--graphs', type=int, default=0)
    ap.add_argument('--split-priority', type=int, default=1)
    ap.add_argument('--out', default='gpurun_out/trace.json')
    ap.add_argument('--host', type=int, default=0,
                    help='1: run_host with fpacked raw frames in, fpack -q 16 image + Rice-coded mask out')
    args = ap.parse_args()
    if args.host:
        return main_host(args)
    tel = args.tel
    raw = synth.make_raw(tel, 4001)[0]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    coeffs = synth.make_xtalk(3)[3]
    raws = [R._to_dev(raw) for _ in range(2)]
    batch = BatchReducer(tel, raw.shape, depth=args.depth, ahead=args.ahead, split_priority=bool(args.split_priority), use_graphs=bool(args.graphs), mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=4)
    nbuf = args.frames
    imgs = [torch.empty(red, dtype=torch.float32, device='cuda') for _ in range(nbuf)]
    masks = [torch.empty(red, dtype=torch.uint8, device='cuda') for _ in range(nbuf)]
    frames = [raws[k % 2] for k in range(args.frames)]
    batch.run(frames, imgs, masks)
    batch.run(frames, imgs, masks)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        batch.run(frames, imgs, masks)
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    prof.export_chrome_trace(args.out)
    analyse(args.out, args.frames)


if __name__ == '__main__':
    main()
