#!/usr/bin/env python
"""Time the Rice codec kernels at BASELINE's full size, L2 flushed between repetitions:
  * bbx_rice_decode, BYTEPIX 2: a whole fpacked raw frame (10600 tiles of 12000 pixels), the frame
    coded by bbx_rice_encode itself (the tests hold its bytes against the oracle's);
  * bbx_rice_encode, BYTEPIX 1: the 10560^2 mask of a reduced frame (three kernels: code, scan, compact);
  * bbx_rice_encode, BYTEPIX 2: the raw frame;
  * bbx_fpack_f32: the reduced float32 image as `fpack -q 16` (row noise, quantisation, Rice code).

    python tools/rice_bench.py [--reps 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timeit(fn, reps, flush):
    times = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), float(min(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    from blackbox_b200 import reduce as bbr, set_bb, synth
    from blackbox_b200.pipeline import FramePipeline
    raw = synth.make_raw('BG3', 4001)[0]
    raw_t = bbr._to_dev(raw)
    H, W = raw.shape
    heap, lens = bbr.rice_encode(raw_t)
    offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
    heap_t, offs_t, lens_t = torch.from_numpy(heap).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(lens).cuda()
    info = dict(bitpix=16, shape=(H, W), bzero=32768.0, bscale=1.0, blocksize=32, bytepix=2)
    out = torch.empty((H, W), dtype=torch.uint16, device='cuda')
    bbr.rice_decode(heap_t, offs_t, lens_t, info, out=out)
    ok = bool(torch.equal(out, raw_t))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    med, mn = timeit(lambda: bbr.rice_decode(heap_t, offs_t, lens_t, info, out=out, check=False), args.reps, flush)
    print('rice_decode BYTEPIX 2, raw frame {}x{}: correct {}, median {:.3f} ms, min {:.3f} ms; compressed {:.1f} MB '
          '({:.2f} of the raw {:.1f} MB), output {:.1f} GB/s'.format(H, W, ok, med, mn, heap.size / 1e6,
                                                                     heap.size / (H * W * 2), H * W * 2 / 1e6,
                                                                     H * W * 2 / med / 1e6))
    shape = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters('BG3', 9, shape)
    pipe = FramePipeline('BG3', raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=synth.make_xtalk(3)[3], niter=4)
    res = pipe.reduce(raw_t)
    mask = res.mask
    enc = bbr.RiceEncoder(shape, 1, mask.device, out_bytes=(8 << 20) + 16 + 4 * shape[0] + 64)
    host = enc.enqueue(mask).cpu()
    total, mlens, mheap, fits = enc.parse(host)
    moffs = np.concatenate(([0], np.cumsum(mlens.astype(np.int64))[:-1]))
    minfo = dict(bitpix=8, shape=shape, bzero=0.0, bscale=1.0, blocksize=32, bytepix=1)
    back = bbr.rice_decode(torch.from_numpy(mheap.copy()).cuda(), moffs, mlens.copy(), minfo)
    ok = fits and bool(torch.equal(back, mask))
    med, mn = timeit(lambda: enc.enqueue(mask), args.reps, flush)
    print('rice_encode BYTEPIX 1, mask {}x{}: round trip {}, median {:.3f} ms, min {:.3f} ms; {:.2f} MB of heap '
          '(1/{:.0f} of the mask), input {:.1f} GB/s'.format(shape[0], shape[1], ok, med, mn, total / 1e6,
                                                           shape[0] * shape[1] / max(total, 1),
                                                           shape[0] * shape[1] / med / 1e6))
    enc2 = bbr.RiceEncoder((H, W), 2, raw_t.device)
    stored = raw_t.view(torch.int16) ^ torch.tensor(-32768, dtype=torch.int16, device='cuda')
    med, mn = timeit(lambda: enc2.enqueue(stored), args.reps, flush)
    print('rice_encode BYTEPIX 2, raw frame: median {:.3f} ms, min {:.3f} ms, input {:.1f} GB/s'.format(
        med, mn, H * W * 2 / med / 1e6))
    del enc2, stored
    # bbx_fpack_f32: the reduced image as `fpack -q 16` (row statistics + quantising Rice coder + scan + compact)
    img = res.img
    fenc = bbr.FpackEncoder(shape, img.device, 16.0)
    got = fenc.parse(fenc.enqueue(img, 4242).cpu().numpy())
    pinfo = dict(bitpix=-32, shape=shape, blocksize=32, bytepix=4, quantize='SUBTRACTIVE_DITHER_1', zdither0=4242, zblank=None)
    poffs = np.concatenate(([0], np.cumsum(got['lengths'].astype(np.int64))[:-1]))
    back = bbr.rice_decode(torch.from_numpy(got['heap'].copy()).cuda(), poffs, got['lengths'].copy(), pinfo,
                           zscale=got['zscale'].copy(), zzero=got['zzero'].copy())
    err = (back - img).abs() / torch.from_numpy(got['zscale'].copy()).cuda().float()[:, None]
    ok2 = bool(got['fits']) and got['skipped'] == 0 and float(err.max()) <= 0.5 + 1e-2
    print('fpack_f32: fits {}, rows not quantised {}, largest error {:.6f} steps'.format(got['fits'], got['skipped'], float(err.max())))
    med, mn = timeit(lambda: fenc.enqueue(img, 4242), args.reps, flush)
    print('fpack_f32 -q 16, image {}x{}: round trip within half a step {}, median {:.3f} ms, min {:.3f} ms; '
          '{:.1f} MB of heap ({:.3f} bytes per pixel, 1/{:.1f} of the float32 image), median ZSCALE {:.4f}'.format(
              shape[0], shape[1], ok2, med, mn, got['total'] / 1e6, got['total'] / (shape[0] * shape[1]),
              4.0 * shape[0] * shape[1] / max(got['total'], 1), float(np.median(got['zscale']))))
    return 0 if (ok and ok2) else 1


if __name__ == '__main__':
    sys.exit(main())
