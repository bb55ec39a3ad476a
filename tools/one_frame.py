#!/usr/bin/env python
"""Reduce a few full-size synthetic frames one after the other on ONE stream (no overlap between
frames): the program profiled by ncu (launch list and --set full captures under profiles/).

    python tools/one_frame.py [--frames 2] [--tel BG3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth  # noqa: E402
from blackbox_b200.pipeline import FramePipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=2)
    ap.add_argument('--tel', default='BG3')
    ap.add_argument('--niter', type=int, default=4)
    args = ap.parse_args()
    tel = args.tel
    raw = synth.make_raw(tel, 4001)[0]
    red = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, 9, red)
    coeffs = synth.make_xtalk(3)[3]
    raw_t = R._to_dev(raw)
    pipe = FramePipeline(tel, raw.shape, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs, niter=args.niter)
    for k in range(args.frames):
        res = pipe.reduce(raw_t)
    torch.cuda.synchronize()
    print('frames', args.frames, 'NCOSMICS', res.header.get('NCOSMICS'), 'NOBJ-SAT', res.header.get('NOBJ-SAT'),
          'redo', res.redo, 'lac iters', res.header.get('LAC-NIT'))


if __name__ == '__main__':
    main()
