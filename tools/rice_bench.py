#!/usr/bin/env python
"""Time bbx_rice_decode16 on a full-size raw frame (10600 tiles of 12000 pixels).  Encoding a whole
frame with the pure-Python test coder takes minutes, so 64 distinct rows of a synthetic BlackGEM
raw frame are encoded once and the 10600 tile descriptors cycle through them; the decoder does the
full work (every tile is decoded and written) and the result is checked.

    python tools/rice_bench.py [--reps 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    from blackbox_b200 import reduce as bbr, synth
    from oracle import rice                      # the encoder: test infrastructure, used to make the input
    rows = synth.make_raw('BG3', 4001, ysize_chan=12)[0]                      # (64, 12000) uint16
    stored = (rows.astype(np.int32) - 32768).astype(np.int16)
    tiles = [rice.encode_tile16(r) for r in stored]
    lens64 = np.array([len(t) for t in tiles], dtype=np.int32)
    offs64 = np.concatenate(([0], np.cumsum(lens64)[:-1])).astype(np.int64)
    heap = torch.from_numpy(np.frombuffer(b''.join(tiles), dtype=np.uint8).copy()).cuda()
    H, W = 10600, 12000
    idx = np.arange(H) % 64
    offs = torch.from_numpy(offs64[idx]).cuda()
    lens = torch.from_numpy(lens64[idx]).cuda()
    info = dict(bitpix=16, shape=(H, W), bzero=32768.0, bscale=1.0, blocksize=32, bytepix=2)
    out = torch.empty((H, W), dtype=torch.uint16, device='cuda')
    bbr.rice_decode(heap, offs, lens, info, out=out)
    ok = bool((out.view(torch.int16).cpu().numpy().view(np.uint16) == rows[idx]).all())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    times = []
    for _ in range(args.reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        bbr.rice_decode(heap, offs, lens, info, out=out, check=False)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    med = float(np.median(times))
    comp = float(lens64[idx].sum())
    print('rice_decode16 full frame {}x{}: correct {}, median {:.3f} ms, min {:.3f} ms; compressed {:.1f} MB '
          '({:.2f} of the raw 254.4 MB), output {:.1f} GB/s'.format(H, W, ok, med, min(times), comp / 1e6,
                                                                     comp / (H * W * 2), H * W * 2 / med / 1e6))
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
