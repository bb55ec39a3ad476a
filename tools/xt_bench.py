"""Crosstalk kernels in isolation on a full 10560^2 frame: the tile kernel bbx_xtalk picks (variant 0;
timed with the per-bit mask counts it is asked for in the pipeline) against the TMA-staged
persistent kernel (5, counts taken on the way) and the generic register-only kernels (4 / 2 / 1
pixels per thread)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackbox_b200 import reduce as R, synth  # noqa: E402
from blackbox_b200._lib import call  # noqa: E402

img = torch.randn(10560, 10560, device='cuda') * 50 + 100
mask = (torch.rand(10560, 10560, device='cuda') < 0.02).to(torch.uint8)
coeffs = np.ascontiguousarray(synth.make_xtalk(3)[3], dtype=np.float64)
bits = R._bits('BG3')


def run(variant):
    call('bbx_xtalk_counts', R._ptr(img), R._ptr(mask), 10560, 10560, 5280, 1320,
         coeffs.ctypes.data_as(C.c_void_p), C.byref(bits), variant, R._ptr(counts) if variant in (0, 5) else None, R._stream())


counts = torch.zeros(136, dtype=torch.int64, device='cuda')
for variant in (0, 5, 4, 2, 1):
    for _ in range(3):
        run(variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run(variant)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('xtalk variant', variant, 'ms', ms, 'GB/s', 1003.6e6 / ms / 1e6)
