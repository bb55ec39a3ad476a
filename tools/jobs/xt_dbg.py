import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from blackbox_b200 import reduce as R, synth
from blackbox_b200._lib import call
H, W = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (128, 1056)
use_mask = int(sys.argv[3]) if len(sys.argv) > 3 else 1
img = torch.randn(H, W, device='cuda') * 50 + 100
mask = (torch.rand(H, W, device='cuda') < 0.02).to(torch.uint8) if use_mask else None
coeffs = np.ascontiguousarray(synth.make_xtalk(3)[3], dtype=np.float64)
bits = R._bits('BG3')
counts = torch.zeros(136, dtype=torch.int64, device='cuda') if use_mask == 1 else None
call('bbx_xtalk_counts', R._ptr(img), R._ptr(mask), H, W, H // 2, W // 8, coeffs.ctypes.data_as(C.c_void_p), C.byref(bits), 0, R._ptr(counts), R._stream())
torch.cuda.synchronize()
print('ok', H, W, use_mask, counts)
