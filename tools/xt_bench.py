import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from blackbox_b200 import reduce as R, synth
R.tel = 'BG3'
img = torch.randn(10560, 10560, device='cuda') * 50 + 100
mask = (torch.rand(10560, 10560, device='cuda') < 0.02).to(torch.uint8)
coeffs = synth.make_xtalk(3)[3]
for px in ('4', '2', '1'):
    os.environ['BBX_XTALK_PX'] = px
    for _ in range(3): R.xtalk_enqueue(img, mask, coeffs, 'BG3')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): R.xtalk_enqueue(img, mask, coeffs, 'BG3')
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('xtalk px', px, 'ms', ms, 'GB/s', 1003.6e6 / ms / 1e6)
