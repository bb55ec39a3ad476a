#!/usr/bin/env python
"""Development aid: phase timing of vos_std_kernel (needs a build with BBX_NVCC_EXTRA=-DVSTD_PROFILE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from blackbox_b200 import reduce as R, set_bb, synth, _lib  # noqa: E402
from blackbox_b200.geometry import Geometry  # noqa: E402

tel = 'BG3'
raw = synth.make_raw(tel, 4001)[0]
raw_t = R._to_dev(raw)
geom = Geometry.from_raw_shape(raw.shape, tel=tel)
gain = [float(x) for x in set_bb.get_par(set_bb.gain, tel)]
lib = _lib.load()
for rep in range(3):
    st = R.overscan_enqueue(raw_t, geom, tel, gain=gain)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)()
    n = C.c_int(0)
    lib.bbx_debug_vstd_clocks(buf, C.byref(n))
    t = [buf[i] for i in range(n.value)]
    print('ticks', n.value, 'deltas [cycles]', [t[i + 1] - t[i] for i in range(len(t) - 1)], 'total', t[-1] - t[0] if t else None)
print('std', st.std_vos.tolist()[:3])
