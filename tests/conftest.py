import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture
def small_bb():
    """Shrink the channel height (columns keep their full 1320 so the column constants of
    os_corr stay meaningful) and restore the settings afterwards."""
    from blackbox_b200 import set_bb
    saved = (set_bb.ysize_chan, set_bb.xsize_chan, dict(set_bb.hos_sat_ypix_lim))

    def apply(ysize_chan=200, xsize_chan=1320, lim=None):
        set_bb.ysize_chan = ysize_chan
        set_bb.xsize_chan = xsize_chan
        q = ysize_chan // 4
        set_bb.hos_sat_ypix_lim = lim or {'BG2': (2 * q, 4 * q), 'BG3': (q, 2 * q),
                                          'BG4': (q, 2 * q)}
        return set_bb

    yield apply
    set_bb.ysize_chan, set_bb.xsize_chan, set_bb.hos_sat_ypix_lim = saved


def float_class_ok(a, b, scale, rtol=1e-5):
    """The float tolerance class of the parity contract: |a-b| <= rtol*max(|a|,|b|) +
    rtol*scale, where scale is the level that was subtracted before (cancellation)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * scale
