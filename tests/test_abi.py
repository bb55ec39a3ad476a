"""The C-ABI library loads without a GPU and exports every symbol include/bbx.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bbx.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bbx_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from blackbox_b200 import build
    return ctypes.CDLL(build.build())


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for must in ('bbx_vos_rowstats', 'bbx_reduce_apply', 'bbx_mask_sat_neighbours', 'bbx_fill_sat_holes',
                 'bbx_xtalk', 'bbx_stack_median', 'bbx_lacosmic', 'bbx_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    from blackbox_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    _lib.load()                        # binds argtypes; raises if the .so is stale


def test_version_and_error_string(lib):
    lib.bbx_version.restype = ctypes.c_int
    lib.bbx_last_error.restype = ctypes.c_char_p
    assert lib.bbx_version() >= 100
    assert isinstance(lib.bbx_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    """Null / inconsistent arguments are rejected before any CUDA call."""
    lib.bbx_last_error.restype = ctypes.c_char_p
    assert lib.bbx_xtalk(None, None, 10, 10, 5, 5, None, None, None) == -1
    assert b'bbx_xtalk' in lib.bbx_last_error()
    assert lib.bbx_stack_median(None, None, 0, ctypes.c_size_t(0), 0, None, 0, None, None) == -1
    assert lib.bbx_medfilt(None, None, 4, 4, 3, None) == -1


def test_product_has_no_cpu_fallback():
    import torch
    from blackbox_b200 import _lib, reduce as bbr
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import numpy as np
    with pytest.raises(_lib.BbxUnavailable):
        bbr.detect_cosmics(np.zeros((8, 8), np.float32), sepmed=False, cleantype='medmask', satlevel=np.inf)
    with pytest.raises(_lib.BbxUnavailable):
        bbr.master_combine([np.zeros((4, 4), np.float32)] * 3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'blackbox_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
