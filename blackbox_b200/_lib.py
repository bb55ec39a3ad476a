"""ctypes binding of libbbx.so (the C ABI declared in include/bbx.h).

There is no CPU fallback: if the library is missing or no CUDA device is present the import
of the compute entry points fails loudly (``BbxUnavailable``).
"""
import ctypes as C
import os

from .geometry import BbxGeom

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'libbbx.so')


class BbxUnavailable(RuntimeError):
    pass


class BbxError(RuntimeError):
    pass


class BbxMaskBits(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('bad', 'cosmic', 'saturated', 'satcon', 'sattrail',
                                       'edge', 'crosstalk')]

    @classmethod
    def from_dict(cls, mv):
        return cls(mv['bad'], mv['cosmic ray'], mv['saturated'], mv['saturated-connected'],
                   mv['satellite trail'], mv['edge'], mv['crosstalk'])


P, I, D, F, SZ = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_size_t
GEOM = C.POINTER(BbxGeom)
BITS = C.POINTER(BbxMaskBits)

# name -> argtypes; restype is int for all but the few listed in _RESTYPES
_SIGNATURES = {
    'bbx_version': [],
    'bbx_vos_rowstats': [P, I, GEOM, P, D, I, P, P],
    'bbx_vos_fit': [P, GEOM, I, D, P, P, P, P, P],
    'bbx_hos_satcount': [P, I, GEOM, P, P, P, I, I, P, P],
    'bbx_hos_stats': [P, I, GEOM, P, P, I, F, P, P, P, P, P, P, P],
    'bbx_vos_std': [P, I, GEOM, P, P, P, P, P],
    'bbx_hos_fit': [P, P, P, P, GEOM, I, I, I, P, P, P, P],
    'bbx_reduce_apply': [P, I, GEOM, P, P, P, P, P, P, P, BITS, P, P, P, P, C.c_uint, P],
    'bbx_reduce_apply_scan': [P, I, GEOM, P, P, P, P, P, P, P, BITS, P, P, P, P, C.c_uint, P, F, F, F, F, P, I, P, P, P],
    'bbx_reduce_apply_stats': [P, I, GEOM, P, P, P, P, P, P, P, BITS, P, P, P, P, C.c_uint, I, P, P, P],
    'bbx_satlevels': [P, P, P, P],
    'bbx_header_means': [P, P, P, P],
    'bbx_mask_sat_neighbours': [P, I, I, I, I, BITS, P],
    'bbx_mask_morph_sparse': [P, I, I, I, I, BITS, P, P, C.c_uint, P, P, P, I, P, P],
    'bbx_mask_morph_sparse_track': [P, I, I, I, I, BITS, P, P, C.c_uint, P, P, P, I, P, P, P, I, P],
    'bbx_fill_holes_work_bytes': [I, I],
    'bbx_fill_sat_holes': [P, I, I, BITS, P, I, P, P],
    'bbx_fill_holes_more': [P, I, I, BITS, P, I, P, P],
    'bbx_count_objects': [P, I, I, I, P, P, P],
    'bbx_mask_counts': [P, SZ, P, P],
    'bbx_xtalk': [P, P, I, I, I, I, P, BITS, P],
    'bbx_xtalk_variant': [P, P, I, I, I, I, P, BITS, I, P],
    'bbx_xtalk_counts': [P, P, I, I, I, I, P, BITS, I, P, P],
    'bbx_stack_median': [P, P, I, SZ, I, P, I, P, P],
    'bbx_stack_median_multi': [P, P, I, SZ, I, P, I, P, I, I, P],
    'bbx_stack_clipped_median': [P, P, I, SZ, D, I, I, P, I, P, P],
    'bbx_lacosmic_work_bytes': [I, I],
    'bbx_lacosmic': [P, P, P, I, I, F, F, F, F, P, I, I, P, P, P],
    'bbx_lacosmic_begin': [P, P, P, I, I, I, I, P, P, P],
    'bbx_lacosmic_iteration': [P, P, P, I, I, F, F, F, F, P, I, I, P, P, P],
    'bbx_lacosmic_finish': [P, P, I, I, I, I, P, P, P, P],
    'bbx_select_work_bytes': [],
    'bbx_masked_lower_median': [P, P, SZ, P, P, P],
    'bbx_medfilt': [P, P, I, I, I, P],
    'bbx_laplace_plus': [P, P, I, I, P],
    'bbx_gain_corr': [P, GEOM, P, P],
    'bbx_binary_inplace': [P, P, SZ, I, P],
    'bbx_mask_or': [P, P, SZ, I, P],
    'bbx_nonlin_corr': [P, I, I, I, I, P, P, P, P, P, I, F, P],
    'bbx_fits_decode': [P, I, I, SZ, P, P],
    'bbx_fits_encode': [P, I, I, SZ, P, P],
    'bbx_rice_decode16': [P, SZ, P, P, I, I, I, I, P, P, P],
    'bbx_rice_decode': [P, SZ, P, P, I, I, I, I, I, P, P, P],
    'bbx_unquantize': [P, I, I, P, P, P, I, I, I, I, P, P],
    'bbx_rice_encode_work_bytes': [I, I, I],
    'bbx_rice_encode_out_bytes': [I, I, I],
    'bbx_rice_encode': [P, I, I, I, P, SZ, P, SZ, P],
    'bbx_fpack_f32_work_bytes': [I, I],
    'bbx_fpack_f32_out_bytes': [I, I],
    'bbx_fpack_f32_heap_offset': [I],
    'bbx_fpack_f32': [P, I, I, F, I, P, P, SZ, P, SZ, P],
    'bbx_chanmed_work_bytes': [],
    'bbx_channel_medians': [P, I, I, I, I, I, P, P, P],
    'bbx_fill_edge': [P, P, I, I, I, I, I, P, P],
}
_RESTYPES = {'bbx_fill_holes_work_bytes': SZ, 'bbx_lacosmic_work_bytes': SZ,
             'bbx_select_work_bytes': SZ, 'bbx_chanmed_work_bytes': SZ,
             'bbx_rice_encode_work_bytes': SZ, 'bbx_rice_encode_out_bytes': SZ,
             'bbx_fpack_f32_work_bytes': SZ, 'bbx_fpack_f32_out_bytes': SZ, 'bbx_fpack_f32_heap_offset': SZ}

EXPORTS = tuple(sorted(list(_SIGNATURES) + ['bbx_last_error']))

_lib = None


def load():
    """Load libbbx.so (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BbxUnavailable(
            'libbbx.so not found at {}: build it with `python -m blackbox_b200.build` '
            '(there is no CPU fallback)'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    lib.bbx_last_error.restype = C.c_char_p
    lib.bbx_last_error.argtypes = []
    missing = []
    for name, argtypes in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if missing:
        raise BbxUnavailable('libbbx.so at {} is stale, it lacks {}: rebuild with '
                             '`python -m blackbox_b200.build --force`'.format(LIB_PATH, missing))
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point; raise BbxError with bbx_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise BbxError('{} failed ({}): {}'.format(
            name, rc, lib.bbx_last_error().decode('utf-8', 'replace')))


def query(name, *args):
    """Call a size-returning entry point."""
    return getattr(load(), name)(*args)
