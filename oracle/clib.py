"""ctypes bindings of oracle/csrc/bbo.c (CPU ORACLE -- test infrastructure only)."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.bbo_lower_median_f.restype = C.c_float
        _lib.bbo_detect_cosmics.restype = C.c_int
        _lib.bbo_rice_encode.restype = C.c_long
        _lib.bbo_rice_decode.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def clip_bounds(v, valid, use_median, maxiters, sig_lo, sig_hi):
    """v, valid: [nslices, len] float64 / uint8 C-contiguous -> (lo, hi) float64 [nslices]."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    ns, n = v.shape
    lo = np.empty(ns)
    hi = np.empty(ns)
    lib().bbo_clip_bounds(_p(v), _p(valid), C.c_long(ns), C.c_long(n), C.c_int(int(use_median)),
                          C.c_int(-1 if (maxiters is None or np.isinf(maxiters)) else int(maxiters)),
                          C.c_double(sig_lo), C.c_double(sig_hi), _p(lo), _p(hi))
    return lo, hi


def clipped_moments(v, valid, lo, hi, ddof=0, want_median=True):
    v = np.ascontiguousarray(v, dtype=np.float64)
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    ns, n = v.shape
    mean = np.empty(ns)
    med = np.empty(ns) if want_median else None
    std = np.empty(ns)
    cnt = np.empty(ns, dtype=np.int64)
    lib().bbo_clipped_moments(_p(v), _p(valid), C.c_long(ns), C.c_long(n), _p(lo), _p(hi),
                              C.c_int(ddof), _p(mean), _p(med), _p(std), _p(cnt))
    return mean, med, std, cnt


def _img_op(name, a, out_dtype=np.float32):
    a = np.ascontiguousarray(a)
    out = np.empty(a.shape, dtype=out_dtype)
    getattr(lib(), name)(_p(a), _p(out), C.c_int(a.shape[0]), C.c_int(a.shape[1]))
    return out


def medfilt3(a):
    return _img_op('bbo_medfilt3', np.asarray(a, np.float32))


def medfilt5(a):
    return _img_op('bbo_medfilt5', np.asarray(a, np.float32))


def medfilt7(a):
    return _img_op('bbo_medfilt7', np.asarray(a, np.float32))


def laplace(a):
    return _img_op('bbo_laplace', np.asarray(a, np.float32))


def subsample(a):
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty((2 * a.shape[0], 2 * a.shape[1]), np.float32)
    lib().bbo_subsample(_p(a), _p(out), C.c_int(a.shape[0]), C.c_int(a.shape[1]))
    return out


def rebin(a):
    a = np.ascontiguousarray(a, np.float32)
    H, W = a.shape[0] // 2, a.shape[1] // 2
    out = np.empty((H, W), np.float32)
    lib().bbo_rebin(_p(a), _p(out), C.c_int(H), C.c_int(W))
    return out


def dilate3(a):
    return _img_op('bbo_dilate3', np.asarray(a, np.uint8), np.uint8)


def dilate5(a, niter):
    a = np.ascontiguousarray(a, np.uint8)
    out = np.empty_like(a)
    lib().bbo_dilate5(_p(a), _p(out), C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(niter))
    return out


def lower_median(a):
    a = np.ascontiguousarray(a, np.float32).ravel()
    return np.float32(lib().bbo_lower_median_f(_p(a), C.c_long(a.size)))


def clean_medmask(clean, crmask, mask, background):
    assert clean.dtype == np.float32 and clean.flags.c_contiguous
    crmask = np.ascontiguousarray(crmask, np.uint8)
    mask = np.ascontiguousarray(mask, np.uint8)
    lib().bbo_clean_medmask(_p(clean), _p(crmask), _p(mask), C.c_int(clean.shape[0]),
                            C.c_int(clean.shape[1]), C.c_float(background))


def detect_cosmics_c(clean, mask, sigclip, sigfrac, objlim, readnoise, satlevel, niter,
                     dump=False):
    """In place on ``clean`` (float32) and ``mask`` (uint8); returns
    (crmask uint8, iterations, ncr_per_iter, background, dumps-or-None)."""
    assert clean.dtype == np.float32 and clean.flags.c_contiguous
    assert mask.dtype == np.uint8 and mask.flags.c_contiguous
    H, W = clean.shape
    crmask = np.zeros((H, W), np.uint8)
    ncr = np.zeros(max(niter, 1), np.int64)
    bg = C.c_float(0)
    dumps = [np.empty((H, W), np.float32) for _ in range(3)] if dump else [None] * 3
    nit = lib().bbo_detect_cosmics(
        _p(clean), _p(mask), _p(crmask), C.c_int(H), C.c_int(W),
        C.c_float(sigclip), C.c_float(sigfrac), C.c_float(objlim), C.c_float(readnoise),
        C.c_float(satlevel), C.c_int(niter), _p(ncr),
        _p(dumps[0]), _p(dumps[1]), _p(dumps[2]), C.byref(bg))
    return crmask, nit, ncr[:nit], np.float32(bg.value), (dumps if dump else None)


def stack_median(frames, scale=None):
    """frames: list of float32 arrays of one shape; returns their per-pixel median."""
    frames = [np.ascontiguousarray(f, np.float32) for f in frames]
    n = len(frames)
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
    sc = None if scale is None else np.ascontiguousarray(scale, np.float32)
    out = np.empty(frames[0].shape, np.float32)
    lib().bbo_stack_median(ptrs, _p(sc), C.c_int(n), C.c_long(out.size), _p(out))
    return out


def rice_encode(a, bytepix):
    """One tile (stored int8 / int16 / int32 pixels) -> coded bytes (bbo_rice_encode)."""
    dt = {1: np.int8, 2: np.int16, 4: np.int32}[bytepix]
    a = np.ascontiguousarray(np.asarray(a).astype(dt))
    cap = a.size * bytepix + a.size // 8 + 64 + bytepix * 40
    out = np.empty(cap, dtype=np.uint8)
    n = lib().bbo_rice_encode(_p(a), C.c_long(a.size), C.c_int(bytepix), _p(out), C.c_long(cap))
    if n < 0:
        raise ValueError('rice_encode: output buffer too small')
    return out[:n].tobytes()


def rice_decode(buf, nx, bytepix):
    """Coded bytes -> nx stored pixels as uint8 / uint16 / uint32 (bbo_rice_decode)."""
    c = np.frombuffer(bytes(buf), dtype=np.uint8)
    out = np.empty(nx, dtype={1: np.uint8, 2: np.uint16, 4: np.uint32}[bytepix])
    if lib().bbo_rice_decode(_p(c), C.c_long(c.size), C.c_long(nx), C.c_int(bytepix), _p(out)) != 0:
        raise ValueError('rice_decode: the stream ends before the tile does')
    return out


# names of the recalled-choice switches of oracle/csrc/bbo.c (same order as its enum) and their alternatives
CHOICES = [
    ('REBIN_ORDER', {1: '2x2 mean summed column-major ((a+c)+b)+d', 2: '2x2 mean summed pairwise (a+b)+(c+d)'}),
    ('LAPLACE_ORDER', {1: '4c -left -right -up -down', 2: '4c - ((l+r)+(u+d))'}),
    ('LAPLACE_EDGE', {1: 'image edge replicated instead of zero padding'}),
    ('CLEAN_MEDIAN', {1: 'medmask: upper median for even counts', 2: 'medmask: mean of the two middle values'}),
    ('BACKGROUND_MEDIAN', {1: 'background level: upper median'}),
    ('SIGCLIP_CMP', {1: "s' >= sigclip / sigcliplow instead of >"}),
    ('OBJLIM_CMP', {1: "s'/f >= objlim instead of >"}),
    ('M5_FLOOR', {1: 'med5 <= 0 -> 1e-5 instead of med5 < 1e-5 -> 1e-5', 2: 'no floor on med5'}),
    ('MEDFILT_FRAME', {1: 'median-filter frame zero instead of a copy of the input'}),
    ('DILATE3_FRAME', {1: '3x3 dilation zero-padded at the frame instead of copying it'}),
    ('CLEAN_FRAME', {1: 'cleaning also in the 2-pixel frame'}),
    ('FINE_MEDIAN7_OF', {1: 'fine structure med3 - med7(image) instead of med3 - med7(med3)'}),
    ('CLIP_INCLUSIVE', {1: 'sigma clip keeps lo < x < hi instead of lo <= x <= hi'}),
    ('CLIP_STD_ABOUT', {1: 'sigma clip std about the centre (median) instead of the mean'}),
    ('CLIP_STOP', {1: 'sigma clip: one more bound computation after maxiters'}),
]


def set_choice(name, value):
    """Flip one recalled choice of the C oracle (0 = the default this repo implements)."""
    names = [n for n, _ in CHOICES]
    assert lib().bbo_num_choices() == len(names)
    lib().bbo_set_choice(C.c_int(names.index(name)), C.c_int(int(value)))


def reset_choices():
    lib().bbo_reset_choices()
