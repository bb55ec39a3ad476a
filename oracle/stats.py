"""CPU ORACLE (test infrastructure only): astropy.stats.sigma_clip / sigma_clipped_stats.

Restated from astropy's published algorithm (>= 4.3 "fast" path: C gufunc that derives the
clipping bounds per slice in float64, Python applies them); astropy itself cannot be installed
here, so this restatement is UNPINNED.  Reference call sites: blackbox.py:6482,
6489, 6500, 6565, 6572, 6652, 6734.

Semantics kept:
  * ``mask_value=v``  -> ``np.ma.masked_values(data, v)``: |x - v| <= 1e-8 + 1e-5|v| masked
  * non-finite values are clipped (astropy also warns)
  * along ``axis`` (None = everything as one slice, C order), at most ``maxiters`` times:
    mean / population std of the survivors in float64 (sequential sums), centre = mean or
    median (even count: mean of the two middle values), keep lo <= x <= hi, stop as soon as
    nothing is rejected
  * the final mask is taken from the FINAL bounds applied to all of the input
  * ``sigma_clipped_stats`` reduces a float64 copy of the data (clipped values = NaN) with
    nan-mean / nan-median / nan-std(ddof): float64 results whatever the input dtype
  * ``sigma_clip(..., masked=True)`` returns a MaskedArray over the ORIGINAL dtype
"""
import numpy as np

from . import clib


def _masked_values(data, value):
    a = np.asarray(data)
    if a.dtype.kind == 'f':
        tol = a.dtype.type(1e-8) + a.dtype.type(1e-5) * a.dtype.type(abs(value))
        return np.abs(a - a.dtype.type(value)) <= tol
    return a == value


def _split(data):
    if isinstance(data, np.ma.MaskedArray):
        return np.asarray(data.data), np.ma.getmaskarray(data).copy()
    a = np.asarray(data)
    return a, np.zeros(a.shape, dtype=bool)


def _to_slices(a, axis):
    """-> (2-D view/copy [nslices, len], function mapping a per-slice vector back to a
    shape broadcastable against ``a``, reduced shape)"""
    if axis is None:
        return a.reshape(1, -1), (lambda v: v.reshape((1,) * a.ndim)), ()
    axis = axis % a.ndim
    moved = np.moveaxis(a, axis, -1)
    red_shape = moved.shape[:-1]
    flat = moved.reshape(-1, moved.shape[-1])
    return flat, (lambda v: np.expand_dims(v.reshape(red_shape), axis)), red_shape


def _bounds(values, masked, axis, sigma_lower, sigma_upper, maxiters, cenfunc):
    flat, expand, red_shape = _to_slices(values, axis)
    mflat, _, _ = _to_slices(masked, axis)
    v64 = np.ascontiguousarray(flat, dtype=np.float64)
    valid = np.ascontiguousarray(~mflat & np.isfinite(v64), dtype=np.uint8)
    lo, hi = clib.clip_bounds(v64, valid, cenfunc == 'median', maxiters, sigma_lower,
                              sigma_upper)
    return v64, valid, lo, hi, expand, red_shape


def sigma_clip(data, sigma=3.0, sigma_lower=None, sigma_upper=None, maxiters=5,
               cenfunc='median', axis=None, masked=True):
    """-> MaskedArray (masked=True) or float64 ndarray with NaN at clipped positions."""
    if cenfunc not in ('mean', 'median'):
        raise ValueError('oracle sigma_clip supports cenfunc mean|median only')
    values, mask = _split(data)
    slo = sigma if sigma_lower is None else sigma_lower
    shi = sigma if sigma_upper is None else sigma_upper
    _, _, lo, hi, expand, _ = _bounds(values, mask, axis, slo, shi, maxiters, cenfunc)
    full = mask | ~np.isfinite(values)
    with np.errstate(invalid='ignore'):
        full |= values < expand(lo)
        full |= values > expand(hi)
    if masked:
        return np.ma.array(values, mask=full, copy=True)
    out = values.astype(np.float64, copy=True)
    out[full] = np.nan
    return out


def sigma_clipped_stats(data, mask=None, mask_value=None, sigma=3.0, sigma_lower=None,
                        sigma_upper=None, maxiters=5, cenfunc='median', std_ddof=0,
                        axis=None):
    """-> (mean, median, std) of the survivors, float64 (NaN for an empty slice)."""
    values, m = _split(data)
    if mask is not None:
        m = m | np.asarray(mask, dtype=bool)
    if mask_value is not None:
        m = m | _masked_values(values, mask_value)
    slo = sigma if sigma_lower is None else sigma_lower
    shi = sigma if sigma_upper is None else sigma_upper
    v64, valid, lo, hi, _, red_shape = _bounds(values, m, axis, slo, shi, maxiters, cenfunc)
    mean, med, std, _ = clib.clipped_moments(v64, valid, lo, hi, ddof=std_ddof)
    if axis is None:
        return np.float64(mean[0]), np.float64(med[0]), np.float64(std[0])
    return mean.reshape(red_shape), med.reshape(red_shape), std.reshape(red_shape)
