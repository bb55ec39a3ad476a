"""Host-side smoothing spline of os_corr (reference: blackbox.py:6683-6723, 6795).

The reference fits ``scipy.interpolate.UnivariateSpline`` (FITPACK curfit) to the first 180
column means of the horizontal overscan and uses it for the first 150 columns -- but then
overwrites every column whose clipped mean is valid (and, for BlackGEM, not affected by
saturation) with that plain mean (blackbox.py:6797-6814).  The spline therefore matters only
for the few columns the GPU fit kernel flags in ``need_spline``; FITPACK stays on the host
(scipy is a dependency of the reference as well) and is evaluated

  * always, in strict mode (``os_corr`` default): reproduces the reference's error behaviour
    (a FITPACK UserWarning triggers one retry with k=3, s=1.5*npoints; a second one raises);
  * only for channels that need it, in the batched pipeline.
"""
import warnings

import numpy as np
from scipy import interpolate

IDX_SWITCH = 150
OVERLAP = 30


def running_median3(y):
    """3-point running median over y[3:], windows clipped to [3, n), computed from the
    un-smoothed values; the 2-point windows at both ends give the mean of the two
    (blackbox.py:6703-6708).  Vectorised: the interior is the median of three shifted views."""
    y = np.array(y, copy=True)
    n = len(y)
    if n <= 3:
        return y
    src = y.copy()
    if n >= 6:
        y[4:n - 1] = np.median(np.stack([src[3:n - 2], src[4:n - 1], src[5:n]]), axis=0)
    for k in {3, n - 1} | ({4} if n < 6 else set()):
        if 3 <= k < n:
            y[k] = np.median(src[max(k - 1, 3):min(k + 2, n)])
    return y


def hos_spline(mean_hos, std_hos, nvalues):
    """Spline values for columns 0..149 of one channel (float64 [150]).

    mean_hos, std_hos: float32 [ncols]; nvalues: int [ncols] (column statistics of the
    horizontal overscan after clipping)."""
    ncols = len(mean_hos)
    valid = nvalues > 1
    xcol = np.arange(ncols) + 1
    err = np.zeros(ncols, dtype=np.float32)
    err[valid] = std_hos[valid] / np.sqrt(nvalues[valid])
    weights = np.zeros(ncols, dtype=np.float32)
    nz = err != 0
    weights[nz] = 1 / err[nz]
    if np.all(valid[0:3]):
        weights[0:3] = 0
    idx = np.arange(min(IDX_SWITCH + OVERLAP, ncols))
    npoints = int(np.sum(valid[idx] & nz[idx]))
    sel = valid[idx]
    y2fit = running_median3(mean_hos[idx][sel])
    xs, ws = xcol[idx][sel], weights[idx][sel]
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        try:
            spl = interpolate.UnivariateSpline(xs, y2fit, w=ws, k=2, s=npoints)
        except UserWarning:
            spl = interpolate.UnivariateSpline(xs, y2fit, w=ws, k=3, s=1.5 * npoints)
    return spl(xcol[0:IDX_SWITCH])
