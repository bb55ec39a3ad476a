"""CPU ORACLE (test infrastructure only): astroscrappy.detect_cosmics, version 1.0.8 behaviour.

astroscrappy is a third-party dependency of the reference (pyproject.toml:25, unpinned; the
reference's code comment ties the call to 1.0.8, blackbox.py:4319) and is not installable
here: the algorithm is restated from the LACosmic paper (van Dokkum 2001, PASP 113, 1420) and
astroscrappy's published implementation -- PARITY UNPINNED.  Only the path the reference
exercises is provided: fsmode='median', cleantype='medmask', sepmed=False, pssl=0
(call site blackbox.py:4323-4332).

Two implementations that must agree bit for bit (tests/test_oracle.py):
  * ``detect_cosmics``        -> C (oracle/csrc/bbo.c), fast enough for full frames
  * ``detect_cosmics_numpy``  -> numpy/scipy.ndimage, literal and slow, small frames only
"""
import numpy as np
from scipy import ndimage

from . import clib

F32 = np.float32


def _check_supported(sepmed, cleantype, fsmode, pssl):
    if sepmed or cleantype != 'medmask' or fsmode != 'median' or pssl != 0.0:
        raise NotImplementedError(
            'oracle detect_cosmics covers sepmed=False, cleantype="medmask", '
            'fsmode="median", pssl=0 (the path blackbox.py:4323 uses)')


def detect_cosmics(indat, inmask=None, sigclip=4.5, sigfrac=0.3, objlim=5.0, gain=1.0,
                   readnoise=6.5, satlevel=65536.0, pssl=0.0, niter=4, sepmed=True,
                   cleantype='meanmask', fsmode='median', psfmodel='gauss', psffwhm=2.5,
                   psfsize=7, psfk=None, psfbeta=4.765, verbose=False, info=None):
    """-> (crmask bool, cleanarr float32); astroscrappy 1.0.8 signature."""
    _check_supported(sepmed, cleantype, fsmode, pssl)
    clean = np.array(indat, dtype=F32, order='C', copy=True)
    clean *= F32(gain)
    mask = (np.zeros(clean.shape, np.uint8) if inmask is None
            else np.ascontiguousarray(inmask, dtype=np.uint8).copy())
    crmask, nit, ncr, bg, _ = clib.detect_cosmics_c(
        clean, mask, F32(sigclip), F32(sigfrac), F32(objlim), F32(readnoise), F32(satlevel),
        int(niter))
    clean /= F32(gain)
    if info is not None:
        info.update(iterations=nit, ncr_per_iter=ncr, background=bg)
    return crmask.astype(bool), clean


# -------------------------------------------------------------------------------------------
# literal numpy/scipy twin (slow)
# -------------------------------------------------------------------------------------------
def _medfilt(a, k):
    r = k // 2
    out = a.copy()
    if a.shape[0] >= k and a.shape[1] >= k:
        med = ndimage.median_filter(a, size=k, mode='nearest')
        out[r:-r, r:-r] = med[r:-r, r:-r]
    return out


def _dilate3(b):
    out = b.copy()
    d = ndimage.binary_dilation(b, structure=np.ones((3, 3), bool))
    out[1:-1, 1:-1] = d[1:-1, 1:-1]
    return out


def laplace_plus(clean):
    """L+ : subsample x2, Laplacian (zero padded), clip negatives, 2x2 block mean; float32
    with one rounding per step in the order of oracle/csrc/bbo.c."""
    H, W = clean.shape
    sub = np.repeat(np.repeat(clean, 2, axis=0), 2, axis=1)
    p = F32(4.0) * sub
    p[:, :-1] -= sub[:, 1:]
    p[:, 1:] -= sub[:, :-1]
    p[:-1, :] -= sub[1:, :]
    p[1:, :] -= sub[:-1, :]
    p[p < 0] = 0
    s = p[0::2, 0::2].copy()
    s += p[0::2, 1::2]
    s += p[1::2, 0::2]
    s += p[1::2, 1::2]
    return s / F32(4.0)


def detect_cosmics_numpy(indat, inmask=None, sigclip=4.5, sigfrac=0.3, objlim=5.0,
                         readnoise=6.5, niter=4, info=None):
    """satlevel=inf, gain=1 path only."""
    clean = np.array(indat, dtype=F32, copy=True)
    mask = np.zeros(clean.shape, bool) if inmask is None else np.asarray(inmask, bool)
    H, W = clean.shape
    good = clean[~mask]
    background = (np.sort(good)[(good.size - 1) // 2] if good.size else F32(0))
    crmask = np.zeros(clean.shape, bool)
    sigclip, objlim, readnoise = F32(sigclip), F32(objlim), F32(readnoise)
    sigcliplow = F32(sigfrac) * sigclip
    rn2 = readnoise * readnoise
    dumps = {}
    for it in range(niter):
        s = laplace_plus(clean)
        m5 = _medfilt(clean, 5)
        m5[m5 < F32(0.00001)] = F32(0.00001)
        noise = np.sqrt(m5 + rn2)
        s = s / (F32(2.0) * noise)
        sp = s - _medfilt(s, 5)
        f = _medfilt(clean, 3)
        f = (f - _medfilt(f, 7)) / noise
        f[f < F32(0.01)] = F32(0.01)
        if it == 0:
            dumps.update(sp=sp.copy(), f=f.copy(), noise=noise.copy())
        goodpix = ~mask
        cr = (sp > sigclip) & goodpix & ((sp / f) > objlim)
        cr = _dilate3(cr) & goodpix & (sp > sigclip)
        cr = _dilate3(cr) & goodpix & (sp > sigcliplow)
        ncr = int(cr.sum())
        crmask |= cr
        if ncr == 0:
            break
        bad = crmask | mask
        ys, xs = np.nonzero(crmask[2:H - 2, 2:W - 2])
        newvals = np.empty(len(ys), F32)
        for n, (y, x) in enumerate(zip(ys + 2, xs + 2)):
            win = clean[y - 2:y + 3, x - 2:x + 3][~bad[y - 2:y + 3, x - 2:x + 3]]
            newvals[n] = np.sort(win)[(win.size - 1) // 2] if win.size else background
        clean[ys + 2, xs + 2] = newvals
    if info is not None:
        info.update(background=background, **dumps)
    return crmask, clean
