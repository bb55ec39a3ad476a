#!/usr/bin/env python
"""Generate tests/golden/golden.json: digests and spot values of the CPU oracle's outputs on
seeded synthetic inputs.

The reference ships no golden vectors and cannot be imported here (astropy, astroscrappy,
zogy are not installable), so these fixtures pin the ORACLE (regression protection for the
checker itself); the pieces of it that call numpy / scipy are the reference's own
dependencies.  Run from the repository root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from blackbox_b200 import set_bb, synth  # noqa: E402
from oracle import lacosmic, reduce as R  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def spots(a, n=6):
    flat = np.asarray(a).ravel()
    idx = np.linspace(0, flat.size - 1, n).astype(int)
    return [float(flat[i]) for i in idx]


def chain_case(tel, seed, ysc=96, niter=3):
    set_bb.ysize_chan = ysc
    q = ysc // 4
    set_bb.hos_sat_ypix_lim = {'BG2': (2 * q, 4 * q), 'BG3': (q, 2 * q), 'BG4': (q, 2 * q)}
    raw, _ = synth.make_raw(tel, seed, nstars=250, ncosmics=80)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
    coeffs = synth.make_xtalk(seed + 2)[3]
    data, mask, hdr, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, coeffs, niter=niter)
    return {
        'tel': tel, 'seed': seed, 'ysize_chan': ysc, 'niter': niter,
        'raw_sha256': digest(raw), 'mask_sha256': digest(mask),
        'image_rounded_sha256': digest(np.round(data.astype(np.float64), 2).astype(np.float32)),
        'image_spots': spots(data), 'mask_counts': {str(b): int(((mask & b) != 0).sum()) for b in (1, 2, 4, 8, 32, 64)},
        'BIASMEAN': float(hdr['BIASMEAN']), 'RDNOISE': float(hdr['RDNOISE']),
        'NOBJ-SAT': int(hdr['NOBJ-SAT']), 'NCOSMICS': float(hdr['NCOSMICS']),
    }


def lacosmic_case(seed):
    rng = np.random.default_rng(seed)
    img = (300 + 17 * rng.standard_normal((96, 128))).astype(np.float32)
    for _ in range(30):
        y, x = rng.integers(0, 96), rng.integers(0, 120)
        img[y, x:x + rng.integers(1, 6)] += rng.uniform(800, 30000)
    crmask, clean = lacosmic.detect_cosmics(img, sigclip=15, sigfrac=0.01, objlim=3, niter=4, readnoise=8.5,
                                            gain=1.0, satlevel=np.inf, cleantype='medmask', sepmed=False)
    return {'seed': seed, 'input_sha256': digest(img), 'crmask_sha256': digest(crmask.astype(np.uint8)),
            'clean_sha256': digest(clean), 'ncr': int(crmask.sum())}


def extras_case(seed):
    """Master combine (plain and sigma-clipped), nonlin_corr, edge-pixel fill, FITS round trip."""
    from scipy import interpolate
    rng = np.random.default_rng(seed)
    saved = (set_bb.ysize_chan, set_bb.xsize_chan)
    set_bb.ysize_chan, set_bb.xsize_chan = 24, 40
    try:
        shape = (48, 320)
        frames = [(1000 + 10 * rng.standard_normal(shape)).astype(np.float32) for _ in range(20)]
        for k in (0, 7, 13):
            hit = rng.random(shape) < 0.03
            frames[k][hit] += rng.uniform(100, 5000, hit.sum()).astype(np.float32)
        plain, _ = R.master_median(frames, 'bias')
        clipped = R.master_median_clipped(frames, sigma=3.0, maxiters=5)
        splines = []
        for i in range(16):
            x = np.linspace(0, 60000, 80)
            y = 2e-3 * np.sin(x / (7000.0 + 300 * i)) - 3e-7 * x + 2e-4 * rng.standard_normal(x.size)
            splines.append(interpolate.UnivariateSpline(x, y, k=3, s=x.size * 4e-8))
        data = rng.uniform(-500, 140000, size=shape).astype(np.float32)
        nl = R.nonlin_corr(data.copy(), splines, tel='BG3')
        mask = np.zeros(shape, dtype=np.uint8)
        mask[:2] = 32
        mask[:, -3:] = 33
        filled = data.copy()
        meds = R.fill_edge_pixels(filled, mask, tel='BG3')
    finally:
        set_bb.ysize_chan, set_bb.xsize_chan = saved
    return {'seed': seed, 'median20_sha256': digest(plain), 'clipped20_sha256': digest(clipped),
            'clipped_differs_from_plain': int((plain != clipped).sum()),
            'nonlin_sha256': digest(nl), 'nonlin_spots': spots(nl),
            'edge_fill_sha256': digest(filled), 'channel_medians': [float(m) for m in meds]}


def main():
    saved = (set_bb.ysize_chan, dict(set_bb.hos_sat_ypix_lim))
    out = {'numpy': np.__version__,
           'chain': [chain_case('ML1', 1001), chain_case('BG3', 4001), chain_case('BG2', 4002, niter=1)],
           'lacosmic': [lacosmic_case(3), lacosmic_case(4)],
           'extras': [extras_case(7)]}
    set_bb.ysize_chan, set_bb.hos_sat_ypix_lim = saved
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
