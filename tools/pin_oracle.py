#!/usr/bin/env python
"""Pin the oracle against the REAL third-party libraries, the day they can be installed.

The reference delegates three pieces of the hot path to libraries that are absent from this image
(no wheels, no network): astropy.stats (sigma_clip / sigma_clipped_stats, blackbox.py:6482-6734),
astroscrappy.detect_cosmics (blackbox.py:4323-4332; the code comment pins 1.0.8) and CFITSIO's
Rice coder behind fpack / astropy.io.fits (blackbox.py:826-840, 1451).  The oracle restates them;
DESIGN.md calls the result "parity unpinned".  This tool removes the word:

    python tools/pin_oracle.py [--write]       # needs astropy and/or astroscrappy and/or fpack on this machine

For every library it finds it runs the SAME seeded inputs through the library and through the
oracle and reports, per piece, whether they agree bit for bit and, if not, which of the oracle's
named switches (oracle/csrc/bbo.c, `clib.CHOICES`) makes them agree.  With --write it regenerates
tests/golden/reference_golden.json with the real libraries in place of the stubs
(tests/golden/make_reference_golden.py --real).  Exit status: 0 = everything found agrees (or
nothing could be checked), 1 = a disagreement that no switch explains.

Nothing of the product runs here; this is test infrastructure."""
import argparse
import itertools
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def have(mod):
    try:
        __import__(mod)
        return True
    except Exception:
        return False


def seeded_frame(tel='ML1', seed=1001, ysc=200):
    """A reduced frame + mask + read noise to feed detect_cosmics with (the golden case's inputs)."""
    from blackbox_b200 import set_bb, synth
    from oracle import reduce as R
    set_bb.ysize_chan = ysc
    raw, _ = synth.make_raw(tel, seed, nstars=400, ncosmics=150)
    shape = (2 * ysc, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
    data, mask, hdr, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, None, steps=('gain', 'os', 'bias', 'mask', 'flat'))
    return data, mask, float(hdr['RDNOISE']), raw


def search_switches(run_oracle, target_equal, names):
    """Try every single switch value, then pairs, until the oracle reproduces the library."""
    from oracle import clib
    table = dict(clib.CHOICES)
    singles = [(n, v) for n in names for v in table[n]]
    for combo in itertools.chain(((s,) for s in singles), itertools.combinations(singles, 2)):
        if len({n for n, _ in combo}) != len(combo):
            continue
        clib.reset_choices()
        try:
            for n, v in combo:
                clib.set_choice(n, v)
            if target_equal(run_oracle()):
                return combo
        finally:
            clib.reset_choices()
    return None


def check_lacosmic(report):
    if not have('astroscrappy'):
        report.append(('astroscrappy.detect_cosmics', 'not installed', None))
        return True
    import astroscrappy
    from oracle import lacosmic as L
    data, mask, rn, _ = seeded_frame()
    kw = dict(sigclip=15.0, sigfrac=0.01, objlim=3.0, gain=1.0, readnoise=rn, satlevel=np.inf, niter=3,
              sepmed=False, cleantype='medmask')
    cr_l, clean_l = astroscrappy.detect_cosmics(data.copy(), inmask=(mask != 0), **kw)
    clean_l = np.asarray(clean_l, dtype=np.float32)

    def run():
        return L.detect_cosmics(data.copy(), (mask != 0), **kw)

    def equal(res):
        return np.array_equal(res[0], cr_l) and np.array_equal(res[1], clean_l, equal_nan=True)

    res = run()
    if equal(res):
        report.append(('astroscrappy.detect_cosmics ' + getattr(astroscrappy, '__version__', '?'),
                       'oracle agrees bit for bit (mask and cleaned image)', None))
        return True
    nd = int((res[0] != cr_l).sum()), int((~((res[1] == clean_l) | (np.isnan(res[1]) & np.isnan(clean_l)))).sum())
    names = ['REBIN_ORDER', 'LAPLACE_ORDER', 'LAPLACE_EDGE', 'CLEAN_MEDIAN', 'BACKGROUND_MEDIAN', 'SIGCLIP_CMP',
             'OBJLIM_CMP', 'M5_FLOOR', 'MEDFILT_FRAME', 'DILATE3_FRAME', 'CLEAN_FRAME', 'FINE_MEDIAN7_OF']
    combo = search_switches(run, equal, names)
    report.append(('astroscrappy.detect_cosmics ' + getattr(astroscrappy, '__version__', '?'),
                   'DIFFERS: {} mask pixels, {} image pixels'.format(*nd),
                   'agrees with switches {}'.format(combo) if combo else 'no single switch or pair explains it'))
    return combo is not None


def check_sigma_clip(report):
    if not have('astropy'):
        report.append(('astropy.stats.sigma_clipped_stats', 'not installed', None))
        return True
    import astropy
    from astropy.stats import sigma_clip, sigma_clipped_stats
    from oracle import stats as S
    rng = np.random.default_rng(11)
    strip = rng.normal(6450.0, 8.5, (530, 174)).astype(np.float32)
    strip[rng.random(strip.shape) < 0.01] += 400.0
    strip[rng.random(strip.shape) < 0.002] = 0.0
    cases = [dict(axis=1, mask_value=0, cenfunc='mean'), dict(mask_value=0, cenfunc='mean'),
             dict(sigma=5, cenfunc='mean'), dict(cenfunc='median')]

    def run():
        out = [S.sigma_clipped_stats(strip, **kw) for kw in cases]
        ma = S.sigma_clip(np.ma.masked_array(strip[:10], strip[:10] > 6470), axis=0, cenfunc='mean', sigma=2.5)
        return out, np.ma.getmaskarray(ma)

    want = [sigma_clipped_stats(strip, **kw) for kw in cases]
    wmask = np.ma.getmaskarray(sigma_clip(np.ma.masked_array(strip[:10], strip[:10] > 6470), axis=0, cenfunc='mean',
                                          sigma=2.5))

    def equal(res):
        out, m = res
        ok = np.array_equal(m, wmask)
        for a, b in zip(out, want):
            for x, y in zip(a, b):
                ok = ok and np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True)
        return ok

    if equal(run()):
        report.append(('astropy.stats ' + astropy.__version__, 'oracle agrees bit for bit (mean, median, std, masks)', None))
        return True
    combo = search_switches(run, equal, ['CLIP_INCLUSIVE', 'CLIP_STD_ABOUT', 'CLIP_STOP'])
    report.append(('astropy.stats ' + astropy.__version__, 'DIFFERS',
                   'agrees with switches {}'.format(combo) if combo else
                   'no switch explains it (check the dtype of the statistics: float64 copy vs input dtype, DESIGN.md section 2)'))
    return combo is not None


def check_rice(report):
    fpack, funpack = shutil.which('fpack'), shutil.which('funpack')
    if not fpack:
        report.append(('CFITSIO Rice coder (fpack)', 'fpack not on PATH', None))
        return True
    from blackbox_b200 import fitsio
    from oracle import rice
    _, _, _, raw = seeded_frame()
    rows = raw[:64]
    ok = True
    with tempfile.TemporaryDirectory() as d:
        plain = os.path.join(d, 'raw.fits')
        fitsio.write_primary(plain, rows)
        subprocess.check_call([fpack, '-D', '-Y', plain])
        ci = fitsio.read_compressed(plain + '.fz')
        heap = np.asarray(ci.heap)
        stored = (rows.astype(np.int32) - 32768).astype(np.int16)
        same_bytes = all(rice.encode_tile(stored[r], 2) == heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes()
                         for r in range(rows.shape[0]))
        same_pixels = all(np.array_equal(rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(),
                                                          rows.shape[1], 2).astype(np.uint16) ^ 0x8000, rows[r])
                          for r in range(rows.shape[0]))
        report.append(('CFITSIO Rice coder (fpack, 16 bit)', 'decoder reads fpack\'s tiles: {}; encoder writes fpack\'s bytes: {}'.format(
            same_pixels, same_bytes), None))
        ok = ok and same_pixels and same_bytes
        mask = np.zeros((64, 1056), np.uint8)
        mask[::7, ::13] = 4
        mplain = os.path.join(d, 'mask.fits')
        fitsio.write_primary(mplain, mask)
        subprocess.check_call([fpack, '-D', '-Y', mplain])
        ci = fitsio.read_compressed(mplain + '.fz')
        heap = np.asarray(ci.heap)
        same8 = all(rice.encode_tile(mask[r], 1) == heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes()
                    for r in range(mask.shape[0]))
        report.append(('CFITSIO Rice coder (fpack, 8 bit)', 'encoder writes fpack\'s bytes: {}'.format(same8), None))
        ok = ok and same8
        # fpack -q 16 of a float image (blackbox.py:836): ZSCALE / ZZERO per row do not depend on the
        # dither seed (fpack draws it from the clock), so they pin fits_quantize_float + FnNoise5_float
        # exactly; the integers are compared with the oracle's for the ZDITHER0 fpack chose
        rng = np.random.default_rng(77)
        img = (rng.normal(300, 12, (48, 1056)) + np.linspace(0, 40, 1056)[None, :]).astype(np.float32)
        img[5] = np.round(img[5])                      # ties: the differences' skip rules
        img[6, 100:400] = 17.0                         # a constant run
        fplain = os.path.join(d, 'red.fits')
        fitsio.write_primary(fplain, img)
        subprocess.check_call([fpack, '-q', '16', '-D', '-Y', fplain])
        ci = fitsio.read_compressed(fplain + '.fz')
        heap = np.asarray(ci.heap)
        same_scale = same_ints = True
        for r in range(img.shape[0]):
            q, sc, ze = rice.fpack_quantize_row(img[r], r, 16.0, ci.info['zdither0'])
            same_scale = same_scale and q is not None and sc == ci.zscale[r] and ze == ci.zzero[r]
            if q is not None and r not in ci.fallback:
                got = rice.decode_tile(heap[ci.offsets[r]:ci.offsets[r] + ci.lengths[r]].tobytes(), img.shape[1], 4)
                same_ints = same_ints and np.array_equal(np.asarray(got).astype(np.int64).astype(np.int32), q)
        report.append(('CFITSIO quantiser (fpack -q 16)', 'ZSCALE / ZZERO of every row equal: {}; quantised integers equal: {}'.format(
            same_scale, same_ints), None if same_scale else 'oracle/rice.py: fn_noise5_row / fpack_quantize_row (float32 against double arithmetic of '
                                                            'the differences, the d2 array length, the zero-point rule)'))
        ok = ok and same_scale and same_ints
        if funpack:
            # our writer's file through funpack
            tiles = [rice.encode_tile(mask[r], 1) for r in range(mask.shape[0])]
            ours = os.path.join(d, 'ours_mask.fits.fz')
            fitsio.write_compressed(ours, np.frombuffer(b''.join(tiles), np.uint8), [len(t) for t in tiles], mask.shape, 8)
            subprocess.check_call([funpack, ours])
            _, back, info = fitsio.read_primary(ours[:-3])
            good = np.array_equal(np.asarray(back), mask)
            report.append(('fitsio.write_compressed -> funpack', 'round trip: {}'.format(good), None))
            ok = ok and good
            # and a float image as the oracle's writer lays it out (bbx_fpack_f32 writes the same table)
            fours = os.path.join(d, 'ours_red.fits.fz')
            _, back_want = rice.write_fz_f32(fours, img, zdither0=4321)
            subprocess.check_call([funpack, fours])
            _, fback, finfo = fitsio.read_primary(fours[:-3])
            fgood = np.array_equal(fitsio.to_native(fback, finfo), back_want)
            report.append(('float .fits.fz (ZSCALE / ZZERO / ZDITHER0) -> funpack', 'values as un-quantised here: {}'.format(fgood), None))
            ok = ok and fgood
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--write', action='store_true', help='regenerate tests/golden/reference_golden.json with the real libraries')
    args = ap.parse_args()
    report = []
    ok = check_sigma_clip(report)
    ok = check_lacosmic(report) and ok
    ok = check_rice(report) and ok
    for name, what, extra in report:
        print('{:<44} {}'.format(name, what))
        if extra:
            print('{:<44} -> {}'.format('', extra))
    if all(w in ('not installed', 'fpack not on PATH') for _, w, _ in report):
        print('nothing to pin against on this machine: the oracle stays "parity unpinned" for these pieces')
    if args.write:
        if not (have('astropy') and have('astroscrappy')):
            print('--write needs both astropy and astroscrappy')
            return 2
        gen = os.path.join(ROOT, 'tests', 'golden', 'make_reference_golden.py')
        return subprocess.call([sys.executable, gen, '--real'])
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
