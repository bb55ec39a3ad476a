#!/usr/bin/env python
"""How much does each RECALLED detail of the two third-party algorithms matter?

astropy's sigma clipping and astroscrappy 1.0.8's detect_cosmics cannot be run in this image, so the
oracle restates them (oracle/csrc/bbo.c) and keeps every detail that was recalled, not read, behind a
named switch (value 0 = what the oracle and the CUDA kernels implement).  This tool runs seeded
frames through every alternative and counts what moves:

  * detect_cosmics on the reduced frame of the reference-made golden case (MeerLICHT, seed 1001) and
    on a BlackGEM frame: pixels whose cosmic-ray flag changes, pixels whose cleaned value changes;
  * os_corr on the raw frames (the sigma-clip switches): BIASMEAN / RDNOISE, pixels of the
    overscan-corrected image that change.

    python tools/oracle_choice_matrix.py [--out profiles/r02_oracle_choice_matrix.txt] [--json ...]

CPU only (the oracle is test infrastructure; nothing of the product runs here)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def frames():
    from blackbox_b200 import set_bb, synth
    out = []
    for tel, seed in (('ML1', 1001), ('BG3', 4001)):
        ysc = 200
        set_bb.ysize_chan = ysc
        q = ysc // 4
        set_bb.hos_sat_ypix_lim = {'BG2': (2 * q, 4 * q), 'BG3': (q, 2 * q), 'BG4': (q, 2 * q)}
        raw, _ = synth.make_raw(tel, seed, nstars=400, ncosmics=150)
        shape = (2 * ysc, 8 * set_bb.xsize_chan)
        mbias, mflat, bpm = synth.make_masters(tel, seed + 1, shape)
        out.append((tel, seed, raw, mbias, mflat, bpm))
        if tel == 'ML1':
            # the same frame without a bad-pixel mask: no 20-pixel 'edge' frame to hide the border rules behind
            out.append((tel + ' (no BPM)', seed, raw, mbias, mflat, np.zeros_like(bpm)))
    return out


def run(tel, raw, mbias, mflat, bpm, niter=4):
    from oracle import reduce as R
    tel = tel.split(' ')[0]
    data, mask, hdr, _ = R.reduce_frame(raw, tel, mbias, mflat, bpm, None, niter=niter,
                                        steps=('gain', 'os', 'bias', 'mask', 'flat', 'cosmics'))
    return data, mask, hdr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    from oracle import clib
    clib.reset_choices()
    fr = frames()
    base = [run(tel, raw, mb, mf, bpm) for tel, _, raw, mb, mf, bpm in fr]
    rows = []
    for name, alts in clib.CHOICES:
        for value, text in alts.items():
            clib.set_choice(name, value)
            try:
                res = []
                for (tel, seed, raw, mb, mf, bpm), (d0, m0, h0) in zip(fr, base):
                    d1, m1, h1 = run(tel, raw, mb, mf, bpm)
                    npix = d0.size
                    cr0, cr1 = (m0 & 2) != 0, (m1 & 2) != 0
                    res.append(dict(frame='{} {}'.format(tel, seed), npix=int(npix), cr_pixels=int(cr0.sum()),
                                    cr_flags_changed=int((cr0 != cr1).sum()),
                                    image_pixels_changed=int((~((d0 == d1) | (np.isnan(d0) & np.isnan(d1)))).sum()),
                                    max_rel_image_change=float(np.nanmax(np.abs(d1 - d0) / np.maximum(np.abs(d0), 1.0))),
                                    d_biasmean=float(h1['BIASMEAN'] - h0['BIASMEAN']),
                                    d_rdnoise=float(h1['RDNOISE'] - h0['RDNOISE']),
                                    d_ncosmics=float(h1['NCOSMICS'] - h0['NCOSMICS'])))
                rows.append(dict(switch=name, value=value, alternative=text, frames=res))
            finally:
                clib.reset_choices()
    # sanity: back at the defaults the results are the baseline's again
    d, m, _ = run(*[fr[0][i] for i in (0, 2, 3, 4, 5)])
    assert np.array_equal(m, base[0][1]) and np.array_equal(d, base[0][0], equal_nan=True)
    lines = ['Recalled choices of the oracle (oracle/csrc/bbo.c, value 0 = implemented) against their alternatives:',
             'what changes on seeded 400 x 10560 frames (full chain up to and including detect_cosmics, niter 4; the third',
             'frame is the first without its bad-pixel mask, whose 20-pixel edge frame otherwise hides every border rule).',
             'Made by tools/oracle_choice_matrix.py on the CPU; "flags" = pixels whose cosmic-ray bit differs, "image" =',
             'pixels of the cleaned image that differ (any amount), of {} pixels per frame.'.format(rows[0]['frames'][0]['npix']), '',
             '%-18s %-2s %-60s %-18s %8s %8s %10s %11s %11s' % ('switch', 'v', 'alternative', 'frame', 'flags', 'image',
                                                                 'max rel', 'd BIASMEAN', 'd RDNOISE')]
    for r in rows:
        for k, f in enumerate(r['frames']):
            lines.append('%-18s %-2s %-60s %-18s %8d %8d %10.2e %11.3e %11.3e' % (
                r['switch'] if k == 0 else '', r['value'] if k == 0 else '', r['alternative'][:60] if k == 0 else '',
                f['frame'], f['cr_flags_changed'], f['image_pixels_changed'], f['max_rel_image_change'],
                f['d_biasmean'], f['d_rdnoise']))
    text = '\n'.join(lines) + '\n'
    print(text)
    if args.out:
        with open(args.out, 'w') as fh:
            fh.write(text)
    if args.json:
        with open(args.json, 'w') as fh:
            json.dump(rows, fh, indent=1)


if __name__ == '__main__':
    main()
