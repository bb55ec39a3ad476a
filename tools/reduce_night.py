#!/usr/bin/env python
"""Reduce a directory of raw FITS frames on one GPU per rank, files in -> files out:

    python tools/reduce_night.py RAW_DIR OUT_DIR --tel BG3 [--mbias F --mflat F --bpm F --xtalk F] [--fpack]

RAW_DIR holds either fpacked raw frames (`*.fits.fz`, as the telescope delivers them: the
Rice-coded heap goes to the GPU and is unpacked there) or uncompressed ones (`*.fits`).  Products:
`<name>_red.fits` + `<name>_mask.fits`, or with --fpack what the reference leaves on disk
(blackbox.py:826-836, 7677-7679): `<name>_red.fits.fz` (`fpack -q 16 -D -Y`, quantised and
Rice-coded on the GPU) + `<name>_mask.fits.fz` (`fpack -D -Y`).  Masters and the bad-pixel mask
may be packed or not.

The file handling of the reference (header checks, QC, calibration-frame selection;
blackbox.py:1100-1460, 1987-2030) is not reproduced: this is the data path only, with the header
keywords the reduction steps set.  Under torchrun every rank takes frames rank, rank + world, ...
"""
import argparse
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from blackbox_b200 import fitsio, reduce as R  # noqa: E402
from blackbox_b200.pipeline import BatchReducer, shard_frames  # noqa: E402


def _master(path, dtype):
    if not path:
        return None
    return R.read_fits_image(path, dtype)[1]


def _base(path):
    name = os.path.basename(path)
    for ext in ('.fits.fz', '.fits'):
        if name.endswith(ext):
            return name[:-len(ext)]
    return os.path.splitext(name)[0]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('raw_dir')
    ap.add_argument('out_dir')
    ap.add_argument('--tel', default='BG3')
    ap.add_argument('--mbias')
    ap.add_argument('--mflat')
    ap.add_argument('--bpm')
    ap.add_argument('--xtalk', help='crosstalk coefficient table (victim source correction)')
    ap.add_argument('--niter', type=int, default=None)
    ap.add_argument('--depth', type=int, default=4)
    ap.add_argument('--chunk', type=int, default=16, help='frames per pass (host ring buffers)')
    ap.add_argument('--fill-edge', action='store_true')
    ap.add_argument('--fpack', action='store_true', help='write _red.fits.fz (fpack -q 16) and _mask.fits.fz')
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    packed = sorted(glob.glob(os.path.join(args.raw_dir, '*.fits.fz')))
    plain = sorted(glob.glob(os.path.join(args.raw_dir, '*.fits')))
    if packed and plain:
        raise SystemExit('{}: fpacked and plain raw frames side by side; one kind per directory'.format(args.raw_dir))
    files = packed or plain
    files = [files[k] for k in shard_frames(len(files), rank, world)]
    if not files:
        return 0
    os.makedirs(args.out_dir, exist_ok=True)
    coeffs = R.read_crosstalk_file(args.xtalk) if args.xtalk else None
    batch, done = None, 0
    for start in range(0, len(files), max(args.chunk, 1)):
        group = files[start:start + max(args.chunk, 1)]
        headers, raws = [], []
        for path in group:
            if packed:
                ci = fitsio.read_compressed(path, pinned=True)
                hdr, shape = ci.header, tuple(ci.info['shape'])
                raws.append(ci)
            else:
                hdr, buf, info = fitsio.read_primary(path, pinned=True)
                if info['bitpix'] != 16 or info['bzero'] != 32768.0:
                    raise fitsio.FitsError('{}: expected a raw 16-bit frame with BZERO 32768'.format(path))
                shape = tuple(info['shape'])
                raws.append(buf.view(torch.uint16).view(shape))
            headers.append(hdr)
        if batch is None:
            batch = BatchReducer(args.tel, shape, depth=args.depth, use_graphs=True, fill_edge=args.fill_edge,
                                 mbias=_master(args.mbias, torch.float32), mflat=_master(args.mflat, torch.float32),
                                 bpm=_master(args.bpm, torch.uint8), coeffs=coeffs, niter=args.niter)
            RH, RW = batch.pipes[0].geom.red_shape
            n = min(max(args.chunk, 1), len(files))
            if args.fpack:
                imgs = [torch.empty(batch.img_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(n)]
                masks = [torch.empty(batch.mask_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(n)]
            else:
                imgs = [torch.empty((RH, RW), dtype=torch.float32).pin_memory() for _ in range(n)]
                masks = [torch.empty((RH, RW), dtype=torch.uint8).pin_memory() for _ in range(n)]
        exptimes = [float(h['EXPTIME'][0]) if 'EXPTIME' in h else 60.0 for h in headers]     # NCOSMICS is a rate
        results = batch.run_host(raws, imgs, masks, fill_header=True, fits=not packed,
                                 exptimes=exptimes, img_fz=args.fpack, mask_fz=args.fpack,
                                 zdither0=[1 + (start + k) % 10000 for k in range(len(group))] if args.fpack else None)
        for k, (path, hdr, res) in enumerate(zip(group, headers, results)):
            base = _base(path)
            out_hdr = {key: v for key, v in hdr.items() if key not in ('COMMENT', 'HISTORY')}
            out_hdr.update({key: (v, '') for key, v in res.header.items()})
            out_hdr['REDFILE'] = (base + '_red', 'BlackBOX reduced image name')
            out_hdr['MASKFILE'] = (base + '_mask', 'BlackBOX mask image name')
            hdr_mask = {key: (v, '') for key, v in res.header_mask.items()}
            if args.fpack:
                fitsio.write_compressed(os.path.join(args.out_dir, base + '_red.fits.fz'), shape=(RH, RW), zbitpix=-32,
                                        header=out_hdr, **res.img_fz)
                heap, lens = res.mask_fz
                fitsio.write_compressed(os.path.join(args.out_dir, base + '_mask.fits.fz'), heap, lens, (RH, RW), 8, hdr_mask)
            elif packed:
                fitsio.write_primary(os.path.join(args.out_dir, base + '_red.fits'), imgs[k].numpy(), out_hdr)
                fitsio.write_primary(os.path.join(args.out_dir, base + '_mask.fits'), masks[k].numpy(), hdr_mask)
            else:
                fitsio.write_primary(os.path.join(args.out_dir, base + '_red.fits'), imgs[k].view(torch.uint8).reshape(-1),
                                     out_hdr, be_bytes=True, shape=(RH, RW), bitpix=-32)
                fitsio.write_primary(os.path.join(args.out_dir, base + '_mask.fits'), masks[k].numpy(), hdr_mask)
        done += len(group)
    return done


if __name__ == '__main__':
    print('reduced', main(), 'frame(s)')
