#!/usr/bin/env python
"""Reduce a directory of raw (uncompressed) FITS frames on one GPU: raw FITS bytes -> pinned host
buffers -> BatchReducer.run_host(fits=True) -> <name>_red.fits + <name>_mask.fits.

    python tools/reduce_night.py RAW_DIR OUT_DIR --tel BG3 [--mbias F --mflat F --bpm F --xtalk F]

The file handling of the reference (header checks, QC, fpack, calibration-frame selection;
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
    _, data, info = fitsio.read_primary(path)
    return np.array(fitsio.to_native(data, info) if dtype != np.uint8 else np.asarray(data), dtype=dtype, order='C', copy=True)


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
    ap.add_argument('--fill-edge', action='store_true')
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    files = sorted(glob.glob(os.path.join(args.raw_dir, '*.fits')))
    files = [files[k] for k in shard_frames(len(files), rank, world)]
    if not files:
        return 0
    os.makedirs(args.out_dir, exist_ok=True)
    headers, raws = [], []
    for path in files:
        hdr, buf, info = fitsio.read_primary(path, pinned=True)
        if info['bitpix'] != 16 or info['bzero'] != 32768.0:
            raise fitsio.FitsError('{}: expected a raw 16-bit frame with BZERO 32768'.format(path))
        headers.append(hdr)
        raws.append(buf.view(torch.uint16).view(info['shape']))
    coeffs = R.read_crosstalk_file(args.xtalk) if args.xtalk else None
    batch = BatchReducer(args.tel, tuple(raws[0].shape), depth=args.depth, use_graphs=True, fill_edge=args.fill_edge,
                         mbias=_master(args.mbias, np.float32), mflat=_master(args.mflat, np.float32),
                         bpm=_master(args.bpm, np.uint8), coeffs=coeffs, niter=args.niter)
    RH, RW = batch.pipes[0].geom.red_shape
    imgs = [torch.empty((RH, RW), dtype=torch.float32).pin_memory() for _ in files]
    masks = [torch.empty((RH, RW), dtype=torch.uint8).pin_memory() for _ in files]
    exptimes = [float(h['EXPTIME'][0]) if 'EXPTIME' in h else 60.0 for h in headers]     # NCOSMICS is a rate
    results = batch.run_host(raws, imgs, masks, fill_header=True, fits=True, exptimes=exptimes)
    for path, hdr, res, img, mask in zip(files, headers, results, imgs, masks):
        base = os.path.splitext(os.path.basename(path))[0]
        out_hdr = {k: v for k, v in hdr.items() if k not in ('COMMENT', 'HISTORY')}
        out_hdr.update({k: (v, '') for k, v in res.header.items()})
        out_hdr['REDFILE'] = (base + '_red', 'BlackBOX reduced image name')
        out_hdr['MASKFILE'] = (base + '_mask', 'BlackBOX mask image name')
        fitsio.write_primary(os.path.join(args.out_dir, base + '_red.fits'), img.view(torch.uint8).reshape(-1), out_hdr,
                             be_bytes=True, shape=(RH, RW), bitpix=-32)
        fitsio.write_primary(os.path.join(args.out_dir, base + '_mask.fits'), mask.numpy(),
                             {k: (v, '') for k, v in res.header_mask.items()})
    return len(files)


if __name__ == '__main__':
    print('reduced', main(), 'frame(s)')
