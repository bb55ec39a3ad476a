"""Host logic of the multi-GPU paths on CPU: gloo, world_size 2 (spawned processes)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blackbox_b200 import distributed as D
from blackbox_b200.pipeline import shard_frames


def test_shard_frames_partition():
    for n, w in [(64, 8), (64, 3), (5, 8), (0, 2)]:
        parts = [shard_frames(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_frames(10, 1, 4) == [1, 5, 9]


def test_stripe_bounds_cover_all_rows():
    for H, w in [(10560, 8), (5280, 8), (10, 4), (7, 2), (3, 8)]:
        rows = []
        for r in range(w):
            a, b = D.stripe_bounds(H, r, w)
            assert 0 <= a <= b <= H
            rows += list(range(a, b))
        assert rows == list(range(H))
    assert D.stripe_bounds(10560, 3, 8) == (3960, 5280)


def _numpy_combine(stripes, scales, flat_fix, bpm_stripe, tel):
    cube = np.stack([s.numpy() for s in stripes]).astype(np.float32)
    if flat_fix:
        for i, sc in enumerate(scales):
            if sc != 0:
                cube[i] /= np.float32(sc)
    out = np.median(cube, axis=0)
    if flat_fix and bpm_stripe is not None:
        out[(bpm_stripe.numpy() == 32) | (out <= 0)] = 1
    return out


def _worker(rank, world, port, H, W, nframes, imgtype, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        frames = [rng.normal(1000 * (1 + 0.1 * i), 30, (H, W)).astype(np.float32) for i in range(nframes)]
        bpm = np.zeros((H, W), np.uint8)
        bpm[0, :] = 32
        medsec = [float(np.median(f)) for f in frames] if imgtype == 'flat' else None
        r0, r1 = D.stripe_bounds(H, rank, world)
        stripes = [torch.from_numpy(f[r0:r1].copy()) for f in frames]
        full = D.master_combine_sharded(stripes, (H, W), imgtype=imgtype, medsec=medsec,
                                        bpm_stripe=torch.from_numpy(bpm[r0:r1].copy()),
                                        combine=_numpy_combine)
        np.save(os.path.join(result_dir, 'rank{}.npy'.format(rank)), full.numpy())
        if rank == 0:
            cube = np.stack(frames)
            if imgtype == 'flat':
                cube = np.stack([f / np.float32(m) for f, m in zip(frames, medsec)])
            want = np.median(cube, axis=0)
            if imgtype == 'flat':
                want[(bpm == 32) | (want <= 0)] = 1
            np.save(os.path.join(result_dir, 'want.npy'), want)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('H,imgtype', [(16, 'bias'), (13, 'flat')])
def test_master_combine_sharded_gloo(H, imgtype, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), H, 24, 5, imgtype, str(tmp_path)), nprocs=world, join=True)
    want = np.load(tmp_path / 'want.npy')
    for r in range(world):
        got = np.load(tmp_path / 'rank{}.npy'.format(r))
        assert got.shape == want.shape
        assert np.array_equal(got, want)            # N-rank result == 1-rank result, bitwise


def test_master_combine_sharded_single_process():
    rng = np.random.default_rng(1)
    frames = [torch.from_numpy(rng.normal(0, 1, (6, 5)).astype(np.float32)) for _ in range(4)]
    out = D.master_combine_sharded(frames, (6, 5), combine=_numpy_combine)
    assert np.array_equal(out.numpy(), np.median(np.stack([f.numpy() for f in frames]), axis=0))
    with pytest.raises(ValueError):
        D.master_combine_sharded(frames, (6, 5), imgtype='flat', combine=_numpy_combine)
    with pytest.raises(ValueError):
        D.master_combine_sharded(frames, (7, 5), combine=_numpy_combine)
