#!/usr/bin/env python
"""Row-stripe sharded master combine on N GPUs over NCCL (BASELINE.json config 5: master bias of
50 binned 5280 x 5280 frames): every rank holds its stripe of each frame, combines
it with the stack-median kernel and one all-gather assembles the master on all ranks.  Checks the
result against the single-GPU combine and prints the device time.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_master.py [--frames 50]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from blackbox_b200 import distributed as D, reduce as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=50)
    ap.add_argument('--size', type=int, default=5280)
    ap.add_argument('--imgtype', default='bias')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dev = torch.device('cuda', torch.cuda.current_device())
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    H = W = args.size
    r0, r1 = D.stripe_bounds(H, rank, world)
    # the same seeded frames on every rank (generated row by row so a stripe equals the rows of the full frame)
    gen = torch.Generator(device=dev)
    stripes, full = [], []
    for k in range(args.frames):
        gen.manual_seed(5000 + k)
        f = torch.randn((H, W), generator=gen, device=dev, dtype=torch.float32) * 8.0 + (1000.0 if args.imgtype == 'flat' else 0.0)
        stripes.append(f[r0:r1].contiguous())
        if rank == 0:
            full.append(f)
        del f
    medsec = [1000.0 + k for k in range(args.frames)] if args.imgtype == 'flat' else None
    for _ in range(2):
        out = D.master_combine_sharded(stripes, (H, W), imgtype=args.imgtype, medsec=medsec)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = D.master_combine_sharded(stripes, (H, W), imgtype=args.imgtype, medsec=medsec)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        want, _ = R.master_combine(full, args.imgtype, medsec=medsec)
        ok = bool(torch.equal(out, want))
        nbytes = (args.frames + 1) * H * W * 4
        print('sharded master: {} frames {}x{} on {} GPU(s): {:.3f} ms (max over ranks), {:.0f} GB/s aggregate, '
              'equal to the single-GPU combine: {}'.format(args.frames, H, W, world, ms.item(), nbytes / ms.item() / 1e6, ok))
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
