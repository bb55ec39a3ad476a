"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed``.

* Science frames are independent (the reference already runs one OS process per frame,
  blackbox.py:378): frame k goes to rank k mod world_size, nothing is exchanged
  (``pipeline.shard_frames``).
* The master-frame combine (master_prep core, blackbox.py:4908-4984) is a per-pixel median, so
  it shards by ROW STRIPES of the stacked frames: rank g holds rows
  [g*ceil(H/G), (g+1)*ceil(H/G)) of every one of the N frames, combines its stripe with the
  stack-median kernel, and a single all-gather (NCCL on GPUs) assembles the master on every
  rank.  That all-gather is the only collective of the whole build.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import set_bb
from .set_bb import get_par


def stripe_rows(H, world_size):
    """Rows per stripe (the last stripe may be shorter or empty)."""
    return (H + world_size - 1) // world_size


def stripe_bounds(H, rank, world_size):
    """[r0, r1) of the rows rank ``rank`` owns."""
    n = stripe_rows(H, world_size)
    r0 = min(rank * n, H)
    return r0, min(r0 + n, H)


class PeerMaster:
    """The gather buffer of the sharded master combine as SYMMETRIC memory: every rank's buffer is
    mapped into every other rank's address space over NVLink / NVSwitch (``torch.distributed.
    _symmetric_memory``), so the stack-median kernel stores each master pixel straight into all of
    them (``bbx_stack_median_multi``) -- compute and all-gather in one kernel, the transfer under
    the loads, no NCCL call on the data path.  ``multicast``: store once to the NVSwitch multicast
    address instead and let the switch replicate (needs multicast support on the fabric)."""

    def __init__(self, shape, group=None, device=None, multicast=False):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        H, W = shape
        self.shape = (H, W)
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nrows = stripe_rows(H, self.world)
        dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.buf = symm.empty((self.world * self.nrows, W), dtype=torch.float32, device=dev)
        self.hdl = symm.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.multicast = bool(multicast) and bool(getattr(self.hdl, 'has_multicast_support', False)) \
            and int(getattr(self.hdl, 'multicast_ptr', 0) or 0) != 0
        self.mc_ptr = int(self.hdl.multicast_ptr) if self.multicast else 0

    def stripe_ptrs(self):
        """Addresses of THIS rank's slot in every rank's buffer (own buffer first), or the one
        multicast address of that slot."""
        off = self.rank * self.nrows * self.shape[1] * 4
        if self.multicast:
            return [self.mc_ptr + off]
        order = [self.rank] + [r for r in range(self.world) if r != self.rank]
        return [self.ptrs[r] + off for r in order]

    def barrier(self):
        """All ranks have reached this point of their current streams (device-side, no host sync)."""
        self.hdl.barrier()


def _default_combine(stripes, scales, flat_fix, bpm_stripe, tel, out=None):
    from . import reduce as R
    res, _ = R.master_combine(stripes, 'flat' if flat_fix else 'bias',
                              medsec=scales if flat_fix else None, bpm=bpm_stripe, tel=tel, out=out)
    return res


def master_combine_sharded(stripes, shape, imgtype='bias', medsec=None, bpm_stripe=None, tel=None,
                           group=None, combine=None, out=None, peers=None):
    """Row-stripe sharded master combine.

    stripes     this rank's row stripe of each of the N frames: float32 [r1-r0, W] tensors
                (CUDA for the product path)
    shape       (H, W) of the full frame
    medsec      flats: the N normalisation medians (header MEDSEC, blackbox.py:4927-4941);
                required here because a frame's normalisation section spans several stripes
    bpm_stripe  flats: this rank's rows of the bad-pixel mask (edge pixels -> 1)
    combine     test hook: callable(stripes, scales, flat_fix, bpm_stripe, tel) -> stripe master
    out         optional float32 [world * stripe_rows(H, world), W] gather buffer to reuse
    peers       a ``PeerMaster`` of this shape: the combine kernel writes every rank's copy itself
                over NVLink (no NCCL all-gather); the result is ``peers.buf``

    Returns the full (H, W) master on every rank (a view of the gather buffer).  The stack-median
    kernel writes this rank's stripe straight into its slot of the gather buffer and the
    all-gather runs in place on it: no staging copy, no zero fill."""
    H, W = shape
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    r0, r1 = stripe_bounds(H, rank, world)
    nrows = stripe_rows(H, world)
    if len(stripes) == 0:
        raise ValueError('no frames to combine')
    for s in stripes:
        if tuple(s.shape) != (r1 - r0, W):
            raise ValueError('rank {}: stripe shape {} != {}'.format(rank, tuple(s.shape), (r1 - r0, W)))
    flat = imgtype == 'flat'
    if flat and medsec is None:
        raise ValueError('sharded flat combine needs the MEDSEC normalisation medians')
    dev = stripes[0].device
    if peers is not None:
        if combine is not None or tuple(peers.shape) != (H, W) or peers.world > 8:
            raise ValueError('peer-memory combine: needs the default combine, a PeerMaster of shape {} and at most '
                             '8 ranks'.format((H, W)))
        from . import reduce as R
        peers.barrier()                      # nobody is still reading the previous master out of these buffers
        if r1 > r0:
            mine = peers.buf[rank * nrows:rank * nrows + (r1 - r0)]
            R.master_combine(stripes, 'flat' if flat else 'bias', medsec=medsec if flat else None, bpm=bpm_stripe,
                             tel=tel, out=mine, out_ptrs=peers.stripe_ptrs(), multicast=peers.multicast)
        peers.barrier()                      # every stripe has landed everywhere
        return peers.buf[:H]
    if out is None:
        out = torch.empty((world * nrows, W), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (world * nrows, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('gather buffer must be a contiguous float32 tensor of shape {}'.format((world * nrows, W)))
    slot = out[rank * nrows:(rank + 1) * nrows]            # this rank's contribution (last rows may be padding)
    if r1 > r0:
        mine = slot[:r1 - r0]
        if combine is not None:
            res = combine(stripes, medsec, flat, bpm_stripe, tel)
            if not isinstance(res, torch.Tensor):
                res = torch.from_numpy(np.ascontiguousarray(res))
            mine.copy_(res)
        else:
            _default_combine(stripes, medsec, flat, bpm_stripe, tel, out=mine)
    if world > 1:
        # in place: the input is this rank's slot of the output (what NCCL calls an in-place all-gather)
        dist.all_gather_into_tensor(out, slot, group=group)
    return out[:H]


def flat_scale_from_region(frames_full, tel=None):
    """np.median over set_bb.flat_norm_sec of whole frames held on one rank (used when MEDSEC
    is not in the header; blackbox.py:4931-4932)."""
    from . import reduce as R
    sec = get_par(set_bb.flat_norm_sec, tel)
    return [float(R.exact_median(R._to_dev(f, torch.float32)[sec])) for f in frames_full]
