#!/usr/bin/env python
"""Benchmark of the per-frame CCD reduction hot path (BASELINE.json metric: reduced frames/s,
10560^2 frames, full chain) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the full chain (gain + overscan + master bias + mask_init + master flat +
LACosmic with 4 iterations + crosstalk; blackbox.py:1479-1902) over a night batch of synthetic
10600x12000 uint16 raw BlackGEM frames: `--batch` (64, BASELINE.json config 4) frames per GPU,
resident in HBM (16.3 GB).  Weak scaling: frames are independent, frame k -> GPU k mod N, no
collective on the data path.  Per GPU `--depth` (12) frames are in flight, the overscan stage
`--ahead` (4) frames in front on a high-priority stream, stages replayed as CUDA graphs.
`--impl reference`: the CPU arm (the oracle port of the reference path on all host cores, bounded
sample per step) as a run of its own, rank 0 only.

Printed JSON (rank 0, one line):
  value      frames/s over all GPUs, raw frames already resident in HBM
  e2e        frames/s through the public API (BatchReducer.run_host) with HOST buffers on both
             sides, H2D / D2H copies inside the timed region, the reference's own files on both
             sides: the fpacked raw frame in (as the telescope delivers it; Rice-decoded on the
             device); out the reduced image as `fpack -q 16 -D -Y` writes it (blackbox.py:836:
             quantised + Rice-coded on the device, bbx_fpack_f32) and the mask as `fpack -D -Y`.
             e2e_f32_image: the same with the float32 image leaving as it is (446 MB per frame);
             e2e_uncompressed: plain uint16 frame in, image + plain mask out (round 1's definition);
             both over fewer steps
  roofline   the dominant single kernel of the chain (largest mean device time among the
             one-kernel stages, CUDA events on the launching stream INSIDE the timed region) against
             the measured HBM peak; roofline_stages lists every stage the same way
  strong_scaling  the same night batch as ONE job of `--batch` frames in total, frame k -> GPU
             k mod N (pipeline.shard_frames; the reference's pool over a fixed file list,
             blackbox.py:378): frames/s incl. pipeline fill / drain at batch/N frames per GPU
  master_sharded  BASELINE.json configs 5 and 2 on the same N GPUs: master bias of 50 binned
             5280^2 frames and master flat of 20 full 10560^2 frames, row stripes per GPU, the
             stack-median kernel writing into its slot of the gather buffer, ONE NCCL all-gather
             (distributed.master_combine_sharded; blackbox.py:4908-4984); time = max over ranks,
             compared bit for bit with the one-GPU combine of the same stack.  fused_allgather: the
             combine kernel stores every master pixel into every rank's buffer itself over NVLink
             (symmetric memory: plain peer stores / NVSwitch multicast) -- no NCCL on the data path
  link       measured pinned host<->device copy bandwidth of this rank-0 GPU (H2D alone, D2H
             alone, both at once): the ceiling of every end-to-end number
  cpu_baseline  the CPU oracle (restatement of the reference's numpy/astropy/astroscrappy
             path) on a bounded sample, one frame-slice per host core (the reference's own
             one-process-per-frame scheme, blackbox.py:378)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TEL = 'BG3'
NITER = 4
FULL_NPIX = 10560 * 10560
# algorithmic bytes per frame
ALGO_BYTES_LAC_ITER = 10 * FULL_NPIX            # img r+w (8 B) + mask r (1 B) + crmask w (1 B)
ALGO_BYTES_CHAIN_4IT = 7563e6
# stage -> (kernels behind it, algorithmic bytes per frame [SURVEY.md 8d], single kernel?)
STAGES = {
    'overscan': ('vos_rowstats, vos_fit, hos_satcount, hos_stats, vos_std, hos_fit, satlevels (S1)', 60e6, False),
    'apply': ('reduce_apply_strip_kernel<uint16> (S2: raw u16 + mbias + mflat + bpm -> img f32 + mask u8)', 1815.6e6, True),
    'mask_morph': ('ms_* / hole_* / ccl kernels (S3)', 223e6, False),
    'lacosmic': ('sp_scan (Laplacian of iteration 1) + sparse candidate / grow / clean kernels on work lists, %d iterations (S4)' % NITER, NITER * 1115.1e6, False),
    'lacosmic_finish': ('cosmic-ray bit + NCOSMICS from the CR list', None, False),
    'xtalk': ('xtalk_tile_kernel, the M-*NUM counts of the mask taken on the way + xtalk_count_sum (S5: img r+w + mask r)', 1003.6e6, False),
    'edge_fill': ('cs_hist x3 + cs_find x3 + cs_fill_edge (3 x img r + mask r)', 3 * 446.1e6 + 111.5e6, False),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=64,
                    help='frames per GPU per step (BASELINE.json config 4: a night batch of 64 frames)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--depth', type=int, default=12, help='frames in flight per GPU')
    ap.add_argument('--ahead', type=int, default=4, help='frames the overscan stage runs ahead')
    ap.add_argument('--split-priority', type=int, default=1,
                    help='1: overscan stage on a high-priority stream of its own')
    ap.add_argument('--graphs', type=int, default=1, help='1: replay the stages as CUDA graphs')
    ap.add_argument('--fill-edge', type=int, default=0,
                    help='1: append the edge-pixel fill (blackbox.py:1958-1974) to the chain (not part of the metric)')
    ap.add_argument('--fits', type=int, default=0,
                    help='1: e2e with FITS data units on the host side (byte swaps on the device)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-masters', action='store_true', help='skip the sharded master-combine workloads')
    ap.add_argument('--no-strong', action='store_true', help='skip the strong-scaling pass')
    ap.add_argument('--cpu-rows', type=int, default=0,
                    help='rows per channel of the CPU sample frame (full frame: 5280); 0 = 330 for the '
                         'cpu_baseline of the GPU arm, scaled to the step count for --impl reference')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample, one process per frame-slice
# ---------------------------------------------------------------------------------------------
def _cpu_worker(job):
    seed, rows = job
    os.environ['OMP_NUM_THREADS'] = '1'
    from blackbox_b200 import set_bb, synth
    from oracle import reduce as R
    set_bb.ysize_chan = rows
    q = rows // 4
    set_bb.hos_sat_ypix_lim = {'BG2': (2 * q, 4 * q), 'BG3': (q, 2 * q), 'BG4': (q, 2 * q)}
    raw, _ = synth.make_raw(TEL, seed)
    shape = (2 * rows, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(TEL, 9, shape)
    coeffs = synth.make_xtalk(3)[3]
    t0 = time.perf_counter()
    R.reduce_frame(raw, TEL, mbias, mflat, bpm, coeffs, niter=NITER)
    return time.perf_counter() - t0


class CpuArm:
    """The oracle on all host cores, one process per frame slice (pool kept alive across steps)."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context('spawn').Pool(cores)
        self.pool.map(_cpu_worker, [(1, 40)] * cores)          # start-up: imports, oracle library load

    def sample(self, rows):
        """-> (frames/s equivalent, wall seconds, description)"""
        jobs = [(5000 + i, rows) for i in range(self.cores)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
        frac = (2 * rows * 8 * 1320) / float(FULL_NPIX)
        value = self.cores * frac / wall
        sample = ('{} slices of {}x10560 px (1/{:.1f} of a 10560^2 frame each), full chain niter={}, '
                  'one process per slice, OMP_NUM_THREADS=1'.format(self.cores, 2 * rows, 1 / frac, NITER))
        return value, wall, sample

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample(rows, cores):
    arm = CpuArm(cores)
    try:
        return arm.sample(rows)
    finally:
        arm.close()


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (the oracle port; the reference itself cannot be
    installed here) on all host cores.  Each step is a bounded sample: a
    330-row slice per core costs ~35 s, so the slice height is scaled to keep the whole
    --steps / --warmup run within a few minutes."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = args.cpu_rows
    if rows <= 0:
        budget = 150.0 / max(args.steps + args.warmup, 1)          # seconds per step
        rows = int(max(96, min(330, 330 * budget / 35.0))) // 4 * 4   # >= 96: keeps the per-slice fixed costs small
    arm = CpuArm(cores)
    vals = []
    try:
        for i in range(args.warmup + args.steps):
            v, wall, sample = arm.sample(rows)
            if i >= args.warmup:
                vals.append((v, wall))
    finally:
        arm.close()
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(w for _, w in vals) / len(vals)
    line = {
        'impl': 'reference', 'metric': 'reduced frames/sec (10560^2 raw, full chain)',
        'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(args):
    return {'workload': 'night batch of synthetic {} 10600x12000 uint16 raw frames ({} per GPU per '
                        'step), full chain: gain+overscan+master bias+mask_init+master flat+'
                        'LACosmic(niter={})+crosstalk -> 10560x10560 f32 image + u8 mask'
                        .format(TEL, args.batch, NITER),
            'frames_per_gpu_per_step': args.batch, 'lacosmic_niter': NITER,
            'l2': 'inputs larger than L2 (254 MB raw + 1 GB masters per frame vs 126 MB L2)',
            'parallelism': 'frame k -> GPU k mod N, no collective; {} frames in flight per GPU, overscan stage '
                           'of the next frame {}ahead of the HBM-bound stage of the current one, stages {}'.format(
                               args.depth, 'on a high-priority stream ' if args.split_priority else '',
                               'replayed as CUDA graphs' if args.graphs else 'launched eagerly')}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from blackbox_b200 import reduce as R, set_bb, synth
    from blackbox_b200.pipeline import BatchReducer

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # ---- synthetic inputs (host, seeded), then resident in HBM ------------------------------
    B = args.batch
    nbase = min(2, B)
    bases = [synth.make_raw(TEL, 4001 + 17 * rank + i)[0] for i in range(nbase)]
    red_shape = (2 * set_bb.ysize_chan, 8 * set_bb.xsize_chan)
    mbias, mflat, bpm = synth.make_masters(TEL, 9, red_shape)
    coeffs = synth.make_xtalk(3)[3]
    raws = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    for k in range(B):
        base = R._to_dev(bases[k % nbase]).view(torch.int16).to(torch.int32) & 0xffff
        if k >= nbase:            # distinct noise realisation per frame
            base = base + torch.randint(-3, 4, base.shape, device=dev, generator=gen, dtype=torch.int32)
        raws.append(base.clamp_(0, 65535).to(torch.int16).view(torch.uint16).contiguous())
    del base
    batch = BatchReducer(TEL, raws[0].shape, depth=args.depth, ahead=args.ahead, split_priority=bool(args.split_priority),
                         use_graphs=bool(args.graphs), fill_edge=bool(args.fill_edge), mbias=mbias, mflat=mflat, bpm=bpm,
                         coeffs=coeffs, niter=NITER)
    pipe = batch.pipes[0]
    nout = args.depth              # frame k -> output buffer k % depth: every (raw, output) pair recurs each step
    out_imgs = [torch.empty(red_shape, dtype=torch.float32, device=dev) for _ in range(nout)]
    out_masks = [torch.empty(red_shape, dtype=torch.uint8, device=dev) for _ in range(nout)]
    out_img, out_mask = out_imgs[0], out_masks[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spline_cols = [0]

    def step():
        redo = 0
        for res in batch.run(raws, out_imgs, out_masks, fill_header=True):
            redo += res.redo
            spline_cols[0] += res.spline_columns
        return redo

    for _ in range(args.warmup):
        step()
    for p in batch.pipes:
        p.enable_stage_timing(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)          # nvidia-smi needs a moment before its first sample
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    redo = 0
    for _ in range(args.steps):
        redo += step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    frames = world * B * args.steps
    value = frames / (ms_total * 1e-3)
    launches = world * args.steps * B * pipe_launches(TEL, NITER)

    # ---- roofline: one LACosmic iteration, timed with CUDA events on the launching stream ---
    roof = stages = None
    alone = measure_kernel_alone(batch, raws, out_imgs, out_masks) if rank == 0 else None
    barrier()
    if rank == 0:
        roof, stages = stage_rooflines(batch.pipes, alone)
    for p in batch.pipes:
        p.enable_stage_timing(False)

    # ---- end to end: pinned host raw in, pinned host image + mask out -----------------------
    e2e = e2e_plain = e2e_f32 = None
    if not args.no_e2e:
        def e2e_entry(mode, steps):
            ms, h2d, d2h = measure_e2e(args, batch, raws, red_shape, barrier, mode=mode, steps=steps)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return {'value': world * B * steps / (float(t.item()) * 1e-3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': int(world * B * h2d), 'd2h_bytes_per_step': int(world * B * d2h),
                    'steps': steps}
        e2e = e2e_entry('fz', args.steps)
        e2e['io'] = ('the files of the reference on both sides -- in: fpacked raw frame (Rice-coded heap + tile '
                     'descriptors, pinned host memory), decoded on the device; out: the reduced image as `fpack -q 16 '
                     '-D -Y` writes it (blackbox.py:836: rows quantised at 1/16 of their noise with subtractive '
                     'dither, Rice-coded: bbx_fpack_f32) + the mask as `fpack -D -Y` (losslessly Rice-coded uint8)')
        e2e_f32 = e2e_entry('f32', max(1, min(args.steps, 3)))
        e2e_f32['io'] = 'in: fpacked raw frame; out: float32 image as it is + Rice-coded uint8 mask'
        e2e_plain = e2e_entry('plain', max(1, min(args.steps, 3)))
        e2e_plain['io'] = 'in: uint16 raw frame; out: float32 image + plain uint8 mask (round 1\'s e2e)'

    # ---- strong scaling: the night batch as one job of B frames in total -----------------------
    strong = None
    if not args.no_strong:
        strong = measure_strong(args, batch, raws, out_imgs, out_masks, rank, world, dev, barrier)

    # ---- sharded master combine (configs 5 and 2): row stripes + one all-gather ------------------
    masters = None
    if not args.no_masters:
        del raws[2:]                       # the stacks need the room of the resident night batch
        torch.cuda.empty_cache()
        masters = measure_master_sharded(args, rank, world, dev, barrier)
    # the link ceiling of the end-to-end number: every rank copies at the same time (at N > 1 the
    # ranks share the host's memory system), rank 0 reports its own figures and the sum over ranks
    link = None
    if not args.no_e2e:
        barrier()
        link = measure_link(dev)
        if world > 1:
            both = torch.tensor([link['both_each_GBs']], dtype=torch.float64, device=dev)
            lo = both.clone()
            dist.all_reduce(both, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            link['all_ranks_concurrently'] = {'ranks': world, 'both_each_GBs_sum': float(both.item()),
                                              'both_each_GBs_min': float(lo.item())}

    if rank == 0:
        cpu = None
        if args.gpus == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, wall, sample = cpu_sample(args.cpu_rows or 330, cores)
            cpu = {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                   'wall_s': wall}
        line = {
            'metric': 'reduced frames/sec (10560^2 raw, full chain)', 'value': value,
            'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args), 'roofline': roof, 'roofline_stages': stages,
            'roofline_stages_note': 'ms_per_frame = device time between CUDA events around the stage INSIDE the timed '
                                    'region, %d frames in flight: the stages of different frames share the GPU, so the '
                                    'figures add up to several times the frame time (ms_per_step / frames); kernels '
                                    'timed alone: profiles/r02_kbench.txt, r02_ncu_full_summary.txt' % args.depth,
            'cpu_baseline': cpu,
            'clocks': clocks, 'e2e': e2e, 'e2e_f32_image': e2e_f32, 'e2e_uncompressed': e2e_plain,
            'gpu_launches': launches,
            'chain_hbm_frac': (ALGO_BYTES_CHAIN_4IT * frames / world / (ms_total * 1e-3)) / (peak_hbm()[0] * 1e9),
            'frames_redone': redo, 'host_spline_columns': spline_cols[0],
            'graph_replays': sum(p.graph_replays for p in batch.pipes),
            'strong_scaling': strong, 'master_sharded': masters, 'link': link,
            'header': 'fill_header=True in both timed paths: every header scalar (BIASM/RDN x16, BIASMEAN, '
                      'RDNOISE, SATLEV x16, NOBJ-SAT, NCOSMICS, LAC-NIT, M-*NUM) arrives in one pinned copy per frame',
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_strong(args, batch, raws, out_imgs, out_masks, rank, world, dev, barrier):
    """Config 4 as the reference runs it: ONE batch of `--batch` frames in total, frame k -> GPU
    k mod N (pipeline.shard_frames), nothing exchanged.  Each rank reduces its share of its resident
    frames; time = max over ranks of the device time of one whole job, mean of 3 jobs after one
    warm-up job (pipeline fill and drain are inside, which is the point)."""
    import torch
    import torch.distributed as dist
    from blackbox_b200.pipeline import shard_frames
    mine = [raws[i] for i in range(len(shard_frames(args.batch, rank, world)))]

    def job():
        if mine:
            batch.run(mine, out_imgs, out_masks)

    job()
    reps = 3
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        job()
        barrier()                          # a job is over when its last rank is
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {'frames_total': args.batch, 'frames_per_gpu': len(mine), 'ms_per_job': ms,
            'value': args.batch / (ms * 1e-3), 'unit': 'frames/s', 'scaling': 'strong',
            'note': 'one job = the whole batch sharded frame k -> GPU k mod N; barrier after every job'}


def measure_master_sharded(args, rank, world, dev, barrier):
    """BASELINE.json config 5 (master bias of 50 2x2-binned 5280^2 frames) and config 2 (master flat
    of 20 full 10560^2 frames) through distributed.master_combine_sharded: every rank holds its
    row stripe of every frame, combines it with bbx_stack_median straight into its slot of the
    gather buffer, one in-place NCCL all-gather assembles the master on every rank.  Rank 0 also
    holds the whole stack and runs the one-GPU combine: the sharded master must equal it bit for
    bit.  Times: CUDA events, 10 repetitions after 3 warm-ups, max over ranks."""
    import torch
    import torch.distributed as dist
    from blackbox_b200 import distributed as D, reduce as R
    peak = peak_hbm()[0]
    out = {}
    cases = (('bias50_5280', 'bias', 50, (5280, 5280), 5000), ('flat20_10560', 'flat', 20, (10560, 10560), 2000))
    for name, imgtype, n, shape, seed0 in cases:
        H, W = shape
        r0, r1 = D.stripe_bounds(H, rank, world)
        gen = torch.Generator(device=dev)
        stripes, full = [], []
        for i in range(n):
            gen.manual_seed(seed0 + i)                       # the same frame on every rank
            if imgtype == 'bias':
                f = torch.randn(shape, device=dev, generator=gen) * 9.0
            else:
                f = (1.0 + 0.01 * torch.randn(shape, device=dev, generator=gen)) * (20000.0 * (0.8 + 0.4 * i / n))
            stripes.append(f[r0:r1].clone())
            if rank == 0 and world > 1:
                full.append(f)
            del f
        medsec = [20000.0 * (0.8 + 0.4 * i / n) for i in range(n)] if imgtype == 'flat' else None
        bpm = None
        if imgtype == 'flat':
            bpm = torch.zeros(shape, dtype=torch.uint8, device=dev)
            bpm[:20] = 32
            bpm[-20:] = 32
            bpm[:, :20] = 32
            bpm[:, -20:] = 32
        gather = torch.empty((world * D.stripe_rows(H, world), W), dtype=torch.float32, device=dev)

        def sharded():
            return D.master_combine_sharded(stripes, shape, imgtype, medsec=medsec,
                                            bpm_stripe=None if bpm is None else bpm[r0:r1], tel=TEL, out=gather)

        def local_only():
            R.master_combine(stripes, imgtype, medsec=medsec, bpm=None if bpm is None else bpm[r0:r1], tel=TEL,
                             out=gather[rank * D.stripe_rows(H, world):rank * D.stripe_rows(H, world) + (r1 - r0)])

        def timed(fn, reps=10, warm=3):
            for _ in range(warm):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        ms = timed(sharded)
        ms_kernel = timed(local_only)
        master = sharded()
        # the same with the all-gather done BY the combine kernel: every master pixel stored into
        # every rank's buffer over NVLink (symmetric memory), plain peer stores and NVSwitch multicast
        peer = {}
        if world > 1:
            for label, mc in (('peer_stores', False), ('multicast', True)):
                try:
                    pm = D.PeerMaster(shape, multicast=mc)
                    if mc and not pm.multicast:
                        peer[label] = {'unavailable': 'no multicast support on this fabric / torch build'}
                        continue
                    fn = lambda: D.master_combine_sharded(stripes, shape, imgtype, medsec=medsec,
                                                          bpm_stripe=None if bpm is None else bpm[r0:r1], tel=TEL, peers=pm)
                    ms_p = timed(fn)
                    got = fn()
                    same = torch.tensor([1 if torch.equal(got.view(torch.int32), master.view(torch.int32)) else 0],
                                        dtype=torch.int32, device=dev)
                    dist.all_reduce(same, op=dist.ReduceOp.MIN)
                    peer[label] = {'ms': ms_p, 'equal_to_allgather_path': bool(same.item())}
                    del pm, got
                except Exception as exc:           # symmetric memory not available here: say so, keep the bench alive
                    peer[label] = {'unavailable': '{}: {}'.format(type(exc).__name__, str(exc)[:200])}
        equal, ms_1gpu = None, ms
        if world > 1:
            flag = torch.ones(1, dtype=torch.int32, device=dev)
            t1 = torch.zeros(1, dtype=torch.float64, device=dev)
            if rank == 0:
                ref = torch.empty(shape, dtype=torch.float32, device=dev)
                one = lambda: R.master_combine(full, imgtype, medsec=medsec, bpm=bpm, tel=TEL, out=ref)
                for _ in range(3):
                    one()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    one()
                e1.record()
                torch.cuda.synchronize()
                t1[0] = e0.elapsed_time(e1) / 10
                same = torch.equal(master.view(torch.int32), ref.view(torch.int32))
                flag[0] = 1 if same else 0
                del ref
            dist.broadcast(flag, 0)
            dist.broadcast(t1, 0)
            equal, ms_1gpu = bool(flag.item()), float(t1.item())
        nbytes = (n + 1) * H * W * 4 + (H * W if bpm is not None else 0)
        for v in peer.values():
            if 'ms' in v:
                v['speedup_vs_1gpu'] = ms_1gpu / v['ms']
                v['frac'] = nbytes / (v['ms'] * 1e-3) / 1e9 / (peak * world)
        best = min([ms] + [v['ms'] for v in peer.values() if 'ms' in v and v.get('equal_to_allgather_path')])
        out[name] = {'frames': n, 'shape': [H, W], 'imgtype': imgtype, 'rows_per_gpu': r1 - r0, 'ms': ms,
                     'fused_allgather': peer, 'ms_best': best, 'speedup_best_vs_1gpu': ms_1gpu / best,
                     'ms_stack_median_stripe': ms_kernel, 'allgather_ms': max(ms - ms_kernel, 0.0),
                     'ms_1gpu': ms_1gpu, 'speedup_vs_1gpu': ms_1gpu / ms, 'equal_to_1gpu': equal,
                     'algorithmic_bytes': nbytes, 'achieved_GBs': nbytes / (ms * 1e-3) / 1e9,
                     'frac': nbytes / (ms * 1e-3) / 1e9 / (peak * world),
                     'frac_note': 'of N x the measured HBM peak; the all-gather is inside the time'}
        del stripes, full, gather, master, bpm
        torch.cuda.empty_cache()
    return out


def measure_link(dev):
    """Pinned host <-> device copy bandwidth of this GPU: H2D alone, D2H alone, both directions at
    once (what BatchReducer.run_host does) -- the ceiling of the end-to-end number."""
    import torch
    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9

    run(True, True, 1)
    both = run(True, True)
    return {'h2d_GBs': run(True, False), 'd2h_GBs': run(False, True), 'both_each_GBs': both, 'unit': 'GB/s',
            'note': 'pinned memory, 512 MiB copies; both_each = per direction with H2D and D2H running concurrently'}


def pipe_launches(tel, niter):
    """Kernels of libbbx.so launched per frame by FramePipeline (memsets and copies not counted;
    equal to the ncu launch list profiles/r02_launches_bench.csv, 66 per BlackGEM frame at niter 4):
    overscan 6 (+1 BlackGEM saturated-column count) + header means 1, fused per-pixel pass 1,
    LACosmic set-up 3 (init, background gather + sample), sparse mask morphology 12, Laplacian scan +
    background rank 2, per iteration 7 (two candidate, three growth kernels, flag clean-up, control)
    + 1 cleaning pass, + 1 rescan per further iteration, cosmic-ray bit + object count 3, crosstalk
    with the per-bit mask counts 2."""
    return (6 + (0 if tel.startswith('ML') else 1) + 1 + 1 + 3 + 12 + 2 + 8 * niter + max(niter - 1, 0) + 3 + 2) if niter > 0 \
        else (6 + (0 if tel.startswith('ML') else 1) + 1 + 1 + 12 + 2)


def measure_kernel_alone(batch, raws, out_imgs, out_masks):
    """The fused per-pixel kernel on its own: one launch per frame of the resident batch, back to
    back on one stream with nothing else on the GPU, every launch between its own pair of CUDA
    events on that stream.  Inputs larger than L2 (each launch reads another 254 MB frame and the
    1 GB of masters).  -> (mean ms, min ms, launches)"""
    import torch
    p = batch.pipes[0]
    torch.cuda.synchronize()
    evs = []
    for rep in range(2):
        evs = []
        for k, raw in enumerate(raws):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p.apply_only(raw, out_imgs[k % len(out_imgs)], out_masks[k % len(out_masks)])
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return sum(ts) / len(ts), min(ts), len(ts)


def peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """kernel name -> DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of the
    committed `ncu --set full` capture, profiles/*_traffic.json); {} if there is none."""
    import glob
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json'))):
        try:
            with open(path) as fh:
                out.update(json.load(fh))
        except (OSError, ValueError):
            pass
    return out


def stage_rooflines(pipes, alone=None):
    """Mean device time of every stage of the chain over the frames of the timed region (CUDA
    events recorded on the launching streams by FramePipeline), each against the measured HBM
    peak with SURVEY.md 8(d)'s algorithmic bytes.  -> (roofline of the dominant single kernel,
    list of all stages)."""
    peak, how = peak_hbm()
    acc = {}
    for p in pipes:
        for name, (ms, n) in p.stage_times_ms().items():
            a = acc.setdefault(name, [0.0, 0])
            a[0] += ms * n
            a[1] += n
    traffic = ncu_traffic()
    stages, best = [], None
    for name, (tot, n) in acc.items():
        ms = tot / n
        kernels, nbytes, single = STAGES.get(name, (name, None, False))
        ent = {'stage': name, 'kernels': kernels, 'ms_per_frame': ms, 'frames': n,
               'algorithmic_bytes': nbytes, 'single_kernel': single}
        if nbytes:
            ent['achieved'] = nbytes / (ms * 1e-3) / 1e9
            ent['frac'] = ent['achieved'] / peak
        stages.append(ent)
        if single and (best is None or ms > best['ms_per_frame']):
            best = ent
    stages.sort(key=lambda e: -e['ms_per_frame'])
    if best is None:
        return None, stages
    kname = best['kernels'].split(' ')[0].split('<')[0]
    in_pipe = {'ms_per_launch': best['ms_per_frame'], 'launches_timed': best['frames'], 'achieved': best['achieved'],
               'frac': best['frac'],
               'note': 'the same kernel between CUDA events INSIDE the timed region: a dozen frames are in flight and '
                       'other streams\' kernels share the SMs and the HBM during the launch, so this is a share of a '
                       'busy GPU, not the kernel\'s own speed'}
    if alone is None or best['stage'] != 'apply':
        roof = dict(in_pipe, bound='hbm', kernel=best['kernels'], peak=peak, unit='GB/s', traffic=traffic.get(kname),
                    algorithmic_bytes=best['algorithmic_bytes'], peak_source=how)
        return roof, stages
    ms, ms_min, n = alone
    ach = best['algorithmic_bytes'] / (ms * 1e-3) / 1e9
    roof = {'bound': 'hbm', 'kernel': best['kernels'], 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
            'frac': ach / peak, 'traffic': traffic.get(kname), 'ms_per_launch': ms, 'ms_min': ms_min,
            'launches_timed': n, 'algorithmic_bytes': best['algorithmic_bytes'], 'peak_source': how,
            'note': 'dominant single kernel of the chain (largest device time among the one-kernel stages), timed '
                    'ALONE in bench.py: one launch per frame of the resident batch back to back on one stream, each '
                    'between its own CUDA events, inputs larger than L2 (another 254 MB frame per launch); peak = the '
                    'measured copy bandwidth (burst)',
            'in_pipeline': in_pipe}
    return roof, stages


def measure_e2e(args, batch, raws, red_shape, barrier, mode='fz', steps=None):
    """Public API with host buffers (BatchReducer.run_host), H2D / D2H copies, the chain and the
    codecs overlapping on their own streams, `--depth` frames in flight.
    mode 'fz' (the headline): every raw frame arrives as the telescope delivers it -- an fpacked
    .fits.fz, i.e. a pinned host buffer with its Rice-coded heap (~1/3 of the frame) plus tile
    descriptors, decoded on the device -- and both products leave as the reference writes them to
    disk: the image as `fpack -q 16 -D -Y` (quantised + Rice-coded on the device, bbx_fpack_f32),
    the mask as `fpack -D -Y` (losslessly Rice-coded uint8).
    mode 'f32': as 'fz', but the float32 image leaves as it is (446 MB per frame).
    mode 'plain': pinned uint16 raw frames in, float32 image + plain uint8 mask out (round 1's e2e).
    -> (ms, h2d bytes per frame, d2h bytes per frame: counted from the copies made)"""
    import torch
    from blackbox_b200 import fitsio, reduce as R
    B = args.batch
    steps = args.steps if steps is None else steps
    nring = min(B, 8)              # 8 distinct pinned raw frames, cycled through the batch
    nout = min(max(2, args.depth), B) if B > 1 else 1
    packed = mode in ('fz', 'f32')
    if mode == 'fz':
        host_img = [torch.empty(batch.img_fz_bytes(), dtype=torch.uint8).pin_memory() for _ in range(nout)]
    else:
        host_img = [torch.empty(red_shape, dtype=torch.float32).pin_memory() for _ in range(nout)]
    if packed:
        ring = []
        for k in range(nring):     # set-up, outside the timed region: fpack the synthetic frames
            heap, lens = R.rice_encode(raws[k])
            offs = np.concatenate(([0], np.cumsum(lens.astype(np.int64))[:-1]))
            info = dict(shape=tuple(raws[k].shape), bitpix=16, bytepix=2, bzero=32768.0, bscale=1.0, blocksize=32)
            ring.append(fitsio.CompressedImage({}, torch.from_numpy(heap).pin_memory(), offs, lens.astype(np.int32), info))
            ring[-1].descriptors()
        h2d = sum(c.heap.numel() + c.descriptors().numel() for c in ring) / len(ring)
        nbytes = batch.mask_fz_bytes()
        host_mask = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(nout)]
        kw = dict(mask_fz=True, img_fz=(mode == 'fz'))
    else:
        ring = [torch.empty(raws[0].shape, dtype=torch.uint16).pin_memory() for _ in range(nring)]
        for k in range(nring):
            ring[k].copy_(raws[k].cpu())
            if args.fits:             # as the data unit of a raw FITS file: big-endian int16, BZERO 32768
                a = ring[k].view(torch.int16).numpy()
                be = (a.view(np.uint16).astype(np.int32) - 32768).astype('>i2')
                a[...] = be.view(np.int16)
        h2d = raws[0].numel() * 2
        host_mask = [torch.empty(red_shape, dtype=torch.uint8).pin_memory() for _ in range(nout)]
        kw = dict(fits=bool(args.fits))
    host_raw = [ring[k % nring] for k in range(B)]

    copied = [0, 0]

    def run(nsteps):
        redo = 0
        for _ in range(nsteps):
            for res in batch.run_host(host_raw, host_img, host_mask, fill_header=True, **kw):
                redo += res.redo
            copied[0] += batch.d2h_bytes
            copied[1] += B
        return redo

    run(1)
    copied[:] = [0, 0]
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    w0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    t1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3
    del host_img, host_mask
    return max(t0.elapsed_time(t1), wall_ms), h2d, copied[0] / max(copied[1], 1)


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               '--nproc-per-node', str(args.gpus), '--master-addr', '127.0.0.1',
               '--master-port', os.environ.get('MASTER_PORT', '29511'), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
