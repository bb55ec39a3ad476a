#!/usr/bin/env python
"""Benchmark of the per-frame CCD reduction hot path (BASELINE.json metric: reduced frames/s,
10560^2 frames, full chain) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the full chain (gain + overscan + master bias + mask_init + master flat +
LACosmic + crosstalk; blackbox.py:1479-1902) over a batch of synthetic 10600x12000 uint16 raw
BlackGEM frames, `--batch` frames per GPU (weak scaling: frames are independent, frame k ->
GPU k mod N, no collective on the data path).

Printed JSON (rank 0, one line):
  value      frames/s over all GPUs, raw frames already resident in HBM
  e2e        frames/s through the public API with HOST (pinned) raw frames in and HOST image +
             mask out, H2D / D2H copies inside the timed region
  roofline   the LACosmic iteration (the dominant unit of work) against the measured HBM peak
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4, help='frames per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--depth', type=int, default=2, help='frames in flight per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-rows', type=int, default=330,
                    help='rows per channel of the CPU sample frame (full frame: 5280)')
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


def cpu_sample(rows, cores):
    """-> (frames/s equivalent, wall seconds, description)"""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    jobs = [(5000 + i, rows) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 40)] * cores)              # start-up, library build / load
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    frac = (2 * rows * 8 * 1320) / float(FULL_NPIX)
    value = cores * frac / wall
    sample = ('{} slices of {}x10560 px (1/{:.0f} of a 10560^2 frame each), full chain niter={}, '
              'one process per slice, OMP_NUM_THREADS=1'.format(cores, 2 * rows, 1 / frac, NITER))
    return value, wall, sample


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        v, wall, sample = cpu_sample(args.cpu_rows, cores)
        if i >= args.warmup:
            vals.append((v, wall))
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
            'parallelism': 'frame k -> GPU k mod N, no collective; {0} frames in flight per GPU on {0} streams'.format(args.depth)}


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
    from blackbox_b200.pipeline import BatchReducer, FramePipeline

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
    batch = BatchReducer(TEL, raws[0].shape, depth=args.depth, mbias=mbias, mflat=mflat, bpm=bpm, coeffs=coeffs,
                         niter=NITER)
    pipe = batch.pipes[0]
    out_imgs = [torch.empty(red_shape, dtype=torch.float32, device=dev) for _ in range(args.depth)]
    out_masks = [torch.empty(red_shape, dtype=torch.uint8, device=dev) for _ in range(args.depth)]
    out_img, out_mask = out_imgs[0], out_masks[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spline_cols = [0]

    def step():
        redo = 0
        for res in batch.run(raws, out_imgs, out_masks):
            redo += res.redo
            spline_cols[0] += res.spline_columns
        return redo

    for _ in range(args.warmup):
        step()
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
    launches = args.steps * B * pipe_launches(TEL, NITER)

    # ---- roofline: one LACosmic iteration, timed with CUDA events on the launching stream ---
    roof = None
    if rank == 0:
        roof = measure_lacosmic_iteration(pipe, raws, out_img, out_mask)

    # ---- end to end: pinned host raw in, pinned host image + mask out -----------------------
    e2e = None
    if not args.no_e2e:
        e2e_ms = measure_e2e(args, pipe, raws, red_shape, dev, barrier)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        raw_bytes = raws[0].numel() * 2
        e2e = {'value': world * B * args.steps / (e2e_ms * 1e-3), 'unit': 'frames/s',
               'h2d_bytes_per_step': B * raw_bytes,
               'd2h_bytes_per_step': B * (out_img.numel() * 4 + out_mask.numel())}

    if rank == 0:
        cpu = None
        if args.gpus == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, wall, sample = cpu_sample(args.cpu_rows, cores)
            cpu = {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                   'wall_s': wall}
        line = {
            'metric': 'reduced frames/sec (10560^2 raw, full chain)', 'value': value,
            'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args), 'roofline': roof, 'cpu_baseline': cpu,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'chain_hbm_frac': (ALGO_BYTES_CHAIN_4IT * frames / world / (ms_total * 1e-3)) / (peak_hbm()[0] * 1e9),
            'frames_redone': redo, 'host_spline_columns': spline_cols[0],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def pipe_launches(tel, niter):
    """Kernels of libbbx.so launched per frame by FramePipeline.enqueue (memsets not counted):
    overscan 8 (+1 BlackGEM saturated-column count), header means 1, fused apply 1, sparse mask
    morphology 11, LACosmic 2 + 7 in the first iteration + 7 per iteration, cosmic bit + object count 3,
    crosstalk 1."""
    return 8 + (0 if tel.startswith('ML') else 1) + 1 + 1 + 11 + 2 + 7 + 7 * niter + 3 + 1


def peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def measure_lacosmic_iteration(pipe, raws, out_img, out_mask, reps=3):
    """Average device time of one LACosmic iteration (stage1 + stage2 + grow + control + clean
    kernels) on full frames, CUDA events on the launching stream."""
    import ctypes as C
    import torch
    from blackbox_b200 import reduce as R, set_bb
    from blackbox_b200._lib import call
    H, W = out_img.shape
    times = []
    for k in range(min(reps, len(raws))):
        # a fresh reduced frame + mask (chain without LACosmic / crosstalk), then begin
        R.overscan_enqueue(raws[k], pipe.geom, pipe.tel, gain=pipe.gain, state=pipe.st)
        call('bbx_header_means', R._ptr(pipe.st.biasm), R._ptr(pipe.st.std_vos), R._ptr(pipe.means), R._stream())
        R.apply_enqueue(raws[k], pipe.geom, pipe.tel, st=pipe.st, gain=pipe.gain, mbias=pipe.mbias,
                        mflat=pipe.mflat, bpm=pipe.bpm, want_mask=True, out_img=out_img, out_mask=out_mask,
                        mwork=pipe.mwork)
        R.mask_morph_enqueue(out_mask, pipe.tel, pipe.mwork)
        call('bbx_lacosmic_begin', R._ptr(out_img), R._ptr(out_mask), R._ptr(pipe.crmask), H, W, NITER, 0,
             R._ptr(pipe.lwork.buf), R._ptr(pipe.lwork.info), R._stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call('bbx_lacosmic_iteration', R._ptr(out_img), R._ptr(out_mask), R._ptr(pipe.crmask), H, W,
             float(set_bb.get_par(set_bb.sigclip, pipe.tel)), float(np.float32(set_bb.sigfrac)),
             float(set_bb.objlim), 0.0, R._ptr(pipe.means[1:]), 0, 0, R._ptr(pipe.lwork.buf),
             R._ptr(pipe.lwork.info), R._stream())
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    peak, how = peak_hbm()
    achieved = ALGO_BYTES_LAC_ITER / (ms * 1e-3) / 1e9
    return {'bound': 'hbm', 'kernel': 'lacosmic_iteration (sp_scan dense pass + sparse candidate/grow/clean kernels)',
            'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
            'traffic': None, 'ms_per_launch': ms, 'algorithmic_bytes': ALGO_BYTES_LAC_ITER,
            'peak_source': how,
            'note': 'lazy LACosmic: one dense Laplacian pass (4 B/px read) per iteration, medians only at candidates; see DESIGN.md'}


def measure_e2e(args, pipe, raws, red_shape, dev, barrier):
    """Public API with host buffers: pinned uint16 raw frames in, pinned f32 image + u8 mask
    out; H2D, chain and D2H overlap on three streams with double buffering."""
    import torch
    B = args.batch
    host_raw = [torch.empty(raws[0].shape, dtype=torch.uint16).pin_memory() for _ in range(B)]
    for k in range(B):
        host_raw[k].copy_(raws[k].cpu())
    host_img = [torch.empty(red_shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_mask = [torch.empty(red_shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d_raw = [torch.empty(raws[0].shape, dtype=torch.uint16, device=dev) for _ in range(2)]
    d_img = [torch.empty(red_shape, dtype=torch.float32, device=dev) for _ in range(2)]
    d_mask = [torch.empty(red_shape, dtype=torch.uint8, device=dev) for _ in range(2)]
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    status = torch.zeros((B, 1), dtype=torch.int32).pin_memory()

    def run(nsteps):
        n = 0
        for _ in range(nsteps):
            for k in range(B):
                b = n % 2
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_cmp[b])              # raw buffer b free again
                    d_raw[b].copy_(host_raw[k], non_blocking=True)
                    ev_in[b].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[b])
                    s_cmp.wait_event(ev_out[b])             # output buffer b copied out
                    pipe.enqueue(d_raw[b], d_img[b], d_mask[b])
                    pipe.enqueue_status(status[k])
                    ev_cmp[b].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b])
                    host_img[b].copy_(d_img[b], non_blocking=True)
                    host_mask[b].copy_(d_mask[b], non_blocking=True)
                    ev_out[b].record(s_out)
                n += 1
        for s in (s_in, s_cmp, s_out):
            s.synchronize()
        if int(status.sum()) != 0:
            raise RuntimeError('e2e: hole filling of a frame did not converge (status {})'.format(status.tolist()))

    run(1)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    w0 = time.perf_counter()
    run(args.steps)
    torch.cuda.synchronize()
    t1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3
    return max(t0.elapsed_time(t1), wall_ms)


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
