#!/usr/bin/env python
"""bench.py -- mocap frames solved/sec (MoSh++ Stage II) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload at N=1: BASELINE.json configs[1] -- SMPL-H (V=6890, 52 joints), one 500-frame synthetic
sequence, 53 markers, fingers on (111 free variables in Step 2).  A "step" is one pass of the hot path
over that sequence: every frame solved by the reference's schedule (Procrustes + 3 dog-legs on the first
frame of a chunk, Step 1 + Step 2 dog-legs on every frame).  At N>1 every rank solves its own sequence of
the same shape (weak scaling, no data-path collective); NCCL is used for the barrier, the
max-over-ranks of the device time and the result gather of the e2e leg.

JSON keys follow the driver contract; see DESIGN.md section 7 for how each number is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'mocap frames solved/sec (Stage-II)'
UNIT = 'frames/s'
WORKLOAD = 'C2'


def algorithmic_bytes(pk, has_velo=True, fingers=True):
    """SURVEY.md 8(d) / BASELINE.md section 3: fp32 bytes of one frame-iteration in the two-kernel
    (materialised J) formulation.  R = residual rows, n = free variables of Step 2."""
    M, n, p_red = pk.n_markers, len(pk.free_step2), pk.p_red
    R = 3 * M + (pk.prior_d + 1 if pk.prior_k else 0) + p_red + (pk.finger_hi - pk.finger_lo if fingers else 0) + 2 * pk.n_dmpl
    b_k1 = 4 * R * (n + 1) + 4 * (n + 3 * M + p_red + 8) + M
    b_k2 = 4 * R * (n + 1) + 4 * (n * (n + 1) // 2 + n)
    return dict(R=R, n=n, B_K1=b_k1, B_K2=b_k2)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms 100` child process from the warm-up to the end of the end-to-end leg."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
            for line in out.strip().splitlines():
                r = [x.strip() for x in line.split(',')]
                if len(r) >= 7:
                    self.rows.append(r)
        except Exception:
            pass

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        reasons = []
        for name, col in (('hw_slowdown', 3), ('hw_thermal_slowdown', 4), ('sw_thermal_slowdown', 5), ('sw_power_cap', 6)):
            if any(len(r) > col and r[col].lower().startswith('active') for r in self.rows):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': float(self.rows[0][1]) if self.rows[0][1].replace('.', '').isdigit() else None,
                'reasons': reasons, 'samples': len(self.rows)}


def make_workload(rank: int, frames: int | None = None):
    from moshpp_b200 import chmosh, synth
    from moshpp_b200.mocap_interface import MocapSession
    d = tempfile.mkdtemp(prefix=f'mosh_bench_r{rank}_')
    case = synth.make_case(d, WORKLOAD, frames=frames, seq_idx=rank)
    pk, opts, flags = chmosh.prepare_stageii(case['cfg'], case['markers_latent'], case['latent_labels'],
                                             case['betas'], case['marker_meta'])
    mocap = MocapSession(case['mocap_fname'], case['cfg'].mocap.unit)
    obs, vis = mocap.frames_for_labels(case['latent_labels'], range(len(mocap)))
    return case, pk, opts, obs, vis


def cpu_reference_fps(case, n_frames: int):
    """The reference's CPU algorithm (float64 oracle in reference-cost mode: full mesh + dense 3V x P
    Jacobian every evaluation, frame-serial) on the first n_frames frames of the workload."""
    from oracle import stageii
    t0 = time.time()
    out = stageii.mosh_stageii(case['mocap_fname'], case['cfg'], case['markers_latent'], case['latent_labels'],
                               case['betas'], case['marker_meta'], mode='reference_cost', max_frames=n_frames)
    dt = time.time() - t0
    st = out['stageii_debug_details']['oracle_stats']
    return st['frames'] / st['elapsed'], st, dt


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get('num_threads', 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path is not runnable here
    (chumpy / psbody.smpl absent); the oracle port in reference-cost mode is timed on the
    host cores, a bounded sample of the workload per step."""
    if rank != 0:
        return
    try:        # torchrun pins OMP_NUM_THREADS=1; the reference arm may use every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    case, pk, opts, obs, vis = make_workload(0)
    n = args.cpu_frames
    vals = []
    for s in range(args.warmup + args.steps):
        fps, st, dt = cpu_reference_fps(case, n)
        if s >= args.warmup:
            vals.append((fps, dt))
    fps = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic (seeded procedural SMPL-H model, markers, motion)',
        'config': {'workload': 'BASELINE configs[1]: SMPL-H 500-frame sequence, 53 markers, Stage II', 'frames': 500,
                   'markers': pk.n_markers, 'free_vars': len(pk.free_step2)},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
                         'sample': f'first {n} frames of the 500-frame workload per step, frame-serial, reference-cost mode '
                                   '(full 6890-vertex mesh and dense 20670x156 Jacobian per evaluation); restated reference, not chumpy'},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--chunk-len', type=int, default=None)
    ap.add_argument('--chunk-warmup', type=int, default=None)
    ap.add_argument('--precision', default='f32', choices=['f32', 'f64'])
    ap.add_argument('--cpu-frames', type=int, default=5, help='frames of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from moshpp_b200 import chmosh, lib

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: the Stage-II path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    case, pk, opts, obs, vis = make_workload(rank)
    F = obs.shape[0]
    chunk_len = args.chunk_len if args.chunk_len is not None else chmosh.auto_chunk_len(F)
    if args.chunk_warmup is None:
        args.chunk_warmup = chmosh.DEFAULT_WARMUP
    prec = {'f32': lib.MOSH2_F32, 'f64': lib.MOSH2_F64}[args.precision]
    model = lib.Model(pk, device=local_rank)
    job = model.job(F, opts, chunk_len=chunk_len, chunk_warmup=args.chunk_warmup, precision=prec)
    job.upload(obs, vis)
    job.sync()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f'cuda:{local_rank}')   # > 126 MB L2

    def flush_l2():
        flush_buf.add_(1)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        flush_l2()
        job.launch()
        job.sync()

    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = []
    for _ in range(args.steps):
        flush_l2()                      # outside the CUDA-event bracket of the step
        job.launch()
        job.sync()
        dev_ms.append(job.kernel_ms())  # CUDA events on the launching stream
    barrier()
    t_wall = time.perf_counter() - t_wall0
    totals = job.totals()
    res = job.download()
    ms_step = float(np.mean(dev_ms))
    if world > 1:
        t = torch.tensor([ms_step], device=f'cuda:{local_rank}')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step_max = float(t.item())
    else:
        ms_step_max = ms_step

    # ---- e2e: host buffers in, host results out, through the C-ABI job calls (H2D + kernel + D2H)
    h2d = obs.size * (4 if prec == lib.MOSH2_F32 else 8) + vis.size
    esz = 4 if prec == lib.MOSH2_F32 else 8
    d2h = F * (pk.p_full + pk.p_red + 3 + pk.n_dmpl + 3 * pk.n_markers + 6) * esz + F * 5 * 4
    for _ in range(2):
        job.upload(obs, vis); job.launch(); job.download()
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        job.upload(obs, vis)
        job.launch()
        res = job.download()            # includes the stream sync
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = float(np.mean(e2e_t)) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device=f'cuda:{local_rank}')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        # result gather of the sharded job (NCCL): every rank's reduced poses to rank 0
        mine = torch.from_numpy(res.pose.astype(np.float32)).to(f'cuda:{local_rank}')
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
    sampler.stop()

    solved = int(((res.status & lib.ST_SOLVED) != 0).sum())
    ab = algorithmic_bytes(pk)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = (ab['B_K1'] + ab['B_K2']) * totals['builds'] / (ms_step * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None

    line = {
        'metric': METRIC, 'value': world * F / (ms_step_max * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step_max, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic (seeded procedural SMPL-H model, markers, motion)',
        'config': {'workload': 'BASELINE configs[1]: SMPL-H 500-frame sequence, 53 markers, Stage II; one sequence per GPU',
                   'frames': F, 'markers': pk.n_markers, 'free_vars': ab['n'], 'residual_rows': ab['R'],
                   'chunk_len': chunk_len, 'chunk_warmup': args.chunk_warmup, 'chunks': job.num_chunks,
                   'l2': 'flushed between timed steps (256 MiB write)', 'frames_solved': solved,
                   'frame_iterations_per_step': totals['builds'], 'residual_evals_per_step': totals['evaluations']},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic,
                     'note': 'achieved = (B_K1+B_K2) x frame-iterations / kernel time, SURVEY 8(d) effective-bandwidth '
                             'definition; the fused kernel keeps J on chip, so DRAM traffic is far below it. peak: '
                             + ('measured (MEASURED_PEAKS.json)' if peaks else 'fallback 6650 GB/s')},
        'e2e': {'value': world * F / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': e2e_ms},
        'gpu_launches': args.steps,
        'clocks': sampler.summary(),
        'wall_s_timed_region': t_wall,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, st, dt = cpu_reference_fps(case, args.cpu_frames)
        line['cpu_baseline'] = {
            'value': fps, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
            'sample': f'first {args.cpu_frames} frames of the same 500-frame workload, frame-serial float64 oracle in '
                      f'reference-cost mode (full mesh + dense Jacobian per evaluation), {dt:.1f} s; restated reference, not chumpy'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
