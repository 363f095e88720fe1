"""N>1 path on CPU: world_size-2 gloo run of the scatter / solve / gather plumbing.
The per-rank solve is the TEST-ONLY host build of the device source, so the numbers are real."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assignment_balances_frames():
    from moshpp_b200.shard import assign_sequences
    a = assign_sequences([4000] * 32, 8)
    assert all(len(x) == 4 for x in a) and sorted(sum(a, [])) == list(range(32))
    b = assign_sequences([100, 900, 500, 500], 2)
    assert sorted(sum(b, [])) == [0, 1, 2, 3]
    loads = [sum([100, 900, 500, 500][i] for i in r) for r in b]
    assert max(loads) - min(loads) <= 100


def _worker(rank, world, port, case_dir, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import ctypes as C
    import torch.distributed as dist
    from moshpp_b200 import build, lib, shard, synth
    from moshpp_b200.mocap_interface import MocapSession
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_seq, F = 3, 6
    rank_dir = os.path.join(case_dir, f'rank{rank}')     # seeded generator: identical content, no concurrent writes to one file
    os.makedirs(rank_dir, exist_ok=True)
    cases = [synth.make_case(rank_dir, 'C4', frames=F, seq_idx=i) for i in range(n_seq)]
    emu = C.CDLL(build.build_emu())

    def solve_one(i, obs, vis):
        pk, cfg = cases[i]['pack'], cases[i]['cfg']
        h = lib.DescHolder(pk)
        opt = lib.make_options(cfg.opt_settings.weights, optimize_fingers=True)
        res = lib.ResultArrays(F, lib.pack_dims(pk))
        o = np.ascontiguousarray(obs, dtype=np.float64)
        v = np.ascontiguousarray(vis, dtype=np.uint8)
        sched = lib.make_schedule(0, 0)
        rc = emu.mosh2_emu_solve(C.byref(h.desc), C.byref(opt), F, o.ctypes.data_as(lib._f64p), v.ctypes.data_as(lib._u8p),
                                 C.byref(sched), lib.MOSH2_F64, C.byref(res.c))
        assert rc == 0
        return np.concatenate([res.fullpose, res.trans], axis=1)

    def solve(mine):         # {seq: (obs, vis) tensors on this rank's device} -> {seq: rows tensor}
        import torch
        return {i: torch.from_numpy(solve_one(i, o.numpy(), v.numpy()).astype(np.float32)) for i, (o, v) in mine.items()}

    obs_list = vis_list = None
    if rank == 0:
        obs_list, vis_list = [], []
        for c in cases:
            mc = MocapSession(c['mocap_fname'], 'mm')
            o, v = mc.frames_for_labels(c['latent_labels'], range(len(mc)))
            obs_list.append(o)
            vis_list.append(v)
    out, assignment = shard.solve_sharded([F] * n_seq, [20] * n_seq, [48 + 3] * n_seq, solve, obs_list, vis_list)
    if rank == 0:
        # every sequence solved locally on rank 0 must equal what came back through scatter + gather
        ref = {i: solve_one(i, obs_list[i].astype(np.float32).astype(np.float64), vis_list[i]) for i in range(n_seq)}
        err = max(float(np.abs(out[i].numpy() - ref[i].astype(np.float32)).max()) for i in range(n_seq))
        q.put((sorted(out.keys()), assignment, err))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_solve_gather_world2(tmp_path):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:                          # a free rendezvous port
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    keys, assignment, err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert keys == [0, 1, 2]
    assert sorted(sum(assignment, [])) == [0, 1, 2] and all(len(a) >= 1 for a in assignment)
    assert err < 1e-6
