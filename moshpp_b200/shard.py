"""Sequence sharding over the GPUs of one box.

Sequences are independent (nothing is shared but the static model), so the path shards with no
data-path collective: every rank solves its own sequences.  ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests) only carries the trivial scatter of observations from rank 0 and the gather of per-frame
results back to it.  One process per GPU; launch with torchrun.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def assign_sequences(frame_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of sequences to ranks (balanced by frame count)."""
    order = sorted(range(len(frame_counts)), key=lambda i: (-frame_counts[i], i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += frame_counts[i]
    for r in range(world_size):
        out[r].sort()
    return out


def _device(dist):
    import torch
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def scatter_observations(obs_list: Optional[List[np.ndarray]], vis_list: Optional[List[np.ndarray]],
                         assignment: List[List[int]], shapes: List[tuple], src: int = 0):
    """Rank ``src`` holds every sequence's (obs F x M x 3, vis F x M); each rank receives the ones assigned
    to it.  ``shapes[i] = (F_i, M_i)`` must be known on every rank (it is derived from the job list)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device(dist)
    mine: Dict[int, tuple] = {}
    reqs = []
    if rank == src:
        for r in range(world):
            for i in assignment[r]:
                o = torch.from_numpy(np.ascontiguousarray(obs_list[i], dtype=np.float32))
                v = torch.from_numpy(np.ascontiguousarray(vis_list[i], dtype=np.uint8))
                if r == src:
                    mine[i] = (o.numpy().astype(np.float64), v.numpy().astype(bool))
                else:
                    reqs.append(dist.isend(o.to(dev), dst=r))
                    reqs.append(dist.isend(v.to(dev), dst=r))
    else:
        for i in assignment[rank]:
            F, M = shapes[i]
            o = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
            v = torch.empty((F, M), dtype=torch.uint8, device=dev)
            dist.recv(o, src=src)
            dist.recv(v, src=src)
            mine[i] = (o.cpu().numpy().astype(np.float64), v.cpu().numpy().astype(bool))
    for q in reqs:
        q.wait()
    return mine


def gather_results(local: Dict[int, Dict[str, np.ndarray]], assignment: List[List[int]],
                   row_widths: List[int], shapes: List[tuple], dst: int = 0):
    """Gathers per-sequence result rows (one float32 matrix F_i x row_widths[i] per sequence, e.g.
    [fullpose | trans | dmpls]) to rank ``dst``.  Returns {seq: matrix} on dst, {} elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device(dist)
    out: Dict[int, np.ndarray] = {}
    reqs = []
    if rank == dst:
        for r in range(world):
            for i in assignment[r]:
                if r == dst:
                    out[i] = np.asarray(local[i], dtype=np.float32)
                else:
                    buf = torch.empty((shapes[i][0], row_widths[i]), dtype=torch.float32, device=dev)
                    dist.recv(buf, src=r)
                    out[i] = buf.cpu().numpy()
    else:
        for i in assignment[rank]:
            t = torch.from_numpy(np.ascontiguousarray(local[i], dtype=np.float32)).to(dev)
            reqs.append(dist.isend(t, dst=dst))
    for q in reqs:
        q.wait()
    return out


def solve_sharded(frame_counts: Sequence[int], n_markers: Sequence[int], row_widths: Sequence[int],
                  solve_fn: Callable[[int, np.ndarray, np.ndarray], np.ndarray],
                  obs_list: Optional[List[np.ndarray]] = None, vis_list: Optional[List[np.ndarray]] = None):
    """Scatter -> per-rank solves -> gather.  ``solve_fn(seq_index, obs, vis)`` returns the F x width
    result matrix of one sequence (on a GPU box it wraps ``lib.Model.solve`` on the rank's device)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    assignment = assign_sequences(list(frame_counts), world)
    shapes = [(int(f), int(m)) for f, m in zip(frame_counts, n_markers)]
    mine = scatter_observations(obs_list, vis_list, assignment, shapes)
    local = {i: solve_fn(i, o, v) for i, (o, v) in mine.items()}
    return gather_results(local, assignment, list(row_widths), shapes), assignment
