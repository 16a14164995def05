"""Trajectory sharding across the GPUs of one box: one process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).

Trajectories are independent (nothing in rk_step, handle_step_adaptive, cfm_general or magnus_42 couples them), so
stepping needs NO collective: each rank integrates the contiguous range `shard_range(N, rank, world)` of the ensemble.
Collectives appear only at the end of a solve: one all_gather of the final states and one all_reduce of the counters
(SURVEY.md §8e). The same code runs on the gloo backend with CPU tensors, which is how the tests cover world_size 2.
"""
from __future__ import annotations

import numpy as np

from .workloads import shard_range


def is_distributed() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def rank_world():
    import torch.distributed as dist
    return (dist.get_rank(), dist.get_world_size()) if is_distributed() else (0, 1)


def my_range(n_total: int):
    r, w = rank_world()
    return shard_range(n_total, r, w)


def gather_states(local, n_total: int, device=None, root=None):
    """Gather of per-rank final states [n_local][d] into the whole ensemble [n_total][d]. Ranks may hold ragged shard sizes
    (ceil(N/G) ranges): shards are padded to the largest one for the collective. `root=None`: all_gather, every rank gets the
    ensemble. `root=r`: only rank r receives it (the others return None) — one device-to-host copy of the whole ensemble
    instead of one per rank, which is what dominates an 8-GPU gather."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not is_distributed():
        return local
    r, w = rank_world()
    per = -(-n_total // w)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    as_real = pad.view(np.float64) if np.iscomplexobj(pad) else pad
    t = torch.from_numpy(as_real.copy())
    if device is not None:
        t = t.to(device)
    if root is None:
        out = [torch.empty_like(t) for _ in range(w)]
        dist.all_gather(out, t)
    else:
        out = [torch.empty_like(t) for _ in range(w)] if r == root else None
        dist.gather(t, out, dst=root)
        if r != root:
            return None
    whole = torch.stack(out).cpu().numpy()  # one device-to-host copy
    parts = []
    for q in range(w):
        lo, hi = shard_range(n_total, q, w)
        a = whole[q]
        if np.iscomplexobj(local):
            a = a.view(np.complex128)
        parts.append(a[: hi - lo])
    return np.concatenate(parts, axis=0)


def reduce_stats(stats: dict, device=None) -> dict:
    """all_reduce of the per-rank counters: sums of accepted / rejected, max of t and of the status bits, min of t."""
    import torch
    import torch.distributed as dist
    acc, rej = int(np.sum(stats["accepted"])), int(np.sum(stats["rejected"]))
    tmin, tmax = float(np.min(stats["t"])), float(np.max(stats["t"]))
    status = int(np.bitwise_or.reduce(stats["status"])) if "status" in stats else 0
    if not is_distributed():
        return dict(accepted=acc, rejected=rej, t_min=tmin, t_max=tmax, status=status)
    sums = torch.tensor([acc, rej], dtype=torch.int64)
    mx = torch.tensor([tmax, -tmin, float(status)], dtype=torch.float64)
    if device is not None:
        sums, mx = sums.to(device), mx.to(device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return dict(accepted=int(sums[0]), rejected=int(sums[1]), t_min=-float(mx[1]), t_max=float(mx[0]), status=int(mx[2]))
