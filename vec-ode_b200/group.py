"""Trajectory sharding across the GPUs of one box.

Trajectories are independent (nothing in rk_step, handle_step_adaptive, cfm_general or magnus_42 couples them), so
stepping needs NO collective: each rank integrates the contiguous range `shard_range(N, rank, world)` of the ensemble.
Collectives appear only at the end of a solve: the gather of the final states and the reduction of the counters
(SURVEY.md §8e).

`Group` is the host mirror of the C ABI's `vo_group_*` (NCCL inside libvecode_b200.so, straight from the solvers' device
state: no host staging, one device-to-host copy on the root). It is formed either from one process per GPU
(`Group.from_torch_distributed`: torch.distributed only carries the 128-byte NCCL id to the other ranks) or inside one
process that drives several GPUs (`Group.local`).

The module-level `gather_states` / `reduce_stats` are the host-array versions over `torch.distributed`; they also run on
the gloo backend with CPU tensors, which is how the CPU tests cover world_size 2.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import check, lib
from .workloads import shard_range


class Group:
    """vo_group: the GPUs that share one trajectory-sharded ensemble. `members` = the contexts this process drives
    (one in the process-per-GPU model)."""

    def __init__(self, handle, ctxs):
        self._h, self.ctxs = handle, list(ctxs)
        self.world = int(lib().vo_group_world(handle))
        self.ranks = [int(lib().vo_group_member_rank(handle, i)) for i in range(len(self.ctxs))]

    @classmethod
    def from_torch_distributed(cls, ctx) -> "Group":
        """One process per GPU (torchrun): rank 0 draws the NCCL id, torch.distributed broadcasts its 128 bytes."""
        import torch
        import torch.distributed as dist
        rank, world = rank_world()
        ident = np.zeros(_cabi.GROUP_ID_BYTES, dtype=np.uint8)
        if world > 1:
            if rank == 0:
                check(lib().vo_group_unique_id(ident.ctypes.data_as(C.c_void_p)))
            t = torch.from_numpy(ident)
            if dist.get_backend() == "nccl":
                t = t.cuda(ctx.device)
            dist.broadcast(t, src=0)
            ident = t.cpu().numpy()
        h = C.c_void_p()
        check(lib().vo_group_create_rank(ctx._h, ident.ctypes.data_as(C.c_void_p), rank, world, C.byref(h)), ctx._h)
        return cls(h, [ctx])

    @classmethod
    def local(cls, ctxs) -> "Group":
        """One process, one context per GPU (ncclCommInitAll)."""
        hs = (C.c_void_p * len(ctxs))(*[c._h.value for c in ctxs])
        h = C.c_void_p()
        check(lib().vo_group_create_local(hs, len(ctxs), C.byref(h)), ctxs[0]._h)
        return cls(h, ctxs)

    def _check(self, code):
        if code != _cabi.VO_OK:
            msg = lib().vo_group_last_error(self._h) or b""
            raise _cabi.VecOdeError(code, msg.decode("utf-8", "replace"))

    def shard(self, n_total: int, member: int = 0):
        return shard_range(n_total, self.ranks[member], self.world)

    def _arr(self, objs):
        return (C.c_void_p * len(objs))(*[o._h.value for o in objs])

    def scatter(self, host, n_total: int, d: int, ens_list, root: int = 0, layout: str = "aos"):
        """The root's whole initial ensemble -> every member's shard (`ens_list[i]`: Ensemble of shape (d, shard length))."""
        ptr = None if host is None else np.ascontiguousarray(host, dtype=np.float64).ctypes.data_as(C.c_void_p)
        self._check(lib().vo_group_scatter(self._h, ptr, _cabi.LAYOUT_AOS if layout == "aos" else _cabi.LAYOUT_SOA, d, n_total, root, self._arr(ens_list)))

    def run(self, solvers, adaptive: bool = False, max_calls: int = 0) -> dict:
        st = _cabi.GroupStats()
        self._check(lib().vo_group_run(self._h, self._arr(solvers), 1 if adaptive else 0, max_calls, C.byref(st)))
        return st.as_dict()

    def gather(self, ens_list, n_total: int, root: int = 0, out=None, layout: str = "aos"):
        """Final states of every shard -> `out` on the root ([n_total][d] for 'aos'); returns `out` there, None elsewhere.
        NCCL from device state into one device ensemble on the root, then ONE device-to-host copy."""
        d = ens_list[0].d
        is_root = root in self.ranks
        if is_root and out is None:
            out = np.empty((n_total, d) if layout == "aos" else (d, n_total))
        ptr = out.ctypes.data_as(C.c_void_p) if is_root else None
        self._check(lib().vo_group_gather(self._h, self._arr(ens_list), n_total, root, ptr, _cabi.LAYOUT_AOS if layout == "aos" else _cabi.LAYOUT_SOA))
        return out if is_root else None

    def gather_placed(self, ens_list, rows, row_off, out, root: int = 0, layout: str = "aos"):
        """Asynchronous general gather (vo_group_gather_placed): rank r holds rows[r] trajectories, the root places them at row
        row_off[r] of `out` ([n][d] for 'aos'). Nothing synchronises; call `sync()` before reading `out`."""
        rows_a = np.ascontiguousarray(rows, dtype=np.int64)
        off_a = np.ascontiguousarray(row_off, dtype=np.int64)
        is_root = root in self.ranks
        host_n = 0 if out is None else (out.shape[0] if layout == "aos" else out.shape[1])
        ptr = out.ctypes.data_as(C.c_void_p) if (is_root and out is not None) else None
        if not is_root:
            host_n = int((off_a + rows_a).max())
        self._check(lib().vo_group_gather_placed(self._h, self._arr(ens_list), root, rows_a.ctypes.data_as(C.c_void_p), off_a.ctypes.data_as(C.c_void_p), ptr,
                                                 _cabi.LAYOUT_AOS if layout == "aos" else _cabi.LAYOUT_SOA, host_n))

    def gather_interleaved(self, ens_list, tot: int, row0: int, out, host_n: int, root: int = 0, layout: str = "aos"):
        """Asynchronous gather of a round-robin sharded block (vo_group_gather_interleaved): rank r holds trajectories
        r, r + G, ... of the block's `tot`; the root places the block, in natural order, at rows [row0, row0 + tot) of `out`."""
        ptr = out.ctypes.data_as(C.c_void_p) if (root in self.ranks and out is not None) else None
        self._check(lib().vo_group_gather_interleaved(self._h, self._arr(ens_list), root, tot, row0, ptr,
                                                      _cabi.LAYOUT_AOS if layout == "aos" else _cabi.LAYOUT_SOA, host_n))

    def sync(self):
        self._check(lib().vo_group_sync(self._h))

    def gather_device(self, ens_list, n_total: int, root: int = 0):
        """As `gather`, but the whole ensemble stays on the root's device: returns an Ensemble view (None off the root)."""
        from .base import Ensemble, _TensorOwner
        h = C.c_void_p()
        self._check(lib().vo_group_gather_device(self._h, self._arr(ens_list), n_total, root, C.byref(h)))
        if not h.value:
            return None
        ctx = self.ctxs[self.ranks.index(root)]
        return Ensemble(ctx, ens_list[0].d, n_total, _handle=h, _owner=_TensorOwner(self))

    def reduce_stats(self, solvers) -> dict:
        st = _cabi.GroupStats()
        self._check(lib().vo_group_reduce_stats(self._h, self._arr(solvers), C.byref(st)))
        return st.as_dict()

    def allreduce(self, values, op: str = "sum") -> np.ndarray:
        """values: [members][n] doubles (n <= 32), reduced over ALL ranks in place."""
        a = np.ascontiguousarray(values, dtype=np.float64).reshape(len(self.ctxs), -1).copy()
        self._check(lib().vo_group_allreduce(self._h, a.ctypes.data_as(C.c_void_p), a.shape[1], {"sum": 0, "max": 1, "min": 2}[op]))
        return a

    def __del__(self):
        try:
            if self._h and all(c._h for c in self.ctxs):
                lib().vo_group_destroy(self._h)
                self._h = None
        except Exception:
            pass


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
    if device is None and dist.get_backend() == "nccl":  # NCCL moves device memory: stage on this process's current GPU
        device = torch.device("cuda", torch.cuda.current_device())
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
    """all_reduce of the per-rank counters: sums of accepted / rejected, min / max of t, and the UNION of the VO_TRAJ_* status
    bits (one 0/1 flag per bit reduced with MAX: a MAX over the bitmasks themselves would lose bits)."""
    import torch
    import torch.distributed as dist
    acc, rej = int(np.sum(stats["accepted"])), int(np.sum(stats["rejected"]))
    tmin, tmax = float(np.min(stats["t"])), float(np.max(stats["t"]))
    status = int(np.bitwise_or.reduce(stats["status"])) if "status" in stats else 0
    if not is_distributed():
        return dict(accepted=acc, rejected=rej, t_min=tmin, t_max=tmax, status=status)
    sums = torch.tensor([acc, rej], dtype=torch.int64)
    bits = [float((status >> b) & 1) for b in range(8)]
    mx = torch.tensor([tmax, -tmin] + bits, dtype=torch.float64)
    if device is not None:
        sums, mx = sums.to(device), mx.to(device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    union = sum((1 << b) for b in range(8) if float(mx[2 + b]) > 0.5)
    return dict(accepted=int(sums[0]), rejected=int(sums[1]), t_min=-float(mx[1]), t_max=float(mx[0]), status=union)
