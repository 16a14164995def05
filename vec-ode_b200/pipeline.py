"""Whole solves from host buffers with transfers hidden behind the integration.

Trajectories are independent, so an ensemble can be cut into chunks that are uploaded, integrated and downloaded
independently. Each chunk gets its own `Context` (its own CUDA stream) and its own host thread (`ctypes` releases the GIL
during native calls; the C ABI allows different contexts on different threads): while one chunk integrates on the SMs, the
copy engines upload the next and download the previous one. Results are those of the un-chunked solve bit for bit — the
kernels see the same trajectories, only grouped differently.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .base import Context, Ensemble
from .workloads import shard_range


class ChunkedSolve:
    """`make_solver(ctx, lo, hi, x0)` builds the solver (an `RK45Solver` with its `Rhs`, tolerances, ...) for trajectories
    [lo, hi) of the ensemble on the given context, starting from the (zero-filled) Ensemble `x0` of that shape."""

    def __init__(self, device: int, arith: str, n: int, d: int, make_solver, parts: int = 4):
        self.n, self.d, self.parts = n, d, parts
        self.chunks = []
        for q in range(parts):
            lo, hi = shard_range(n, q, parts)
            if hi <= lo:
                continue
            ctx = Context(device, arith=arith)
            x0 = Ensemble(ctx, d, hi - lo)
            self.chunks.append((lo, hi, ctx, x0, make_solver(ctx, lo, hi, x0)))
        self.pool = ThreadPoolExecutor(max_workers=max(parts, 1))

    @staticmethod
    def _one(chunk, host_in, host_out, adaptive):
        lo, hi, ctx, x0, solver = chunk
        x0.upload(host_in[lo:hi], "aos")
        solver.reset(x0)
        st = solver.run(adaptive=adaptive)
        solver.current()[1].to_host("aos", out=host_out[lo:hi])
        return st

    def solve(self, host_in: np.ndarray, host_out: np.ndarray, adaptive: bool = False):
        """host_in / host_out: [n][d] float64, C-contiguous (pinned memory makes the copies asynchronous to the SMs).
        Returns the per-chunk ODEState list; every chunk is complete (and its stream idle) on return."""
        assert host_in.shape == (self.n, self.d) and host_out.shape == (self.n, self.d)
        futs = [self.pool.submit(self._one, c, host_in, host_out, adaptive) for c in self.chunks]
        return [f.result() for f in futs]

    @property
    def launch_count(self) -> int:
        return sum(c[2].launch_count for c in self.chunks)

    def close(self):
        self.pool.shutdown()


class ShardedChunkedSolve:
    """The same pipeline for ONE ensemble sharded by trajectory over the GPUs of a `Group` (one process per GPU): this rank
    uploads, integrates and hands over its shard in `parts` chunks, and every chunk is gathered to the root as soon as it is
    done — NCCL from the chunk's device state into the root's gather buffer, then one device-to-host copy straight into the
    chunk's place in the root's host array (vo_group_gather_placed) — while the later chunks still integrate. The root's
    host link carries the whole ensemble once; nothing else crosses PCIe on the way back.

    `make_solver(ctx, lo, hi, x0)` as for ChunkedSolve, with [lo, hi) indices into THIS RANK's shard. All ranks issue the
    chunk gathers in the same order (chunk 0, 1, ...) on the group's own stream, which follows each chunk's stream through
    an event (vo_ctx_wait_for)."""

    def __init__(self, group, n_total: int, d: int, make_solver, parts: int = 4, arith: str = "fast"):
        import threading
        self.group, self.n_total, self.d, self.parts = group, n_total, d, parts
        self.gctx = group.ctxs[0]
        self.rank, self.world = group.ranks[0], group.world
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        n_local = self.hi - self.lo
        self.local = ChunkedSolve(self.gctx.device, arith, n_local, d, make_solver, parts=parts)
        # chunk q of rank r: rows and where they go in the whole ensemble
        self.rows, self.off = [], []
        for q in range(parts):
            rows_q, off_q = [], []
            for r in range(self.world):
                slo, shi = shard_range(n_total, r, self.world)
                lo, hi = shard_range(shi - slo, q, parts)
                rows_q.append(max(hi - lo, 0)), off_q.append(slo + min(lo, shi - slo))
            self.rows.append(rows_q), self.off.append(off_q)
        self._turn, self._cv = 0, threading.Condition()
        self._placeholder = None  # a rank whose chunk q is empty still takes part in the gather with a 1-row dummy of 0 rows

    def _one(self, q, chunk, host_in, host_out, adaptive, root):
        st = None
        if chunk is not None:
            lo, hi, ctx, x0, solver = chunk
            x0.upload(host_in[lo:hi], "aos")
            solver.reset(x0)
            st = solver.run(adaptive=adaptive)
        with self._cv:  # the gathers are enqueued in chunk order on every rank
            self._cv.wait_for(lambda: self._turn == q)
            if chunk is not None:
                self.gctx.wait_for(chunk[2])
                ens = chunk[4].current()[1]
            else:
                ens = self._dummy()
            self.group.gather_placed([ens], self.rows[q], self.off[q], host_out, root=root)
            self._turn += 1
            self._cv.notify_all()
        return st

    def _dummy(self):
        if self._placeholder is None:
            self._placeholder = Ensemble(self.gctx, self.d, 1)
        return self._placeholder

    def solve(self, host_in_local: np.ndarray, host_out_full, adaptive: bool = False, root: int = 0):
        """host_in_local: this rank's shard [n_local][d]; host_out_full: [n_total][d] on the root (ignored elsewhere).
        Returns this rank's per-chunk ODEState list; the whole ensemble is in host_out_full on the root on return."""
        self._turn = 0
        by_q = {q: None for q in range(self.parts)}
        for q, c in enumerate(self.local.chunks):
            by_q[q] = c
        assert all(self.rows[q][self.rank] == (0 if by_q[q] is None else by_q[q][1] - by_q[q][0]) for q in range(self.parts))
        pool = self.local.pool
        futs = [pool.submit(self._one, q, by_q[q], host_in_local, host_out_full, adaptive, root) for q in range(self.parts)]
        sts = [f.result() for f in futs]
        self.group.sync()
        return [s for s in sts if s is not None]

    @property
    def launch_count(self) -> int:
        return self.local.launch_count
