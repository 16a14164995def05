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


def chunk_range(n: int, q: int, parts: int):
    """Chunk q of `parts` of n trajectories, sizes tapering towards the end (32 / 28 / 24 / 16 % for four parts): the transfer of the LAST
    chunk is the one nothing overlaps, so it is the smallest; the others still leave the copy engines less to do than the SMs."""
    if parts <= 1:
        return 0, n
    w = np.linspace(1.0 + 0.35, 1.0 - 0.35, parts)
    if parts == 4:
        w = np.array([0.32, 0.28, 0.24, 0.16])
    edges = np.concatenate([[0.0], np.cumsum(w / w.sum())])
    lo, hi = int(round(edges[q] * n)), int(round(edges[q + 1] * n))
    return lo, (n if q == parts - 1 else hi)


class ChunkedSolve:
    """`make_solver(ctx, lo, hi, x0)` builds the solver (an `RK45Solver` with its `Rhs`, tolerances, ...) for trajectories
    [lo, hi) of the ensemble on the given context, starting from the (zero-filled) Ensemble `x0` of that shape."""

    def __init__(self, device: int, arith: str, n: int, d: int, make_solver, parts: int = 4):
        self.n, self.d, self.parts = n, d, parts
        self.chunks = []
        for q in range(parts):
            lo, hi = chunk_range(n, q, parts)
            if hi <= lo:
                continue
            # earlier chunks are more urgent: the chunks then finish one after the other (not all together at the end), so the
            # download of a finished chunk overlaps the integration of the next
            ctx = Context(device, arith=arith, urgency=parts - q)
            x0 = Ensemble(ctx, d, hi - lo)
            self.chunks.append((lo, hi, ctx, x0, make_solver(ctx, lo, hi, x0)))
        self.pool = ThreadPoolExecutor(max_workers=max(parts, 1))

    @staticmethod
    def _one(chunk, host_in, host_out, adaptive):
        lo, hi, ctx, x0, solver = chunk
        x0.upload(host_in[lo:hi], "aos")
        solver.reset(x0)
        st = solver.run(adaptive=adaptive)
        solver.state().to_host("aos", out=host_out[lo:hi])
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
    chunk's place in the root's host array — while the later chunks still integrate. The root's host link carries the whole
    ensemble once; nothing else crosses PCIe on the way back.

    Sharding: contiguous ceil(N/G) ranges (`interleave=False`, vo_group_gather_placed), or round-robin (`interleave=True`:
    rank r holds trajectories r, r + G, ...; vo_group_gather_interleaved) — the static interleave that balances an adaptive
    ensemble whose cost varies along the trajectory index, e.g. config 3's mu sweep.

    `make_solver(ctx, lo, hi, x0)` as for ChunkedSolve, with [lo, hi) indices into THIS RANK's shard. All ranks issue the
    chunk gathers in the same order (chunk 0, 1, ...) on the group's own stream, which follows each chunk's stream through
    an event (vo_ctx_wait_for)."""

    def __init__(self, group, n_total: int, d: int, make_solver, parts: int = 4, arith: str = "fast", interleave: bool = False):
        import threading
        self.group, self.n_total, self.d, self.parts, self.interleave = group, n_total, d, parts, interleave
        self.gctx = group.ctxs[0]
        self.rank, self.world = group.ranks[0], group.world
        G, r = self.world, self.rank
        self.plan = []  # per chunk: (local lo, local hi, gather arguments)
        if interleave:
            self.n_local = max(0, -(-(n_total - r) // G))
            n0 = -(-n_total // G)  # rank 0's count, the largest: chunk boundaries in units of G consecutive trajectories
            for q in range(parts):
                lo, hi = chunk_range(n0, q, parts)
                row0 = min(lo * G, n_total)  # (an empty trailing chunk of a tiny ensemble must still name a row inside the host array)
                tot = max(0, min(hi * G, n_total) - row0)
                mine = max(0, -(-(tot - r) // G))
                self.plan.append((lo, lo + mine, (tot, row0)))
        else:
            slo, shi = shard_range(n_total, r, G)
            self.n_local = shi - slo
            for q in range(parts):
                rows_q, off_q = [], []
                for rr in range(G):
                    a, b = shard_range(n_total, rr, G)
                    lo, hi = chunk_range(b - a, q, parts)
                    rows_q.append(max(hi - lo, 0)), off_q.append(a + min(lo, b - a))
                lo, hi = chunk_range(self.n_local, q, parts)
                self.plan.append((lo, max(lo, hi), (rows_q, off_q)))
        self.chunks = []
        for lo, hi, _ in self.plan:
            if hi <= lo:
                self.chunks.append(None)
                continue
            ctx = Context(self.gctx.device, arith=arith, urgency=parts - len(self.chunks))  # chunk 0 first: its gather overlaps the rest
            x0 = Ensemble(ctx, d, hi - lo)
            self.chunks.append((lo, hi, ctx, x0, make_solver(ctx, lo, hi, x0)))
        self.pool = ThreadPoolExecutor(max_workers=max(parts, 1))
        self._turn, self._cv = 0, threading.Condition()
        self.trace = None  # set to a list to record (chunk, phase, t_begin, t_end) host timestamps of the next solve (tools/e2e_trace.py)
        self._placeholder = None  # a rank whose chunk is empty still takes part in the gather

    def _one(self, q, host_in, host_out, adaptive, root):
        import time
        chunk, st = self.chunks[q], None
        tr = self.trace
        t0 = time.perf_counter()
        if chunk is not None:
            lo, hi, ctx, x0, solver = chunk
            x0.upload(host_in[lo:hi], "aos")
            t1 = time.perf_counter()
            solver.reset(x0)
            st = solver.run(adaptive=adaptive)
            if tr is not None:
                tr.append((q, "upload", t0, t1)), tr.append((q, "run", t1, time.perf_counter()))
        t2 = time.perf_counter()
        with self._cv:  # the gathers are enqueued in chunk order on every rank
            self._cv.wait_for(lambda: self._turn == q)
            if chunk is not None:
                self.gctx.wait_for(chunk[2])
                ens = chunk[4].state()  # no read-back: the gather is enqueued behind the chunk's stream
            else:
                if self._placeholder is None:
                    self._placeholder = Ensemble(self.gctx, self.d, 1)
                ens = self._placeholder
            args = self.plan[q][2]
            if self.interleave:
                self.group.gather_interleaved([ens], args[0], args[1], host_out, self.n_total, root=root)
            else:
                self.group.gather_placed([ens], args[0], args[1], host_out, root=root)
            self._turn += 1
            self._cv.notify_all()
        if tr is not None:
            tr.append((q, "gather_enqueue", t2, time.perf_counter()))
        return st

    def solve(self, host_in_local: np.ndarray, host_out_full, adaptive: bool = False, root: int = 0):
        """host_in_local: this rank's shard [n_local][d]; host_out_full: [n_total][d] on the root (ignored elsewhere).
        Returns this rank's per-chunk ODEState list; the whole ensemble is in host_out_full on the root on return."""
        assert host_in_local.shape == (self.n_local, self.d)
        self._turn = 0
        futs = [self.pool.submit(self._one, q, host_in_local, host_out_full, adaptive, root) for q in range(self.parts)]
        sts = [f.result() for f in futs]
        if self.trace is not None:
            import time
            t0 = time.perf_counter()
            self.group.sync()
            self.trace.append((-1, "final_sync", t0, time.perf_counter()))
        else:
            self.group.sync()
        return [s for s in sts if s is not None]

    @property
    def launch_count(self) -> int:
        return sum(c[2].launch_count for c in self.chunks if c is not None)
