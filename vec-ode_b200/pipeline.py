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
        self.pool = ThreadPoolExecutor(max_workers=len(self.chunks))

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
