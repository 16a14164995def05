"""One large state across several GPUs: a periodic 1-D grid state `V = Array1<f64>` of the reference
(`src/impls/ndarray.rs:8-33`) cut into contiguous slabs, one process per GPU.

Design: ghost zones wide enough that the UNCHANGED single-GPU stage kernels run on every slab. Each rank holds its m
points plus H ghost points per side, copies of its periodic neighbours' edge points. An s-stage explicit RK step with a
3-point stencil reads one more neighbour per stage, so one step corrupts at most s points at each end of the local array
(where the kernels' periodic wrap-around reads the wrong data) and leaves every other point with exactly the bits the
single-GPU solve produces: same kernels, same operations, same order. With H = k*s the ranks exchange ghost points once
every k steps — the only communication of the fixed-step path: one all-gather of 2H doubles per rank (NCCL over NVLink on
GPUs, gloo in the CPU tests).

Adaptive stepping (`step_adaptive`, ode.rs:311-344) needs ONE error norm of the whole state per attempt: every rank runs
try_step on its slab and reduces x_err over its OWNED points only (vo_adaptive_try), the accumulators are combined with one
all-reduce of a single double (sum of squares for the L2 norm, sum for L1, max for Linf), and every rank hands the same global
norm to the controller (vo_adaptive_handle), so (t, h) and every accept / reject decision are identical on all ranks. A
rejected attempt leaves the state untouched, so only ACCEPTED steps use up ghost points.
"""
from __future__ import annotations

import numpy as np

from .group import is_distributed, rank_world
from .workloads import shard_range


class PeriodicSlab:
    """Index bookkeeping of one rank's slab of a periodic grid of `d_total` points: `m` owned points [lo, hi), `halo` ghost
    points on each side; local index j <-> global index (lo - halo + j) mod d_total."""

    def __init__(self, d_total: int, rank: int, world: int, halo: int):
        self.d_total, self.rank, self.world, self.halo = d_total, rank, world, halo
        self.lo, self.hi = shard_range(d_total, rank, world)
        self.m = self.hi - self.lo
        if self.m < halo:
            raise ValueError(f"slab of {self.m} points is narrower than its ghost zone ({halo})")
        self.local_len = self.m + 2 * halo
        self.left, self.right = (rank - 1) % world, (rank + 1) % world

    def global_index(self) -> np.ndarray:
        return (np.arange(self.lo - self.halo, self.hi + self.halo, dtype=np.int64)) % self.d_total

    def scatter(self, u_global: np.ndarray) -> np.ndarray:
        """The local array (ghosts filled) of a state given on the whole grid."""
        return np.ascontiguousarray(u_global[self.global_index()])

    def interior(self, local):
        return local[self.halo:self.halo + self.m]

    def exchange(self, x):
        """Refresh the ghost points of the 1-D tensor `x` (torch, CPU or CUDA, length local_len) from the neighbours' edges."""
        import torch
        H, m = self.halo, self.m
        edges = torch.cat([x[H:2 * H], x[m:m + H]])  # my first H and my last H owned points
        if self.world == 1 or not is_distributed():
            x[:H] = edges[H:]
            x[H + m:] = edges[:H]
            return
        import torch.distributed as dist
        staged = dist.get_backend() == "gloo" and edges.is_cuda  # gloo moves host memory
        send = edges.cpu() if staged else edges.contiguous()
        out = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(out, send)
        lg, rg = out[self.left][H:], out[self.right][:H]  # left neighbour's last H points, right neighbour's first H
        x[:H] = lg.to(x.device) if staged else lg
        x[H + m:] = rg.to(x.device) if staged else rg

    def gather(self, local_interior: np.ndarray) -> np.ndarray:
        """The whole grid state on every rank from the owned points of each."""
        from .group import gather_states
        return gather_states(np.ascontiguousarray(local_interior)[:, None], self.d_total)[:, 0]


class _DeviceView:
    """__cuda_array_interface__ over a raw device pointer, so torch can address a solver's state buffer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class HeatSlabSolver:
    """Fixed-step RK on the semi-discretised periodic heat equation (config 4's RHS) with the grid split over the ranks of
    the process group. Same stepping semantics as `RK45Solver.step()` for a single large state (N = 1, lock-step control on
    the host); every rank runs the same (t, dt) sequence, so the only data-path exchange is the ghost refresh."""

    def __init__(self, ctx, d_total: int, u0_global_fn, kappa: float, t0: float, tf: float, h: float, tableau=None, steps_per_exchange: int = 4, fused: bool = False, adaptive: bool = False,
                 rhs_factory=None, radius: int = 1):
        from . import base
        rank, world = rank_world()
        self.ctx = ctx
        self.tableau = tableau or base.ButcherTableu.builtin("RK4")
        self.k = int(steps_per_exchange)
        # a stencil of radius R corrupts R points per stage at each slab end: ghost zone = k * s * R
        self.slab = PeriodicSlab(d_total, rank, world, self.k * self.tableau.num_stages() * radius)
        u0 = u0_global_fn(self.slab.global_index())  # every rank evaluates its own points (and ghosts) of the initial state
        # rhs_factory(ctx, local_len): a user stencil (base.Rhs.custom_stencil) in place of the compiled-in heat equation
        self.rhs = rhs_factory(ctx, self.slab.local_len) if rhs_factory else base.Rhs(ctx, "HEAT1D", self.slab.local_len, [kappa])
        self.solver = base.RK45Solver(self.rhs, t0, tf, base.Ensemble.from_host(ctx, u0[None, :]), h, tableau=self.tableau)
        if not adaptive:
            self.solver.no_adaptive()
        self._norm_kind, self._adaptive = "L2", adaptive
        if fused:  # whole-step kernel (rk_heat_fused.cuh): its periodic wrap corrupts the same s points per step at the slab ends
            self.solver.set_fused_step()
        self._fused = fused
        self._views = {}
        self._since_exchange = 0  # ghosts are fresh at construction
        self.exchanges = 0

    def _state_tensor(self):
        import torch
        ens = self.solver.current()[1]
        ptr = ens.device_ptr
        if ptr not in self._views:  # x and next_x swap on every accepted step (ode.rs:184-188): two buffers alternate
            self._views[ptr] = torch.as_tensor(_DeviceView(ptr, self.slab.local_len), device=f"cuda:{self.ctx.device}")
        return self._views[ptr]

    def step(self):
        """One call of the reference's `step()` for the distributed state."""
        # the exchange is torch work on torch's current stream; a ctx with a stream of its own is fenced on both sides
        self._refresh_ghosts_if_due()
        st = self.solver.step()
        if st.counts["Step"]:
            self._since_exchange += 1
        return st

    def _refresh_ghosts_if_due(self):
        if self._since_exchange == self.k:
            import torch
            tstream = torch.cuda.current_stream(self.ctx.device)
            foreign = self.ctx.stream != (tstream.cuda_stream or 1)
            if foreign:
                self.ctx.sync()
            self.slab.exchange(self._state_tensor())
            if foreign:
                tstream.synchronize()
            self._since_exchange, self.exchanges = 0, self.exchanges + 1

    def step_adaptive(self):
        """One call of the reference's `step_adaptive()` (ode.rs:336-344) for the distributed state: the error norm is taken over
        the whole grid (one all-reduce of one double per attempt), the controller runs identically on every rank."""
        import math
        self._refresh_ghosts_if_due()
        H, m = self.slab.halo, self.slab.m
        acc, ev, done = self.solver.adaptive_try(H, H + m)
        if done is not None:  # Chkpt / End: no norm, nothing to combine
            return done
        kind = self._norm_kind
        custom = not isinstance(kind, str)  # a base.NormFn: its own join and finish
        use_max = kind.join == "max" if custom else kind == "LINF"
        if self.slab.world > 1 and is_distributed():
            import torch
            import torch.distributed as dist
            t = torch.tensor([acc], dtype=torch.float64, device=("cpu" if dist.get_backend() == "gloo" else f"cuda:{self.ctx.device}"))
            dist.all_reduce(t, op=dist.ReduceOp.MAX if use_max else dist.ReduceOp.SUM)
            acc = float(t.item())
        st = self.solver.adaptive_handle(kind.finish_host(acc, self.slab.d_total) if custom else (math.sqrt(acc) if kind == "L2" else acc))
        if st.counts["Step"]:
            self._since_exchange += 1
        return st

    def with_tolerance(self, atol: float, rtol: float, norm: str = "L2"):
        """with_tolerance (ode.rs:296-306) + the norm of the whole state: "L2", "L1", "LINF", or a base.NormFn built with `finish_py`
        (the host-side twin of its finish statements, applied to the all-reduced accumulator)."""
        self.solver.with_tolerance(atol, rtol)
        self.solver.with_norm(norm)
        self._norm_kind = norm
        self._adaptive = True
        return self

    def run(self, max_calls: int = 0, adaptive: bool = False):
        st, calls = None, 0
        while st is None or (st.kind == "Ok" and (max_calls <= 0 or calls < max_calls)):
            st = self.step_adaptive() if adaptive else self.step()
            calls += 1
        return st

    def reset(self, x0):
        """Restart from the local state `x0` (an Ensemble of the slab's length, ghosts included and fresh)."""
        self.solver.reset(x0)
        if not self._adaptive:
            self.solver.no_adaptive()
        if self._fused:
            self.solver.set_fused_step()
        self._since_exchange = 0

    def local_interior(self) -> np.ndarray:
        return self.slab.interior(self.solver.current()[1].to_host()[0])

    def gather(self) -> np.ndarray:
        return self.slab.gather(self.local_interior())
