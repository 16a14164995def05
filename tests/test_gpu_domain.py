"""Row N4 on the GPU: the heat state split into slabs with ghost zones runs the unchanged single-GPU stage kernels and must
reproduce the single-GPU solve bit for bit on the owned points (vec-ode_b200/domain.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_STEPS, H_STEP = 9, 0.2


def _reference(vo, ctx, d_total):
    rhs = vo.Rhs(ctx, "HEAT1D", d_total, [1.0])
    s = vo.RK45Solver(rhs, 0.0, 1.0e9, vo.Ensemble.from_host(ctx, vo.workloads.heat_u0(d_total)[None, :]), H_STEP, tableau=vo.ButcherTableu.builtin("RK4"))
    s.no_adaptive()
    for _ in range(N_STEPS + 1):  # the first call is the Chkpt at t0
        s.step()
    return s.current()[1].to_host()[0]


@pytest.mark.parametrize("fused", [False, True])  # stage path / whole-step kernel on the slab
@pytest.mark.parametrize("d_total,k", [(4099, 1), (1 << 16, 4), ((1 << 18) + 2, 3)])  # plain stage kernel / TMA-staged stage kernel
def test_single_rank_slab_with_periodic_ghosts_bitwise(vo, ctx, d_total, k, fused):
    ref = _reference(vo, ctx, d_total)
    ds = vo.domain.HeatSlabSolver(ctx, d_total, lambda j: vo.workloads.heat_u0_at(j, d_total), 1.0, 0.0, 1.0e9, H_STEP, steps_per_exchange=k, fused=fused)
    for _ in range(N_STEPS + 1):
        ds.step()
    assert ds.exchanges == (N_STEPS - 1) // k
    assert np.array_equal(ds.local_interior(), ref)


def _worker(rank, world, port, d_total, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import vecode_b200 as vo
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # both ranks share cuda:0 here; NCCL needs one GPU per rank
    ctx = vo.Context(0, arith="strict")
    ds = vo.domain.HeatSlabSolver(ctx, d_total, lambda j: vo.workloads.heat_u0_at(j, d_total), 1.0, 0.0, 1.0e9, H_STEP, steps_per_exchange=k)
    for _ in range(N_STEPS + 1):
        ds.step()
    full = ds.gather()
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "ref.npy"), _reference(vo, ctx, d_total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("d_total,k", [((1 << 17) + 6, 2)])
def test_two_ranks_on_one_gpu_bitwise(tmp_path, d_total, k):
    port = 32500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, d_total, k, str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "full.npy"), np.load(tmp_path / "ref.npy"))


# ---- adaptive stepping of the distributed state (one global error norm per attempt: vo_adaptive_try / vo_adaptive_handle) ---
def _rough_u0_at(vo, j, d_total):
    j = np.asarray(j)
    return vo.workloads.heat_u0_at(j, d_total) + 0.25 * np.cos(np.pi * j) + 0.1 * np.sin(2.0 * np.pi * 17.0 * j / d_total)


def _adaptive_reference(vo, ctx, d_total, tf, rtol):
    """The single-GPU adaptive solve (tree-reduced norm, controller on the host) and its event sequence, call by call."""
    rhs = vo.Rhs(ctx, "HEAT1D", d_total, [1.0])
    s = vo.RK45Solver(rhs, 0.0, tf, vo.Ensemble.from_host(ctx, _rough_u0_at(vo, np.arange(d_total), d_total)[None, :]), 0.01).with_tolerance(rtol, rtol)
    s.with_step_range(1e-6, 1.0).with_init_step(0.01)
    events = []
    while True:
        st = s.step_adaptive()
        events.append([k for k in ("Step", "Chkpt", "Reject", "End") if st.counts[k]][0])
        if st.kind != "Ok":
            break
    return s.current()[1].to_host()[0], events, s.stats()


def _adaptive_slab_run(vo, ctx, d_total, tf, rtol, k, fused=False):
    ds = vo.domain.HeatSlabSolver(ctx, d_total, lambda j: _rough_u0_at(vo, j, d_total), 1.0, 0.0, tf, 0.01, tableau=vo.ButcherTableu.builtin("RKF45_REF"),
                                  steps_per_exchange=k, fused=fused, adaptive=True)
    ds.with_tolerance(rtol, rtol)
    ds.solver.with_step_range(1e-6, 1.0).with_init_step(0.01)
    events = []
    while True:
        st = ds.step_adaptive()
        events.append([kk for kk in ("Step", "Chkpt", "Reject", "End") if st.counts[kk]][0])
        if st.kind != "Ok":
            break
    return ds, events


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("d_total,k", [(4099, 1), (1 << 16, 3)])
def test_adaptive_single_rank_slab_follows_the_single_state_solve(vo, ctx, d_total, k, fused):
    """ode.rs:311-344 on a slab with periodic ghosts: same Step / Reject sequence as the single-state solve; the states agree to
    rounding (the two norms sum the same squares in a different order, so h differs in the last bits)."""
    tf, rtol = 3.0, 1e-6
    ref, ref_events, stats = _adaptive_reference(vo, ctx, d_total, tf, rtol)
    ds, events = _adaptive_slab_run(vo, ctx, d_total, tf, rtol, k, fused)
    assert stats["rejected"][0] > 0 and events == ref_events
    assert np.abs(ds.local_interior() - ref).max() <= 1e-12
    with pytest.raises(vo.VecOdeError):  # a second try without handling the first
        s2 = vo.domain.HeatSlabSolver(ctx, 4096, lambda j: _rough_u0_at(vo, j, 4096), 1.0, 0.0, 1.0, 0.01, tableau=vo.ButcherTableu.builtin("RKF45_REF"), adaptive=True)
        s2.solver.adaptive_try(0, 10), s2.solver.adaptive_try(0, 10), s2.solver.adaptive_try(0, 10)


def _adaptive_worker(rank, world, port, d_total, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import vecode_b200 as vo
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = vo.Context(0, arith="strict")
    ds, events = _adaptive_slab_run(vo, ctx, d_total, 3.0, 1e-6, k)
    full = ds.gather()
    if rank == 0:
        ref, ref_events, _ = _adaptive_reference(vo, ctx, d_total, 3.0, 1e-6)
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "ref.npy"), ref)
        np.save(os.path.join(out_dir, "same_events.npy"), np.array([events == ref_events and "Reject" in events]))
    dist.barrier()
    dist.destroy_process_group()


def test_adaptive_two_ranks_on_one_gpu(tmp_path):
    port = 34500 + (os.getpid() % 2000)
    mp.spawn(_adaptive_worker, args=(2, port, (1 << 15) + 6, 2, str(tmp_path)), nprocs=2, join=True)
    assert bool(np.load(tmp_path / "same_events.npy")[0])
    assert np.abs(np.load(tmp_path / "full.npy") - np.load(tmp_path / "ref.npy")).max() <= 1e-12


def test_slab_with_a_user_stencil_of_radius_two_bitwise(vo, ctx):
    """A user stencil (vo_rhs_create_custom_stencil) in place of the compiled-in heat equation: ghost zones of k * s * R points, the
    owned points of the slab equal the single-state solve bit for bit (position-independent body: a slab sees local indices)."""
    d_total, k, R = (1 << 16) + 4, 2, 2
    body = "du = (-u[0] + 16.0 * u[1] - 30.0 * u[2] + 16.0 * u[3] - u[4]) * (p[0] / 12.0);"
    u0 = vo.workloads.heat_u0(d_total)
    s = vo.RK45Solver(vo.Rhs.custom_stencil(ctx, body, d_total, R, [0.4]), 0.0, 1.0e9, vo.Ensemble.from_host(ctx, u0[None, :]), H_STEP,
                      tableau=vo.ButcherTableu.builtin("RK4")).no_adaptive()
    for _ in range(N_STEPS + 1):
        s.step()
    ref = s.current()[1].to_host()[0]
    ds = vo.domain.HeatSlabSolver(ctx, d_total, lambda j: vo.workloads.heat_u0_at(j, d_total), 1.0, 0.0, 1.0e9, H_STEP, steps_per_exchange=k,
                                  rhs_factory=lambda c, n: vo.Rhs.custom_stencil(c, body, n, R, [0.4]), radius=R)
    assert ds.slab.halo == k * 4 * R
    for _ in range(N_STEPS + 1):
        ds.step()
    assert np.array_equal(ds.local_interior(), ref)


def test_adaptive_slab_with_a_user_norm(vo, ctx):
    """The user's norm on a distributed state: every piece reports the accumulator of the functor's `map` over its own points
    (vo_adaptive_try with VO_NORM_CUSTOM numbers the components globally within the slab), the pieces are joined and the host applies
    the functor's `finish` twin — same Step / Reject sequence as the single-state solve with the same norm."""
    d_total, tf, rtol = 1 << 14, 2.0, 1e-6
    mk = lambda: vo.NormFn(ctx, "m = e * e;", "sum", "r = sqrt(acc / n);", finish_py=lambda acc, n: float(np.sqrt(acc / n)))
    rhs = vo.Rhs(ctx, "HEAT1D", d_total, [1.0])
    s = vo.RK45Solver(rhs, 0.0, tf, vo.Ensemble.from_host(ctx, _rough_u0_at(vo, np.arange(d_total), d_total)[None, :]), 0.01).with_tolerance(rtol, rtol)
    s.with_step_range(1e-6, 1.0).with_init_step(0.01).with_norm(mk())
    ref_events = []
    while True:
        st = s.step_adaptive()
        ref_events.append([k for k in ("Step", "Chkpt", "Reject", "End") if st.counts[k]][0])
        if st.kind != "Ok":
            break
    ds = vo.domain.HeatSlabSolver(ctx, d_total, lambda j: _rough_u0_at(vo, j, d_total), 1.0, 0.0, tf, 0.01, tableau=vo.ButcherTableu.builtin("RKF45_REF"),
                                  steps_per_exchange=2, adaptive=True)
    ds.with_tolerance(rtol, rtol, norm=mk())
    ds.solver.with_step_range(1e-6, 1.0).with_init_step(0.01)
    events = []
    while True:
        st = ds.step_adaptive()
        events.append([k for k in ("Step", "Chkpt", "Reject", "End") if st.counts[k]][0])
        if st.kind != "Ok":
            break
    assert events == ref_events
    assert np.abs(ds.local_interior() - s.current()[1].to_host()[0]).max() <= 1e-12
