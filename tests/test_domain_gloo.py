"""Row N4 on CPU: a periodic grid state split over world_size-2 gloo processes. Each rank steps its slab (with ghost zones)
using the CPU oracle's heat RHS on the local array — whose periodic wrap-around reads the wrong data at the slab ends,
exactly as the unchanged GPU stage kernels do — and refreshes ghosts with the SAME `PeriodicSlab.exchange` the GPU ranks
run. The gathered owned points must equal the single-process solve bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAPPA, H_STEP = 1.0, 0.2


def _steps(ol, tab, u, n):
    """n calls of step() past the initial Chkpt on one (local or global) array."""
    x, out, _ = ol.rk_solve("HEAT1D", [KAPPA], tab, 0.0, 1.0e9, u, H_STEP, no_adaptive=True, max_calls=n + 1)
    return x


def _worker(rank, world, port, d_total, k, n_steps, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import vecode_b200 as vo
    from oracle import oracle_lib as ol
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = ol.builtin_tableau(1)  # RK4
    slab = vo.domain.PeriodicSlab(d_total, rank, world, k * tab[3])
    local = torch.from_numpy(vo.workloads.heat_u0_at(slab.global_index(), d_total))
    done = 0
    while done < n_steps:
        n = min(k, n_steps - done)
        local = torch.from_numpy(_steps(ol, tab, local.numpy(), n))
        done += n
        if done < n_steps:
            slab.exchange(local)
    full = slab.gather(slab.interior(local.numpy()))
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,d_total,k", [(2, 203, 2), (2, 96, 1), (3, 200, 2)])  # (3, 200): ragged slabs of 67, 67, 66 points
def test_slabs_match_single_process_bitwise(tmp_path, oracle, vo, world, d_total, k):
    n_steps = 7
    port = 31500 + (os.getpid() % 2000) + d_total + world
    mp.spawn(_worker, args=(world, port, d_total, k, n_steps, str(tmp_path)), nprocs=world, join=True)
    ref = _steps(oracle, oracle.builtin_tableau(1), vo.workloads.heat_u0(d_total), n_steps)
    full = np.load(tmp_path / "full.npy")
    assert full.shape == (d_total,) and np.array_equal(full, ref)


def test_slab_bookkeeping_and_single_process_wrap(vo):
    import torch
    slab = vo.domain.PeriodicSlab(10, 0, 1, 3)
    assert slab.m == 10 and slab.local_len == 16 and list(slab.global_index()) == [7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2]
    x = torch.zeros(16, dtype=torch.float64)
    x[3:13] = torch.arange(10, dtype=torch.float64)
    slab.exchange(x)
    assert x.tolist() == [7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2]
    with pytest.raises(ValueError):
        vo.domain.PeriodicSlab(10, 0, 4, 4)  # slabs of 3 points cannot feed 4 ghost points


# ---- adaptive stepping of the distributed state: one global error norm per attempt (domain.py: step_adaptive) ----------------
def rough_u0_at(j, d_total):
    """The smooth initial state of config 4 plus the grid's stiffest mode and a mid-range one: the step size then has to grow
    through a few rejections."""
    import vecode_b200 as vo
    j = np.asarray(j)
    return vo.workloads.heat_u0_at(j, d_total) + 0.25 * np.cos(np.pi * j) + 0.1 * np.sin(2.0 * np.pi * 17.0 * j / d_total)


def _adaptive_worker(rank, world, port, d_total, k, rtol, tf, out_dir):
    """The logic of HeatSlabSolver.step_adaptive with the CPU oracle's rk_step in place of the GPU stage kernels: every rank
    attempts a step on its slab, reduces x_err over its owned points, all-reduces the accumulator, runs handle_step_adaptive
    (ode.rs:311-334) with the global norm; only accepted steps use up ghost points."""
    sys.path.insert(0, ROOT)
    import math
    import torch
    import torch.distributed as dist
    import vecode_b200 as vo
    from oracle import oracle_lib as ol
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = ol.builtin_tableau(0)  # the reference's RKF45 (error estimate on)
    slab = vo.domain.PeriodicSlab(d_total, rank, world, k * tab[3])
    local = torch.from_numpy(rough_u0_at(slab.global_index(), d_total))
    H, m = slab.halo, slab.m
    t, h, prev_h, since, events = 0.0, 0.01, 0.01, 0, []
    alpha, pw, min_dt, max_dt = 0.9, 1.0 / 3.0, 1e-6, 1.0
    first = True
    while True:
        if first:  # ode.rs:144-145: the first call is the Chkpt at t0
            first = False
            events.append(1)
            continue
        rem = tf - t
        if abs(rem) <= 2.220446049250313e-16:
            events.append(3)
            break
        dt = rem if rem < h else h
        if since == k:
            slab.exchange(local)
            since = 0
        xf, xe, _ = ol.rk_step("HEAT1D", [KAPPA], tab, t, dt, local.numpy())
        acc = torch.tensor([float(np.sum(xe[H:H + m] * xe[H:H + m]))], dtype=torch.float64)
        dist.all_reduce(acc)
        dxn = math.sqrt(float(acc.item()))
        f = rtol / dxn
        fp = min(max(alpha * f ** pw, 0.3), 2.0)
        prev_h, h = h, min(max(fp * h, min_dt), max_dt)
        if f <= 1.0:
            events.append(2)
            continue
        local = torch.from_numpy(xf)
        t += dt
        since += 1
        events.append(0)
    full = slab.gather(slab.interior(local.numpy()))
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "events.npy"), np.array(events))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,d_total,k", [(2, 192, 2), (3, 200, 1)])
def test_adaptive_slabs_follow_the_single_process_solve(tmp_path, oracle, vo, world, d_total, k):
    rtol, tf = 1e-6, 3.0
    port = 33500 + (os.getpid() % 2000) + d_total + world
    mp.spawn(_adaptive_worker, args=(world, port, d_total, k, rtol, tf, str(tmp_path)), nprocs=world, join=True)
    rx, ro, trace = oracle.rk_solve("HEAT1D", [KAPPA], oracle.builtin_tableau(0), 0.0, tf, rough_u0_at(np.arange(d_total), d_total), 0.01, adaptive=True, rtol=rtol,
                                    max_dt=1.0, trace_cap=4096)
    events = np.load(tmp_path / "events.npy")
    assert ro.n_reject > 0  # the case exercises rejections
    assert np.array_equal(events, trace[:ro.n_calls, 0].astype(int))  # the same Step / Reject / Chkpt / End sequence, call for call
    np.testing.assert_allclose(np.load(tmp_path / "full.npy"), rx, rtol=0, atol=1e-13)
