"""Worker of tests/test_gpu_group.py::test_two_process_group_under_torchrun (one process per GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import vecode_b200 as vo
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = vo.Context(local, arith="strict")
    g = vo.group.Group.from_torch_distributed(ctx)
    assert g.world == world and g.ranks == [rank]
    n, tf = 40_003, 1.0
    lo, hi = g.shard(n)
    mu = vo.workloads.vdp_mu(n, hi - lo, lo)

    def make(c, a, b, x0):
        return vo.RK45Solver(vo.Rhs(c, "VDP", 2, [mu[a:b].copy()]), 0.0, tf, x0, 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    s = make(ctx, 0, hi - lo, vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(hi - lo)))
    tot = g.run([s], adaptive=True)
    got = g.gather([s.current()[1]], n, root=0)
    # the pipelined form: chunks gathered one by one into their place
    sh = vo.pipeline.ShardedChunkedSolve(g, n, 2, make, parts=3, arith="strict")
    full = np.zeros((n, 2)) if rank == 0 else None
    sh.solve(vo.workloads.vdp_x0(hi - lo), full, adaptive=True)
    # round-robin sharding (rank r holds trajectories r, r + G, ...), gathered back into natural order
    mu_all = vo.workloads.vdp_mu(n)
    mu_il = mu_all[rank::world].copy()

    def make_il(c, a, b, x0):
        return vo.RK45Solver(vo.Rhs(c, "VDP", 2, [mu_il[a:b].copy()]), 0.0, tf, x0, 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    shi = vo.pipeline.ShardedChunkedSolve(g, n, 2, make_il, parts=3, arith="strict", interleave=True)
    full_il = np.zeros((n, 2)) if rank == 0 else None
    shi.solve(vo.workloads.vdp_x0(len(mu_il)), full_il, adaptive=True)
    if rank == 0:
        ref = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu_all]), 0.0, tf, vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(n)), 1e-3,
                            tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
        ref.run(adaptive=True)
        rx, rs = ref.current()[1].to_host(), ref.stats()
        assert np.array_equal(got, rx), "gathered ensemble differs from the single-GPU solve"
        assert np.array_equal(full, rx), "chunk-wise gathered ensemble differs from the single-GPU solve"
        assert np.array_equal(full_il, rx), "round-robin sharded ensemble differs from the single-GPU solve"
        assert tot["accepted"] == int(rs["accepted"].sum()) and tot["rejected"] == int(rs["rejected"].sum()) and tot["n_done"] == n
    else:
        assert got is None
    # ---- one grid state over the two GPUs, ADAPTIVE, with the library's own all-reduce between the two halves of step_adaptive
    # (vo_adaptive_try -> vo_group_allreduce -> vo_adaptive_handle: what a compiled host without torch.distributed does)
    import math
    d_total, k_ex = (1 << 15) + 6, 2
    tab = vo.ButcherTableu.builtin("RKF45_REF")

    def rough(j):
        j = np.asarray(j)
        return vo.workloads.heat_u0_at(j, d_total) + 0.25 * np.cos(np.pi * j) + 0.1 * np.sin(2.0 * np.pi * 17.0 * j / d_total)
    ds = vo.domain.HeatSlabSolver(ctx, d_total, rough, 1.0, 0.0, 3.0, 0.01, tableau=tab, steps_per_exchange=k_ex, adaptive=True)
    ds.with_tolerance(1e-6, 1e-6)
    ds.solver.with_step_range(1e-6, 1.0).with_init_step(0.01)
    events, H, m = [], ds.slab.halo, ds.slab.m
    while True:
        ds._refresh_ghosts_if_due()
        acc, ev, done = ds.solver.adaptive_try(H, H + m)
        if done is None:
            tot_acc = float(g.allreduce([[acc]], "sum")[0, 0])  # NCCL inside libvecode_b200.so
            st = ds.solver.adaptive_handle(math.sqrt(tot_acc))
            if st.counts["Step"]:
                ds._since_exchange += 1
        else:
            st = done
        events.append([kk for kk in ("Step", "Chkpt", "Reject", "End") if st.counts[kk]][0])
        if st.kind != "Ok":
            break
    full_grid = ds.gather()
    if rank == 0:
        rs_ = vo.RK45Solver(vo.Rhs(ctx, "HEAT1D", d_total, [1.0]), 0.0, 3.0, vo.Ensemble.from_host(ctx, rough(np.arange(d_total))[None, :]), 0.01).with_tolerance(1e-6, 1e-6)
        rs_.with_step_range(1e-6, 1.0).with_init_step(0.01)
        ref_events = []
        while True:
            st = rs_.step_adaptive()
            ref_events.append([kk for kk in ("Step", "Chkpt", "Reject", "End") if st.counts[kk]][0])
            if st.kind != "Ok":
                break
        assert events == ref_events and "Reject" in events, "adaptive slabs: the event sequence differs from the single-state solve"
        assert np.abs(full_grid - rs_.current()[1].to_host()[0]).max() <= 1e-12
    dist.barrier()
    if rank == 0:
        open(sys.argv[1], "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
