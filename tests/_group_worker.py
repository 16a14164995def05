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
    dist.barrier()
    if rank == 0:
        open(sys.argv[1], "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
