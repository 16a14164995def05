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
