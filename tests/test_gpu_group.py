"""vo_group_* (NCCL inside libvecode_b200.so): one ensemble sharded by trajectory over the GPUs of a box, the final gather from
device state and the reduction of the counters. A world of one runs on any GPU box; the two-rank cases (one process driving
two GPUs, and two processes under torchrun) need two GPUs and are skipped otherwise. Every case ends bit-identical to the
single-GPU solve of the whole ensemble (trajectories are independent; parameters are indexed by global trajectory number)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _whole(vo, ctx, n, tf):
    mu = vo.workloads.vdp_mu(n)
    s = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu]), 0.0, tf, vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(n)), 1e-3,
                      tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    s.run(adaptive=True)
    return s.current()[1].to_host(), s.stats()


@pytest.mark.parametrize("n_dev", [1, 2])
def test_local_group_scatter_run_gather_reduce(vo, n_dev):
    """One process, one host thread, n_dev GPUs (vo_group_create_local = ncclCommInitAll)."""
    if _n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    n, tf = 5001, 1.0  # ragged shards, odd sizes
    ctxs = [vo.Context(i, arith="strict") for i in range(n_dev)]
    g = vo.group.Group.local(ctxs)
    assert g.world == n_dev and g.ranks == list(range(n_dev))
    x0 = vo.workloads.vdp_x0(n) + 1e-3 * np.arange(n)[:, None]
    shards = [vo.Ensemble(ctxs[i], 2, g.shard(n, i)[1] - g.shard(n, i)[0]) for i in range(n_dev)]
    g.scatter(x0, n, 2, shards, root=0)
    for i in range(n_dev):
        lo, hi = g.shard(n, i)
        assert np.array_equal(shards[i].to_host(), x0[lo:hi])
    solvers = []
    for i in range(n_dev):
        lo, hi = g.shard(n, i)
        rhs = vo.Rhs(ctxs[i], "VDP", 2, [vo.workloads.vdp_mu(n, hi - lo, lo)])
        solvers.append(vo.RK45Solver(rhs, 0.0, tf, vo.Ensemble.from_host(ctxs[i], vo.workloads.vdp_x0(hi - lo)), 1e-3,
                                     tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6))
    tot = g.run(solvers, adaptive=True)
    ref_x, ref_stats = _whole(vo, ctxs[0], n, tf)
    finals = [s.current()[1] for s in solvers]
    for root in range(n_dev):
        for layout in ("aos", "soa"):
            got = g.gather(finals, n, root=root, layout=layout)
            assert np.array_equal(got if layout == "aos" else got.T, ref_x)
    dev = g.gather_device(finals, n, root=n_dev - 1)
    assert np.array_equal(dev.to_host(), ref_x)
    assert tot["accepted"] == int(ref_stats["accepted"].sum()) and tot["rejected"] == int(ref_stats["rejected"].sum())
    assert tot["n_traj"] == n and tot["n_done"] == n and tot["n_nonfinite"] == 0 and tot["n_stuck"] == 0
    assert tot["t_min"] == tf and tot["t_max"] == tf
    # the general, asynchronous form: shards placed at caller-chosen rows of a larger host array
    out = np.full((n + 7, 2), -1.0)
    rows = [g.shard(n, i)[1] - g.shard(n, i)[0] for i in range(n_dev)]
    offs = [g.shard(n, i)[0] + 7 for i in range(n_dev)]
    g.gather_placed(finals, rows, offs, out, root=0)
    g.sync()
    assert np.array_equal(out[7:], ref_x) and np.all(out[:7] == -1.0)
    red = g.allreduce([[1.0 + i, -2.0 * i] for i in range(n_dev)], "sum")
    assert np.array_equal(red, np.tile([sum(1.0 + i for i in range(n_dev)), sum(-2.0 * i for i in range(n_dev))], (n_dev, 1)))
    # lock-step solvers (no per-trajectory arrays): the counters come from the host-side state machine
    ls = []
    for i in range(n_dev):
        lo, hi = g.shard(n, i)
        rhs = vo.Rhs(ctxs[i], "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS))
        ls.append(vo.RK45Solver(rhs, 0.0, 0.01, vo.Ensemble.from_host(ctxs[i], vo.workloads.lorenz_x0(hi - lo, first=lo)), 1e-3, tableau=vo.ButcherTableu.builtin("RK4")))
    tot = g.run(ls)
    assert tot["accepted"] == n * ls[0].stats()["accepted"][0] and tot["n_done"] == n and tot["rejected"] == 0


def test_two_process_group_under_torchrun(vo, tmp_path):
    """The launch model of bench.py: one process per GPU, torch.distributed only to carry the NCCL id."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 1000
    out = tmp_path / "ok"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_group_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert out.read_text() == "ok"
