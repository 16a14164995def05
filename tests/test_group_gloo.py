"""The N > 1 path on CPU: world_size-2 gloo processes shard an ensemble by trajectory, integrate their shard with the
CPU oracle (standing in for the device path, which needs a GPU), then run the SAME gather / reduce code the GPU ranks
run, and the gathered result must equal the single-process solve bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import vecode_b200 as vo
    from oracle import oracle_lib as ol
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = vo.group.my_range(n_total)
    mu = vo.workloads.vdp_mu(n_total, hi - lo, lo)
    x0 = vo.workloads.vdp_x0(hi - lo)
    r = ol.rk_ensemble("VDP", mu[:, None], ol.builtin_tableau(2), 0.0, 2.0, x0, 1e-3, adaptive=True, rtol=1e-6)
    full = vo.group.gather_states(r["x"], n_total)
    rooted = vo.group.gather_states(r["x"], n_total, root=1)  # only rank 1 receives the ensemble
    assert (rooted is None) == (rank != 1) and (rooted is None or np.array_equal(rooted, full))
    # rank 0 reports DONE|STUCK (5), rank 1 DONE|NONFINITE (3): the reduction must return their union (7), not the maximum
    red = vo.group.reduce_stats(dict(accepted=r["accepted"], rejected=r["rejected"], t=r["t"], status=np.full(hi - lo, 5 if rank == 0 else 3, np.int32)))
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "red.npy"), np.array([red["accepted"], red["rejected"], red["t_min"], red["t_max"], red["status"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [101, 64])
def test_two_rank_shard_gather_reduce(tmp_path, oracle, vo, n_total):
    port = 29500 + (os.getpid() % 2000) + n_total
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    mu = vo.workloads.vdp_mu(n_total)
    ref = oracle.rk_ensemble("VDP", mu[:, None], oracle.builtin_tableau(2), 0.0, 2.0, vo.workloads.vdp_x0(n_total), 1e-3, adaptive=True, rtol=1e-6)
    full, red = np.load(tmp_path / "full.npy"), np.load(tmp_path / "red.npy")
    assert full.shape == (n_total, 2) and np.array_equal(full, ref["x"])
    assert int(red[0]) == int(ref["accepted"].sum()) and int(red[1]) == int(ref["rejected"].sum())
    assert red[2] == ref["t"].min() and red[3] == ref["t"].max() and int(red[4]) == 7


def test_gather_complex_states_single_process(vo):
    z = np.arange(12).reshape(6, 2) * (1 + 2j)
    assert np.array_equal(vo.group.gather_states(z, 6), z)
