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
