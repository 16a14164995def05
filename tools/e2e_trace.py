#!/usr/bin/env python
"""Host-side timeline of one sharded e2e solve of config 3 (pipeline.ShardedChunkedSolve): per chunk the upload, the integration,
the wait for the gather's turn + its enqueue, and the final sync, on every rank. Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_trace.py [--parts 4]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", type=int, default=4)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import vecode_b200 as vo
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    bench.numa_bind(torch, local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = vo.group.Group.from_torch_distributed(vo.Context(local, arith="fast", urgency=8))
    n_total = bench.N_TRAJ * world
    mu_il = vo.workloads.vdp_mu(n_total)[rank::world].copy()
    tab = vo.ButcherTableu.builtin("DOPRI5")

    def make(ctx, lo, hi, x0):
        return vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu_il[lo:hi].copy()]), 0.0, 20.0, x0, 1e-3, tableau=tab).with_tolerance(1e-6, 1e-6)
    sh = vo.pipeline.ShardedChunkedSolve(group, n_total, 2, make, parts=args.parts, arith="fast", interleave=True)
    pin_in = torch.from_numpy(vo.workloads.vdp_x0(sh.n_local)).pin_memory()
    pin_full = torch.empty((n_total, 2), dtype=torch.float64).pin_memory() if rank == 0 else None
    full = None if pin_full is None else pin_full.numpy()
    for rep in range(args.reps):
        dist.barrier()
        torch.cuda.synchronize()
        sh.trace = [] if rep == args.reps - 1 else None
        t0 = time.perf_counter()
        sh.solve(pin_in.numpy(), full, adaptive=True)
        t1 = time.perf_counter()
        dist.barrier()
        if rank == 0:
            print(f"rep {rep}: rank-0 solve {1e3 * (t1 - t0):.2f} ms", flush=True)
    for r in range(world):
        dist.barrier()
        if r == rank and (rank in (0, 1, world - 1)):
            for q, ph, a, b in sorted(sh.trace, key=lambda e: e[2]):
                print(f"rank {rank} chunk {q:2d} {ph:15s} {1e3 * (a - t0):7.2f} -> {1e3 * (b - t0):7.2f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
