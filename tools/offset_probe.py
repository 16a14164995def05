#!/usr/bin/env python
"""Where does a short timed region lose time?  total(K) = offset + K * slope for K launches after a synchronize.

For every K: synchronise, record an event, enqueue K one-event sweeps (the bench's run_steps), record an event; repeated R times.
Prints the median event time per K, the host time spent enqueueing, and a least-squares (offset, slope).
  python tools/offset_probe.py [--workload vdp_dopri5|lorenz_rk4] [--arith fast|strict]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vdp_dopri5")
    ap.add_argument("--arith", default="fast")
    ap.add_argument("--reps", type=int, default=15)
    args = ap.parse_args()
    import torch
    import vecode_b200 as vo
    torch.cuda.set_device(0)
    W = bench.WORKLOADS[args.workload]
    ctx = vo.Context.on_torch_stream(0, arith=args.arith)
    nb = bench.n_batches_for(W, bench.N_TRAJ)
    w = W(vo, ctx, 0, 1, nb)
    w.run_steps(nb)
    w.run_steps(3000)  # spin-up
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = []
    for K in (1, 2, 4, 7, 10, 20, 40, 80, 160, 320):
        ms, host = [], []
        for _ in range(args.reps):
            w.run_steps(5)
            torch.cuda.synchronize()
            ev0.record()
            h0 = time.perf_counter()
            w.run_steps(K)
            h1 = time.perf_counter()
            ev1.record()
            torch.cuda.synchronize()
            ms.append(ev0.elapsed_time(ev1) * 1e3), host.append((h1 - h0) * 1e6)
        rows.append({"K": K, "us_median": float(np.median(ms)), "us_min": float(np.min(ms)), "host_enqueue_us": float(np.median(host))})
        print(json.dumps(rows[-1]), flush=True)
    big = [r for r in rows if r["K"] >= 20]
    A = np.array([[1.0, r["K"]] for r in big])
    off, slope = np.linalg.lstsq(A, np.array([r["us_median"] for r in big]), rcond=None)[0]
    print(json.dumps({"workload": args.workload, "arith": args.arith, "batches": nb, "offset_us": float(off), "slope_us_per_launch": float(slope)}))


if __name__ == "__main__":
    main()
