import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import vecode_b200 as vo
ctx = vo.Context.on_torch_stream(0, arith="fast")
N, n = 100_000, 64
H0, H1 = vo.workloads.schrodinger_system(n)
gp = vo.workloads.schrodinger_drive(N)
psi0 = np.zeros((N, n), dtype=np.complex128); psi0[:, 0] = 1.0
for deg in (1, 2, 4, 8, 0):
    sp = vo.DenseBasisSplit(ctx, np.stack([-1j * H0, -1j * H1]), taylor_degree=deg)
    for dyn in (False, True):
        s = vo.ExpCFMSolver(sp, gp, 0.0, 1e9, psi0, 0.1).no_adaptive()
        if dyn:
            s.dynamic_grouping()
        s.run(max_calls=4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run(max_calls=10); e1.record(); torch.cuda.synchronize()
        print(f"forced degree {deg or 'auto'} dyn={dyn}: {e0.elapsed_time(e1) / 10:.3f} ms per step", flush=True)
