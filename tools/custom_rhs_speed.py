"""Launch time of the register-resident RK4 kernel with a built-in and with a run-time compiled (NVRTC) Lorenz-63 RHS."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vecode_b200 as vo

BODY = "dx[0] = p[0]*(x[1]-x[0]); dx[1] = x[0]*(p[1]-x[2]) - x[1]; dx[2] = x[0]*x[1] - p[2]*x[2];"
n, nb, rounds = 1_000_000, 16, 200
for arith in ("fast", "strict"):
    ctx = vo.Context.on_torch_stream(0, arith=arith)
    x0 = vo.Ensemble.from_host(ctx, vo.workloads.lorenz_x0(n))
    for name, rhs in (("builtin", vo.Rhs(ctx, "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS))), ("custom", vo.Rhs.custom(ctx, BODY, 3, list(vo.workloads.LORENZ_PARAMS)))):
        solvers = [vo.RK45Solver(rhs, 0.0, 1e9, x0, 1e-3, tableau=vo.ButcherTableu.builtin("RK4")) for _ in range(nb)]
        vo.step_many(solvers, False, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vo.step_many(solvers, False, rounds); e1.record(); torch.cuda.synchronize()
        print(f"{arith} {name}: {e0.elapsed_time(e1) / (rounds * nb) * 1e3:.2f} us per launch of 10^6 trajectories")
        del solvers
