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

# ---- a user stencil on one grid state against the compiled-in heat kernels (stage path, RK4 step of 2^26 points) and a 24-component
# ---- pointwise user RHS on the stage path
import numpy as np
d = 1 << 26
for arith in ("fast", "strict"):
    ctx = vo.Context.on_torch_stream(0, arith=arith)
    u0 = vo.Ensemble.from_host(ctx, vo.workloads.heat_u0(d)[None, :])
    for name, rhs in (("builtin HEAT1D", vo.Rhs(ctx, "HEAT1D", d, [1.0])), ("user stencil", vo.Rhs.custom_stencil(ctx, "du = p[0] * ((u[0] + u[2]) - 2.0 * u[1]);", d, 1, [1.0]))):
        s = vo.RK45Solver(rhs, 0.0, 1e9, u0, 0.25, tableau=vo.ButcherTableu.builtin("RK4"))
        vo.step_many([s], False, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vo.step_many([s], False, 20); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{arith} {name}: {ms:.3f} ms per RK4 step of 2^26 points = {104.0 * d / (ms * 1e-3) / 1e12:.2f} TB/s of the 104-byte model")
        del s
D, n = 24, 1_000_000
body = "\n".join(f"dx[{c}] = p[0] * (x[{(c + 1) % D}] - x[{c}]) - x[{c}] * x[{c}] * x[{c}];" for c in range(D))
ctx = vo.Context.on_torch_stream(0, arith="fast")
x0 = vo.Ensemble.from_host(ctx, np.random.default_rng(0).uniform(-1, 1, (n, D)))
s = vo.RK45Solver(vo.Rhs.custom(ctx, body, D, [1.0]), 0.0, 1e9, x0, 1e-3, tableau=vo.ButcherTableu.builtin("RK4"))
vo.step_many([s], False, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); vo.step_many([s], False, 10); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"fast 24-component user RHS, stage path: {ms:.3f} ms per RK4 step of 10^6 trajectories = {104.0 * D * n / (ms * 1e-3) / 1e12:.2f} TB/s of the 13-pass model")
