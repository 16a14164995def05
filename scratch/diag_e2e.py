import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vecode_b200 as vo
N = 1_000_000
for arith in ("strict", "fast", "strict", "fast"):
    ctx = vo.Context.on_torch_stream(0, arith=arith)
    x0h = vo.workloads.lorenz_x0(N)
    rhs = vo.Rhs(ctx, "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS))
    ex0 = vo.Ensemble(ctx, 3, N)
    sol = vo.RK45Solver(rhs, 0.0, 1.0, ex0, 1e-3, tableau=vo.ButcherTableu.builtin("RK4"))
    pin = torch.from_numpy(x0h.copy()).pin_memory(); pout = torch.empty_like(pin).pin_memory()
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ex0.upload(pin.numpy(), "aos"); t1 = time.perf_counter()
        sol.reset(ex0); torch.cuda.synchronize(); t2 = time.perf_counter()
        l0 = ctx.launch_count
        st = sol.run(); t3 = time.perf_counter()
        torch.cuda.synchronize(); t4 = time.perf_counter()
        sol.current()[1].to_host("aos", out=pout.numpy()); t5 = time.perf_counter()
        print(f"{arith} rep{rep}: upload {1e3*(t1-t0):.2f} reset {1e3*(t2-t1):.2f} run-cpu {1e3*(t3-t2):.2f} run-sync {1e3*(t4-t3):.2f} download {1e3*(t5-t4):.2f} ms; launches {ctx.launch_count-l0} steps {st.counts['Step']//N} finite {np.isfinite(pout.numpy()).all()}")
    # adaptive vdp
    mu = vo.workloads.vdp_mu(N)
    rhs2 = vo.Rhs(ctx, "VDP", 2, [mu])
    e2 = vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(N))
    s2 = vo.RK45Solver(rhs2, 0.0, 20.0, e2, 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    for rep in range(2):
        s2.reset(e2); torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = ctx.launch_count
        st = s2.run(adaptive=True); torch.cuda.synchronize(); t1 = time.perf_counter()
        stt = s2.stats()
        att = stt["accepted"] + stt["rejected"]
        print(f"{arith} vdp rep{rep}: {1e3*(t1-t0):.1f} ms launches {ctx.launch_count-l0} attempts total {att.sum()} max/traj {att.max()} min {att.min()} argmax {att.argmax()} status {np.unique(stt['status'])}")
