"""User-defined right-hand sides (vo_rhs_create_custom): the C-ABI replacement of the reference's RHS closure
`f: FnMut(T, &V, &mut V)` (src/base/rk.rs:97), compiled at run time into the same fused kernels as the built-in families.

Bars: a body that restates a built-in family gives the built-in kernel's bits on every path; a body with no built-in
counterpart is bit-exact against the pure-Python restatement of rk.rs / ode.rs driving the same function (strict
arithmetic, fixed step), and within rtol with equal accept / reject counts on adaptive runs.
"""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LORENZ_BODY = """
dx[0] = p[0] * (x[1] - x[0]);
dx[1] = x[0] * (p[1] - x[2]) - x[1];
dx[2] = x[0] * x[1] - p[2] * x[2];
"""
# Lotka-Volterra with a time-dependent harvest term: no built-in family, uses t and four parameters
LV_BODY = """
const double xy = x[0] * x[1];
dx[0] = p[0] * x[0] - p[1] * xy;
dx[1] = p[3] * xy - p[2] * x[1] - (0.01 * t) * x[1];
"""


def lv_f(p):
    def f(t, x, dx):
        xy = x[0] * x[1]
        dx[0] = p[0] * x[0] - p[1] * xy
        dx[1] = p[3] * xy - p[2] * x[1] - (0.01 * t) * x[1]
    return f


def _lorenz_pair(vo, ctx, n):
    x0 = vo.workloads.lorenz_x0(n)
    builtin = vo.Rhs(ctx, "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS))
    custom = vo.Rhs.custom(ctx, LORENZ_BODY, 3, list(vo.workloads.LORENZ_PARAMS))
    return x0, builtin, custom


@pytest.mark.parametrize("n", [2048, 1001])  # TMA-staged kernels / odd N: register-prefetch kernels
@pytest.mark.parametrize("tab", ["RK4", "RKF45_REF"])
def test_custom_lorenz_fixed_equals_builtin_bits(vo, ctx, n, tab):
    x0, builtin, custom = _lorenz_pair(vo, ctx, n)
    out = []
    for rhs in (builtin, custom):
        s = vo.RK45Solver(rhs, 0.0, 0.05, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin(tab))
        assert s.run().kind == "Done"
        out.append(s.current()[1].to_host())
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("n", [4096, 640, 333])  # two-trajectory kernel / one-trajectory staged kernel / odd N
def test_custom_lorenz_adaptive_equals_builtin_bits(vo, n, arith):
    ctx = vo.Context(0, arith=arith)
    x0, builtin, custom = _lorenz_pair(vo, ctx, n)
    out = []
    for rhs in (builtin, custom):
        s = vo.RK45Solver(rhs, 0.0, 0.5, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5"))
        s.with_tolerance(1e-6, 1e-7)
        assert s.run(adaptive=True).kind == "Done"
        st = s.stats()
        out.append((s.current()[1].to_host(), st["accepted"], st["rejected"]))
    if arith == "strict":
        assert np.array_equal(out[0][0], out[1][0])
    else:  # nvcc and NVRTC are free to contract differently
        assert np.abs(out[0][0] - out[1][0]).max() <= 1e-9 * np.abs(out[0][0]).max()
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    ctx.close()


def test_custom_stage_path_and_eval_equal_builtin_bits(vo, ctx):
    n = 515
    x0, builtin, custom = _lorenz_pair(vo, ctx, n)
    tableau = vo.ButcherTableu.builtin("DOPRI5")
    got = []
    for rhs in (builtin, custom):
        X0 = vo.Ensemble.from_host(ctx, x0)
        solver = vo.RK45Solver(rhs, 0.0, 1.0, X0, 0.01, tableau=tableau)
        nx, xe, dx = vo.Ensemble(ctx, 3, n), vo.Ensemble(ctx, 3, n), vo.Ensemble(ctx, 3, n)
        K = [vo.Ensemble(ctx, 3, n) for _ in range(7)]
        solver.try_step(0.37, 0.0123, nx, xe, K)
        rhs(0.2, X0, dx)
        got.append([nx.to_host(), xe.to_host(), dx.to_host()] + [k.to_host() for k in K])
    for a, b in zip(*got):
        assert np.array_equal(a, b)


def _lv_inputs(n):
    rng = np.random.default_rng(5)
    params = np.stack([1.0 + 0.2 * rng.random(n), 0.4 + 0.1 * rng.random(n), 0.8 + 0.2 * rng.random(n), 0.1 + 0.05 * rng.random(n)], axis=1)
    x0 = np.stack([8.0 + rng.random(n), 3.0 + rng.random(n)], axis=1)
    return params, x0


def test_custom_lotka_volterra_fixed_bit_exact_vs_python_restatement(vo, ctx, oracle):
    """The reference's rk_step / step() restated in pure Python (oracle/vecode_oracle.py) drives the same closure."""
    from oracle import vecode_oracle as po
    n = 130
    params, x0 = _lv_inputs(n)
    rhs = vo.Rhs.custom(ctx, LV_BODY, 2, [params[:, q].copy() for q in range(4)])
    for tab_name, tab_id in (("RKF45_REF", 0), ("DOPRI5", 2)):
        s = vo.RK45Solver(rhs, 0.0, 0.3, vo.Ensemble.from_host(ctx, x0), 0.01, tableau=vo.ButcherTableu.builtin(tab_name))
        assert s.run().kind == "Done"
        got = s.current()[1].to_host()
        ac, b, be, ns = oracle.builtin_tableau(tab_id)
        for i in range(0, n, 13):
            r = po.RKSolver(lv_f(params[i]), (list(ac), list(b), None if be is None else list(be), ns), 0.0, 0.3, list(x0[i]), 0.01)
            r.run()
            assert np.array_equal(got[i], np.array(r.x)), (tab_name, i, got[i], r.x)


def test_custom_lotka_volterra_adaptive_vs_python_restatement(vo, ctx, oracle):
    from oracle import vecode_oracle as po
    n = 2048
    params, x0 = _lv_inputs(n)
    rhs = vo.Rhs.custom(ctx, LV_BODY, 2, [params[:, q].copy() for q in range(4)])
    s = vo.RK45Solver(rhs, 0.0, 2.0, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5"))
    s.with_tolerance(1e-6, 1e-6)
    assert s.run(adaptive=True).kind == "Done"
    got, st = s.current()[1].to_host(), s.stats()
    ac, b, be, ns = oracle.builtin_tableau(2)
    for i in range(0, n, 256):
        r = po.RKSolver(lv_f(params[i]), (list(ac), list(b), list(be), ns), 0.0, 2.0, list(x0[i]), 1e-3)
        r.with_tolerance(1e-6, 1e-6)
        r.run(adaptive=True)
        assert np.abs(got[i] - np.array(r.x)).max() <= 1e-6 * 20  # rtol at t_end, a few dozen accepted steps
        assert (int(st["accepted"][i]), int(st["rejected"][i])) == (r.n_accept, r.n_reject)


def test_custom_rhs_compile_error_is_reported(vo, ctx):
    with pytest.raises(vo.VecOdeError) as ei:
        vo.Rhs.custom(ctx, "dx[0] = undefined_symbol;", 1, [])
    assert "rhs_body(1)" in str(ei.value) and "undefined_symbol" in str(ei.value)
    with pytest.raises(vo.VecOdeError):
        vo.Rhs.custom(ctx, "dx[0] = x[0];", 33, [])  # d > 32


def test_custom_rhs_largest_shape_eight_components_eight_per_trajectory_parameters(vo, ctx):
    """d = 8 with 8 per-trajectory parameters: the widest tiles (18 staged rows, > 48 KB of shared memory per CTA) and the
    highest register pressure the run-time compiled kernels see. dx_c = p_c x_c has the closed form x_c(0) exp(p_c t); the
    fixed-step run must also agree bit for bit between the register-resident kernel and the stage-path kernel of the same body."""
    n, d = 4096, 8
    rng = np.random.default_rng(11)
    p = -2.0 * rng.random((n, d))
    x0 = 0.5 + rng.random((n, d))
    body = "\n".join(f"dx[{c}] = p[{c}] * x[{c}];" for c in range(d))
    custom = vo.Rhs.custom(ctx, body, d, [p[:, c].copy() for c in range(d)])
    # adaptive DoPri5 on the two-trajectory control kernel
    s = vo.RK45Solver(custom, 0.0, 1.0, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-8, 1e-8)
    assert s.run(adaptive=True).kind == "Done"
    got = s.current()[1].to_host()
    assert np.abs(got - x0 * np.exp(p * 1.0)).max() <= 1e-6
    assert np.all(s.stats()["status"] == 1)
    # fixed-step RKF45 (the reference's literal tableau): register-resident kernel == stage-path kernel
    out = []
    for stage in (False, True):
        s = vo.RK45Solver(custom, 0.0, 0.2, vo.Ensemble.from_host(ctx, x0), 0.01)
        s.set_stage_path(stage)
        assert s.run().kind == "Done"
        out.append(s.current()[1].to_host())
    assert np.array_equal(out[0], out[1])


# ---- user stencils on one grid state, and pointwise systems wider than the register-resident kernels take -----------------------
HEAT_STENCIL = "du = p[0] * ((u[0] + u[2]) - 2.0 * u[1]);"


@pytest.mark.parametrize("d", [1023, 4096, (1 << 18) + 5])
@pytest.mark.parametrize("tab", ["RK4", "RKF45_REF"])
def test_user_stencil_restating_the_heat_equation_gives_the_builtin_bits(vo, ctx, d, tab):
    """The compiled-in HEAT1D family written as a user stencil (vo_rhs_create_custom_stencil): same operations, same order, so the
    strict-arithmetic results are bit-identical — fixed steps with and without the error estimate, and vo_rhs_eval."""
    u0 = vo.workloads.heat_u0(d)
    outs = []
    for rhs in (vo.Rhs(ctx, "HEAT1D", d, [0.7]), vo.Rhs.custom_stencil(ctx, HEAT_STENCIL, d, 1, [0.7])):
        s = vo.RK45Solver(rhs, 0.0, 1.0e9, vo.Ensemble.from_host(ctx, u0[None, :]), 0.2, tableau=vo.ButcherTableu.builtin(tab))
        s.run(max_calls=8)
        dx = vo.Ensemble(ctx, d, 1)
        rhs(0.0, vo.Ensemble.from_host(ctx, u0[None, :]), dx)
        outs.append((s.current()[1].to_host()[0], dx.to_host()[0]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_user_stencil_adaptive_and_fourth_order_laplacian(vo, ctx, oracle):
    """A radius-2 stencil with a time-dependent source at one grid point, adaptive RKF45 on one state: against the pure-Python
    restatement of rk.rs / ode.rs driving the same closure (accepted / rejected counts equal, states to rounding)."""
    from oracle import vecode_oracle as po
    d = 300
    body = "du = (-u[0] + 16.0 * u[1] - 30.0 * u[2] + 16.0 * u[3] - u[4]) * (p[0] / 12.0) + (j == 7 ? p[1] * sin(3.0 * t) : 0.0);"
    u0 = vo.workloads.heat_u0(d) + 0.2 * np.cos(np.pi * np.arange(d))
    rhs = vo.Rhs.custom_stencil(ctx, body, d, 2, [0.6, 0.5])
    s = vo.RK45Solver(rhs, 0.0, 1.5, vo.Ensemble.from_host(ctx, u0[None, :]), 0.01).with_tolerance(1e-7, 1e-7)
    assert s.run(adaptive=True).kind == "Done"

    def f(t, x, dx):
        n = len(x)
        for j in range(n):
            dx[j] = (-x[j - 2] + 16.0 * x[j - 1] - 30.0 * x[j] + 16.0 * x[(j + 1) % n] - x[(j + 2) % n]) * (0.6 / 12.0) + (0.5 * math.sin(3.0 * t) if j == 7 else 0.0)
    ac, b, be, ns = oracle.builtin_tableau(0)
    r = po.RKSolver(f, (list(ac), list(b), list(be), ns), 0.0, 1.5, list(u0), 0.01).with_tolerance(1e-7, 1e-7)
    r.run(adaptive=True)
    st = s.stats()
    assert (int(st["accepted"][0]), int(st["rejected"][0])) == (r.n_accept, r.n_reject)
    assert np.abs(s.current()[1].to_host()[0] - np.array(r.x)).max() <= 1e-12
    with pytest.raises(vo.VecOdeError):  # an ensemble of grids is not what a stencil handle is for
        vo.RK45Solver(rhs, 0.0, 1.0, vo.Ensemble(ctx, d, 3), 0.01).run(max_calls=3)


def test_wide_pointwise_user_rhs_on_the_stage_path(vo, ctx, oracle):
    """24 coupled components per trajectory (a ring of cubic oscillators): beyond the register-resident kernels (d <= 8), so the
    solver takes the stage path — fixed-step bit-exact against the pure-Python restatement, adaptive with equal counts."""
    from oracle import vecode_oracle as po
    D, n = 24, 700
    body = "\n".join(f"dx[{c}] = p[0] * (x[{(c + 1) % D}] - x[{c}]) - x[{c}] * x[{c}] * x[{c}];" for c in range(D))
    rng = np.random.default_rng(8)
    x0 = rng.uniform(-1.0, 1.0, (n, D))
    kap = rng.uniform(0.5, 2.0, n)
    rhs = vo.Rhs.custom(ctx, body, D, [kap])

    def mk(k):
        def f(t, x, dx):
            for c in range(D):
                dx[c] = k * (x[(c + 1) % D] - x[c]) - x[c] * x[c] * x[c]
        return f
    ac, b, be, ns = oracle.builtin_tableau(2)
    s = vo.RK45Solver(rhs, 0.0, 0.2, vo.Ensemble.from_host(ctx, x0), 0.01, tableau=vo.ButcherTableu.builtin("DOPRI5"))
    assert s.run().kind == "Done"
    got = s.current()[1].to_host()
    for i in range(0, n, 97):
        r = po.RKSolver(mk(float(kap[i])), (list(ac), list(b), list(be), ns), 0.0, 0.2, list(x0[i]), 0.01)
        r.run()
        assert np.array_equal(got[i], np.array(r.x)), i
    s = vo.RK45Solver(rhs, 0.0, 1.0, vo.Ensemble.from_host(ctx, x0), 0.01, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    assert s.run(adaptive=True).kind == "Done"
    got, st = s.current()[1].to_host(), s.stats()
    for i in range(0, n, 233):
        r = po.RKSolver(mk(float(kap[i])), (list(ac), list(b), list(be), ns), 0.0, 1.0, list(x0[i]), 0.01).with_tolerance(1e-6, 1e-6)
        r.run(adaptive=True)
        assert (int(st["accepted"][i]), int(st["rejected"][i])) == (r.n_accept, r.n_reject)
        assert np.abs(got[i] - np.array(r.x)).max() <= 1e-9


@pytest.mark.parametrize("radius,body", [(1, HEAT_STENCIL), (2, "du = (-u[0] + 16.0 * u[1] - 30.0 * u[2] + 16.0 * u[3] - u[4]) * (p[0] / 12.0) + 1e-3 * sin(t) * (j % 5);"),
                                         (3, "du = p[0] * (u[0] - u[6]) + 0.1 * (u[2] + u[4] - 2.0 * u[3]);"), (4, "du = p[0] * (u[0] + u[8] - 2.0 * u[4]);")])
@pytest.mark.parametrize("tab", ["RK4", "DOPRI5"])
def test_user_stencil_tma_staged_kernel_equals_the_plain_kernel_bitwise(vo, ctx, radius, body, tab):
    """Large even grids take the TMA-staged stencil kernel (rk_stage_stencil.cuh: stage_stencil_tma_kernel), everything else the
    plain one: same operations per point, so the strict results must be the same bits — fixed steps with the error estimate on,
    grid sizes that end in a full tile, a partial tile and a two-point remainder."""
    for d in (8192, 8192 + 514, 4 * 1024 + 2):
        u0 = vo.workloads.heat_u0(d) + 0.1 * np.cos(np.pi * np.arange(d))
        outs = []
        for plain in ("", "1"):
            if plain:
                os.environ["VECODE_STENCIL_PLAIN"] = "1"
            else:
                os.environ.pop("VECODE_STENCIL_PLAIN", None)
            try:
                rhs = vo.Rhs.custom_stencil(ctx, body, d, radius, [0.3])
                s = vo.RK45Solver(rhs, 0.0, 1.0e9, vo.Ensemble.from_host(ctx, u0[None, :]), 0.05, tableau=vo.ButcherTableu.builtin(tab))
                s.run(max_calls=6)
                outs.append(s.current()[1].to_host()[0])
            finally:
                os.environ.pop("VECODE_STENCIL_PLAIN", None)
        assert np.array_equal(outs[0], outs[1]), (d, np.abs(outs[0] - outs[1]).max())
        assert np.isfinite(outs[0]).all() and np.abs(outs[0] - u0).max() > 1e-6
