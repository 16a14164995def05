"""User-defined norms on the B200: the reference's `Normed<T, V>` is the user's trait impl (src/base/ode.rs:9-11; RK45Solver::norm,
rk.rs:302-304) and ExpCFMSolver takes a NormFn closure (src/exp/cfm.rs:105, 214-216). Here the functor crosses the C ABI as source
(vo_normfn_create) and is compiled into the control kernels. Checked against the pure-Python restatement driving the same norm
as a Python callable, and — for the built-in 2-norm restated as a functor — against the compiled-in kernels bit for bit."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WRMS_MAP, WRMS_FINISH = "m = (e * e + im * im) / (1.0 + i);", "r = sqrt(acc / n);"


def wrms(v):
    acc = 0.0
    for i, e in enumerate(v):
        acc = acc + (e * e + 0.0 * 0.0) / (1.0 + i)
    return math.sqrt(acc / len(v))


def test_norm_on_its_own(vo, ctx):
    rng = np.random.default_rng(1)
    f = vo.NormFn(ctx, WRMS_MAP, "sum", WRMS_FINISH)
    g = vo.NormFn(ctx, "m = fabs(e) * (i + 1);", "max")
    for d, n in [(5, 1000), (64, 33), (100_003, 1), (5000, 3)]:
        x = rng.standard_normal((n, d))
        e = vo.Ensemble.from_host(ctx, x)
        w = 1.0 / (1.0 + np.arange(d))
        np.testing.assert_allclose(e.norm(f), np.sqrt((x * x * w).sum(axis=1) / d), rtol=1e-13)
        np.testing.assert_allclose(e.norm(g), (np.abs(x) * (np.arange(d) + 1)).max(axis=1), rtol=1e-15)
    with pytest.raises(vo.VecOdeError) as ei:
        vo.NormFn(ctx, "m = nope;")
    assert "norm_map_body(1)" in str(ei.value)


@pytest.mark.parametrize("n", [2500, 300])  # TMA-staged / register-prefetch control kernels
def test_builtin_two_norm_restated_as_a_functor_gives_the_builtin_bits(vo, ctx, n):
    mu = vo.workloads.vdp_mu(n)
    res = []
    for norm in ("L2", vo.NormFn(ctx, "m = e * e;", "sum", "r = sqrt(acc);")):
        s = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu]), 0.0, 2.0, vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(n)), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5"))
        s.with_tolerance(1e-6, 1e-6).with_norm(norm)
        assert s.run(adaptive=True).kind == "Done"
        res.append((s.current()[1].to_host(), s.stats()))
    assert np.array_equal(res[0][0], res[1][0])
    for k in ("accepted", "rejected", "h", "dx_norm"):
        assert np.array_equal(res[0][1][k], res[1][1][k]), k


def test_weighted_norm_on_a_builtin_family_against_the_python_restatement(vo, ctx, oracle):
    from oracle import vecode_oracle as po
    n = 2048
    mu = vo.workloads.vdp_mu(n)
    f = vo.NormFn(ctx, WRMS_MAP, "sum", WRMS_FINISH)
    ac, b, be, ns = oracle.builtin_tableau(2)
    runs = {}
    for path in (0, 1):  # register-resident kernels / stage path (norm evaluated by the functor's own kernel ahead of the commit)
        s = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu]), 0.0, 2.0, vo.Ensemble.from_host(ctx, vo.workloads.vdp_x0(n)), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5"))
        s.with_tolerance(1e-7, 1e-7).with_norm(f)
        if path:
            s.set_stage_path(True)
        assert s.run(adaptive=True).kind == "Done"
        runs[path] = (s.current()[1].to_host(), s.stats())
    got, st = runs[0]
    assert np.array_equal(got, runs[1][0]) and np.array_equal(st["accepted"], runs[1][1]["accepted"]) and np.array_equal(st["rejected"], runs[1][1]["rejected"])
    x0 = vo.workloads.vdp_x0(n)
    for i in range(0, n, 256):
        m = float(mu[i])

        def vdp(t, x, dx, m=m):
            dx[0] = x[1]
            dx[1] = (m * (1.0 - x[0] * x[0])) * x[1] - x[0]
        r = po.RKSolver(vdp, (list(ac), list(b), list(be), ns), 0.0, 2.0, list(x0[i]), 1e-3).with_tolerance(1e-7, 1e-7)
        r.norm_kind = wrms
        r.run(adaptive=True)
        assert (int(st["accepted"][i]), int(st["rejected"][i])) == (r.n_accept, r.n_reject), i
        # strict arithmetic follows the restatement bit for bit except where glibc's pow is not correctly rounded (~0.1 % of calls)
        assert np.abs(got[i] - np.array(r.x)).max() <= 1e-10, i
        assert abs(st["dx_norm"][i] - r.dx_norm) <= 1e-9 * r.dx_norm
    # the norm changes the solve: the plain 2-norm takes a different number of steps
    s2 = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu]), 0.0, 2.0, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-7, 1e-7)
    s2.run(adaptive=True)
    assert s2.stats()["accepted"].sum() != st["accepted"].sum()


def test_weighted_norm_with_a_user_rhs(vo, ctx, oracle):
    from oracle import vecode_oracle as po
    from test_gpu_custom_rhs import LV_BODY, _lv_inputs, lv_f
    n = 1500
    params, x0 = _lv_inputs(n)
    rhs = vo.Rhs.custom(ctx, LV_BODY, 2, [params[:, q].copy() for q in range(4)])
    s = vo.RK45Solver(rhs, 0.0, 2.0, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-6, 1e-6)
    s.with_norm(vo.NormFn(ctx, "m = fabs(e);", "max"))
    assert s.run(adaptive=True).kind == "Done"
    got, st = s.current()[1].to_host(), s.stats()
    ac, b, be, ns = oracle.builtin_tableau(2)
    for i in range(0, n, 250):
        r = po.RKSolver(lv_f(params[i]), (list(ac), list(b), list(be), ns), 0.0, 2.0, list(x0[i]), 1e-3).with_tolerance(1e-6, 1e-6)
        r.norm_kind = lambda v: max(abs(e) for e in v)
        r.run(adaptive=True)
        assert (int(st["accepted"][i]), int(st["rejected"][i])) == (r.n_accept, r.n_reject)
        assert np.abs(got[i] - np.array(r.x)).max() <= 2e-5


def test_weighted_norm_on_one_large_state(vo, ctx, oracle):
    """Stage path, one state of d components: the functor's tree reduction feeds the host controller (ode.rs:311-334)."""
    from oracle import vecode_oracle as po
    d = 1024
    j = np.arange(d)
    u0 = vo.workloads.heat_u0(d) + 0.25 * np.cos(np.pi * j)
    f = vo.NormFn(ctx, WRMS_MAP, "sum", WRMS_FINISH)
    s = vo.RK45Solver(vo.Rhs(ctx, "HEAT1D", d, [1.0]), 0.0, 1.0, vo.Ensemble.from_host(ctx, u0[None, :]), 0.01).with_tolerance(1e-6, 1e-6).with_norm(f)
    assert s.run(adaptive=True).kind == "Done"

    def heat(t, x, dx):
        nn = len(x)
        for k in range(nn):
            dx[k] = 1.0 * ((x[k - 1] + x[(k + 1) % nn]) - 2.0 * x[k])
    ac, b, be, ns = oracle.builtin_tableau(0)
    r = po.RKSolver(heat, (list(ac), list(b), list(be), ns), 0.0, 1.0, list(u0), 0.01).with_tolerance(1e-6, 1e-6)
    r.norm_kind = wrms
    r.run(adaptive=True)
    st = s.stats()
    assert (int(st["accepted"][0]), int(st["rejected"][0])) == (r.n_accept, r.n_reject)
    assert np.abs(s.current()[1].to_host()[0] - np.array(r.x)).max() <= 1e-12


def test_user_norm_in_the_exponential_integrators(vo, ctx):
    """ExpCFMSolver's NormFn (cfm.rs:105, 214-216): the 2-norm restated as a functor reproduces the compiled-in kernel's step sequence;
    a weighted norm is checked against the controller restated on the host from the embedded error of the pure-Python oracle."""
    from oracle import exp_oracle as eo
    from test_gpu_exp import _system
    n, N, h, rtol = 16, 21, 0.2, 1e-5
    B0, B1, gp, psi0 = _system(vo, n, N)
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    a = vo.ExpCFMSolver(sp, gp, 0.0, 2.0, psi0, h).with_tolerance(rtol, rtol)
    b = vo.ExpCFMSolver(sp, gp, 0.0, 2.0, psi0, h).with_tolerance(rtol, rtol).with_norm(vo.NormFn(ctx, "m = e * e + im * im;", "sum", "r = sqrt(acc);"))
    a.run(adaptive=True), b.run(adaptive=True)
    assert np.array_equal(a.stats()["accepted"], b.stats()["accepted"]) and np.array_equal(a.stats()["rejected"], b.stats()["rejected"])
    assert np.abs(a.current()[1] - b.current()[1]).max() <= 1e-13
    c = vo.ExpCFMSolver(sp, gp, 0.0, 2.0, psi0, h).with_tolerance(rtol, rtol).with_norm(vo.NormFn(ctx, WRMS_MAP, "sum", WRMS_FINISH))
    c.step_adaptive(), c.step_adaptive()  # Chkpt at t0, then one attempt
    st = c.stats()
    spy = eo.BasisSplit([[[(complex(z).real, complex(z).imag) for z in row] for row in B] for B in (B0, B1)])
    for i in range(0, N, 5):
        g = eo.gen_cos([tuple(r) for r in gp[i]], 2, 2)
        x0 = [(z.real, z.imag) for z in psi0[i]]
        _, xe = eo.cfm_general(spy, lambda ts: [g(t) for t in ts], 0.0, x0, h, eo.C_GAUSS_LEGENDRE_4, eo.CFM_R4_J2_GL, eo.CFM_R2_J1_GL)
        acc = 0.0
        for k, (re, im) in enumerate(xe):
            acc += (re * re + im * im) / (1.0 + k)
        dxn = math.sqrt(acc / n)
        fct = rtol / dxn
        new_h = min(max(min(max(0.9 * fct ** (1.0 / 3.0), 0.3), 2.0) * h, 1e-6), 1.0)
        assert abs(st["dx_norm"][i] - dxn) <= 1e-12 * dxn and abs(st["h"][i] - new_h) <= 1e-12 * new_h
        assert int(st["accepted"][i]) == int(not (fct <= 1.0))
