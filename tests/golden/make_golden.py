"""Regenerates tests/golden/rk_known_answers.json from the CPU oracle.

The reference crate cannot be compiled here (no Rust toolchain) and its own tests assert nothing, so these known
answers come from the two independent restatements of its source (oracle/vecode_oracle.cpp and
oracle/vecode_oracle.py), which must agree BIT FOR BIT before anything is written. Floats are stored with repr()
round-trip precision. Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_lib as ol  # noqa: E402
from oracle import vecode_oracle as po  # noqa: E402


def py_solve(rhs_id, params, tab_name, t0, tf, x0, h, adaptive=False, no_adaptive=False, rtol=None, norm=0, t_list=None):
    s = po.RKSolver(po.RHS[rhs_id](params), po.TABLEAUX[tab_name], t0, tf, x0, h)
    if no_adaptive:
        s.no_adaptive()
    if rtol is not None:
        s.with_tolerance(rtol, rtol)
    s.norm_kind = norm
    if t_list is not None:
        s.t_list = list(t_list)
    s.run(adaptive=adaptive)
    return s


def both(name, rhs, rhs_id, params, tab_name, t0, tf, x0, h, **kw):
    tab = ol.builtin_tableau(ol.TABLEAU_ID[tab_name])
    okw = dict(kw)
    if "rtol" in okw and okw["rtol"] is not None:
        okw["atol"] = okw["rtol"]
    else:
        okw.pop("rtol", None)
    norm = okw.pop("norm", 0)
    cx, co, _ = ol.rk_solve(rhs, params, tab, t0, tf, x0, h, norm=norm, **okw)
    ps = py_solve(rhs_id, params, tab_name, t0, tf, x0, h, norm=norm, **kw)
    assert list(cx) == ps.x, (name, list(cx), ps.x)
    assert (co.n_accept, co.n_reject, co.t, co.h) == (ps.n_accept, ps.n_reject, ps.t, ps.h), name
    return dict(x=[float(v) for v in cx], t=co.t, h=co.h, accepted=co.n_accept, rejected=co.n_reject, calls=co.n_calls)


def main():
    g = {}
    g["test_rk45_2"] = both("rk45_2", "DIAG_LINEAR", 0, [-1.0, -2.0], "RKF45_REF", 0.0, 2.0, [1.0, 1.0], 1e-4)
    g["test_rk45_2_no_adaptive"] = both("rk45_2na", "DIAG_LINEAR", 0, [-1.0, -2.0], "RKF45_REF", 0.0, 2.0, [1.0, 1.0], 1e-4,
                                        no_adaptive=True)
    g["test_rk45_f64"] = both("rk45_f64", "DIAG_LINEAR", 0, [-1.0], "RKF45_REF", 0.0, 2.0, [1.0], 1e-4, adaptive=True, rtol=1e-10)
    for rtol in (1e-6, 1e-8, 1e-10):
        g[f"harmonic_rtol_{rtol:g}"] = both("harm", "HARMONIC2D", 1, [1.0], "RKF45_REF", 0.0, 10.0, [1.0, 0.0], 1e-3, adaptive=True,
                                            rtol=rtol)
    for mu in (0.5, 5.0, 20.0):
        for tab in ("DOPRI5", "RKF45_REF"):
            g[f"vdp_mu{mu:g}_{tab}"] = both("vdp", "VDP", 3, [mu], tab, 0.0, 20.0, [2.0, 0.0], 1e-3, adaptive=True, rtol=1e-6)
    g["lorenz_rk4_100"] = both("lorenz", "LORENZ63", 2, [10.0, 28.0, 8.0 / 3.0], "RK4", 0.0, 0.1, [1.0, 1.0, 1.0], 1e-3)
    g["harmonic_tlist"] = both("tl", "HARMONIC2D", 1, [1.0], "RKF45_REF", 0.0, 1.0, [1.0, 0.0], 0.01, t_list=[0.0, 0.35, 0.5, 1.0])
    # the complex flavour of the reference's first test (RK45ComplexSolver, nalgebra.rs:52-70): C++ complex arithmetic
    tab = ol.builtin_tableau(0)
    z, co = ol.rk_solve_c64([-1.0, -2.0], tab, 0.0, 2.0, [1.0 + 0j, 1.0 + 0j], 1e-4)
    g["test_rk45_1_complex"] = dict(x=[z[0].real, z[0].imag, z[1].real, z[1].imag], accepted=co.n_accept)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rk_known_answers.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", out)
    for k, v in g.items():
        print(k, v)


if __name__ == "__main__":
    main()
