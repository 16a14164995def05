"""pipeline.ChunkedSolve: the ensemble in independent chunks on their own streams and host threads. Trajectories do not
interact, so the chunked solve must reproduce the single-solver result bit for bit (fixed-step and adaptive)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("parts", [1, 3, 4])
def test_chunked_lorenz_fixed_step_bitwise(vo, ctx, parts):
    n = 5000
    x0 = vo.workloads.lorenz_x0(n)
    tab = vo.ButcherTableu.builtin("RK4")
    s = vo.RK45Solver(vo.Rhs(ctx, "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS)), 0.0, 0.1, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=tab)
    st_ref = s.run()
    ref = s.current()[1].to_host()

    def make(c, lo, hi, e):
        return vo.RK45Solver(vo.Rhs(c, "LORENZ63", 3, list(vo.workloads.LORENZ_PARAMS)), 0.0, 0.1, e, 1e-3, tableau=tab)
    cs = vo.pipeline.ChunkedSolve(0, "strict", n, 3, make, parts=parts)
    out = np.empty_like(x0)
    for _ in range(2):  # a second solve reuses the chunks' solvers through reset()
        sts = cs.solve(x0, out)
        assert all(st.kind == "Done" for st in sts) and sum(st.counts["Step"] for st in sts) == st_ref.counts["Step"]
        assert np.array_equal(out, ref)
    assert cs.launch_count > 0
    cs.close()


def test_chunked_vdp_adaptive_bitwise(vo, ctx):
    n = 4096
    mu, x0 = vo.workloads.vdp_mu(n), vo.workloads.vdp_x0(n)
    tab = vo.ButcherTableu.builtin("DOPRI5")
    s = vo.RK45Solver(vo.Rhs(ctx, "VDP", 2, [mu]), 0.0, 3.0, vo.Ensemble.from_host(ctx, x0), 1e-3, tableau=tab).with_tolerance(1e-6, 1e-6)
    st = s.run(adaptive=True)
    ref = s.current()[1].to_host()

    def make(c, lo, hi, e):
        return vo.RK45Solver(vo.Rhs(c, "VDP", 2, [mu[lo:hi].copy()]), 0.0, 3.0, e, 1e-3, tableau=tab).with_tolerance(1e-6, 1e-6)
    cs = vo.pipeline.ChunkedSolve(0, "strict", n, 2, make, parts=4)
    out = np.empty_like(x0)
    sts = cs.solve(x0, out, adaptive=True)
    assert np.array_equal(out, ref)
    assert sum(x.counts["Step"] for x in sts) == st.counts["Step"] and sum(x.counts["Reject"] for x in sts) == st.counts["Reject"]
    cs.close()
