"""Composite exponential splits (src/exp/split_exp.rs, row N1) on the B200 against dense-expm restatements."""
import numpy as np
import pytest

from oracle import split_oracle as so

pytestmark = pytest.mark.gpu


def _setup(vo, ctx, n=16, N=21, seed=2):
    rng = np.random.default_rng(seed)

    def herm():
        G = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
        return (G + G.conj().T) / (2 * np.sqrt(n))

    basis = np.stack([-1j * herm(), -1j * herm(), -1j * herm()])   # A = span{B0, B1}, B = span{B2}
    sp = vo.DenseBasisSplit(ctx, basis)
    psi = rng.standard_normal((N, n)) + 1j * rng.standard_normal((N, n))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    la = rng.uniform(0.1, 0.9, (N, 2)) + 0j
    lb = rng.uniform(0.1, 0.9, (N, 1)) + 0j
    return basis, sp, psi, la, lb


def _apply(vo, split, l, psi):
    import torch
    N, n = psi.shape
    x = torch.from_numpy(psi.view(np.float64).reshape(N, n, 2).copy()).cuda()
    y = torch.empty_like(x)
    split.map_exp(split.exp(l), x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    return y.cpu().numpy().reshape(N, 2 * n).view(np.complex128)


@pytest.mark.parametrize("name,ref", [("CommutativeExpSplit", so.commutative), ("StrangSplit", so.strang), ("SemiComplexO4ExpSplit", so.semi_complex_o4),
                                      ("TripleJumpExpSplit", so.triple_jump), ("RKNR4ExpSplit", so.rknr4)])
def test_composite_splits_match_dense_expm(vo, ctx, name, ref):
    basis, sp, psi, la, lb = _setup(vo, ctx)
    split = getattr(vo, name)(sp, [0, 1], [2])
    got = _apply(vo, split, vo.DirectSumL(la, lb), psi)
    for i in range(psi.shape[0]):
        A = la[i, 0] * basis[0] + la[i, 1] * basis[1]
        B = lb[i, 0] * basis[2]
        want = ref(A, B, psi[i])
        assert np.abs(got[i] - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), (name, i, np.abs(got[i] - want).max())


@pytest.mark.parametrize("name,order", [("StrangSplit", 2), ("SemiComplexO4ExpSplit", 4), ("TripleJumpExpSplit", 4), ("RKNR4ExpSplit", 4)])
def test_composite_split_orders(vo, ctx, name, order):
    """exp(dt(A+B)) is approximated to the advertised order: halving dt divides the local error by ~2^(order+1).
    (The RKN split is 4th order for this generic A, B too; its name only says it was optimised for RKN problems.)"""
    basis, sp, psi, la, lb = _setup(vo, ctx, N=8)
    split = getattr(vo, name)(sp, [0, 1], [2])
    from scipy.linalg import expm
    errs = []
    for dt in (0.4, 0.2):
        got = _apply(vo, split, vo.DirectSumL(la * dt, lb * dt), psi)
        e = 0.0
        for i in range(psi.shape[0]):
            L = dt * (la[i, 0] * basis[0] + la[i, 1] * basis[1] + lb[i, 0] * basis[2])
            e = max(e, np.abs(got[i] - expm(L) @ psi[i]).max())
        errs.append(e)
    ratio = errs[0] / errs[1]
    print(name, "local errors", errs, "ratio", ratio)
    assert 0.6 * 2 ** (order + 1) <= ratio <= 1.6 * 2 ** (order + 1)


def test_multi_exp_and_lin_zero(vo, ctx):
    basis, sp, psi, la, lb = _setup(vo, ctx, N=5)
    st = vo.StrangSplit(sp, [0, 1], [2])
    z = st.lin_zero(5)
    assert z.a.shape == (5, 2) and z.b.shape == (5, 1) and not z.a.any()
    us = st.multi_exp(vo.DirectSumL(la, lb), [0.5, 2.0])
    assert np.array_equal(us[0], st.exp(vo.DirectSumL(0.5 * la, 0.5 * lb))) and us[1].shape == (3, 5, 3)
    got = _apply(vo, st, z, psi)  # exp(0) = identity
    assert np.array_equal(got, psi)


def test_exp_split_midpoint_solver(vo, ctx):
    """ExpSplitMidpointSolver (split_exp.rs:613-685): the literal scheme (generator at t, both halves dt/2, A B A)."""
    n, N, h, steps = 16, 12, 0.05, 10
    rng = np.random.default_rng(4)
    basis, sp, psi, _, _ = _setup(vo, ctx, n=n, N=N)
    basis2 = basis[:2]
    sp2 = vo.DenseBasisSplit(ctx, basis2)
    gp = vo.workloads.schrodinger_drive(N)
    s = vo.ExpSplitMidpointSolver(sp2, [0], gp, 0.0, h * steps, psi, h)
    st = s.run()
    assert st.kind == "Done" and st.counts["Step"] >= N * steps
    got = s.current()[1]
    for i in range(N):
        amp, om, ph = gp[i, 0]
        x, t = psi[i].copy(), 0.0
        for _ in range(int(st.counts["Step"] // N)):
            dt = min(h, h * steps - t)
            x = so.split_exp_midpoint_step(lambda tt: basis2[0], lambda tt: amp * np.cos(om * tt + ph) * basis2[1], t, x, dt)
            t += dt
        assert np.abs(got[i] - x).max() <= 1e-12
    with pytest.raises(vo.VecOdeError):
        s.step_adaptive()
