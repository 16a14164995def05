"""Exponential integrators (src/exp) on the B200 against the CPU oracle and an eigendecomposition reference.

The reference pins only the SCHEME (nodes, weights, composition order); exp/map_exp/commutator are user-supplied there,
so the arithmetic bar is floating-point: <= 1e-12 against (a) the oracle's scaled-Taylor restatement and (b)
V diag(e^{lambda}) V^-1 from numpy/LAPACK for constant generators, plus unitarity of the propagated state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _small_system(n):
    rng = np.random.default_rng(n)
    k = np.arange(n)
    H0 = np.diag((k - (n - 1) / 2) * 0.05) + np.diag(np.full(n - 1, 0.5), 1) + np.diag(np.full(n - 1, 0.5), -1)
    G = rng.uniform(-1, 1, (n, n)) + 1j * rng.uniform(-1, 1, (n, n))
    return H0.astype(complex), (G + G.conj().T) / (2 * np.sqrt(n))


def _system(vo, n, N, seed=3):
    H0, H1 = vo.workloads.schrodinger_system(n) if n == 64 else _small_system(n)
    gp = vo.workloads.schrodinger_drive(N)
    rng = np.random.default_rng(seed)
    psi0 = rng.standard_normal((N, n)) + 1j * rng.standard_normal((N, n))
    psi0 /= np.linalg.norm(psi0, axis=1, keepdims=True)
    return -1j * H0, -1j * H1, gp, psi0


@pytest.mark.parametrize("n", [16, 64])
def test_map_exp_against_eigendecomposition(vo, ctx, n):
    """map_exp(exp(L), x) == V diag(e^{lambda}) V^-1 x for L = c0 B0 + c1 B1 with per-system coefficients."""
    import torch
    N = 37  # ragged: not a multiple of the 16-system tile
    B0, B1, _, psi0 = _system(vo, n, N)
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    rng = np.random.default_rng(5)
    coef = np.zeros((N, 2), dtype=complex)
    coef[:, 0] = rng.uniform(0.05, 1.2, N)          # up to ||L||_1 ~ 3: exercises sub-stepping
    coef[:, 1] = rng.uniform(-0.3, 0.3, N) + 0.05j * rng.uniform(-1, 1, N)
    x = torch.from_numpy(psi0.view(np.float64).reshape(N, n, 2).copy()).cuda()
    y = torch.empty_like(x)
    sp.map_exp(sp.exp(coef), x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    got = y.cpu().numpy().reshape(N, n * 2).view(np.complex128)
    for i in range(N):
        L = coef[i, 0] * B0 + coef[i, 1] * B1
        lam, V = np.linalg.eig(L)
        ref = V @ (np.exp(lam) * np.linalg.solve(V, psi0[i]))
        assert np.abs(got[i] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (i, np.abs(got[i] - ref).max())


@pytest.mark.parametrize("n,M", [(16, 1), (16, 4), (24, 2), (24, 3), (32, 1), (32, 4), (40, 2), (48, 2), (48, 3), (64, 1), (64, 3)])
def test_map_exp_on_every_compiled_shape_family(vo, ctx, n, M):
    """The shared-basis split beyond config 5's (64, 2): every n that is a multiple of 8 from 16 to 64 with the M values compiled in for it
    (the whole basis has to fit one SM's shared memory) — map_exp against scipy expm per system, ragged batch."""
    import torch
    from scipy.linalg import expm
    N = 19
    rng = np.random.default_rng(100 * n + M)
    Bs = []
    for _ in range(M):
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        Bs.append(-1j * (G + G.conj().T) / (2.0 * np.sqrt(n)))
    basis = np.stack(Bs)
    coef = (rng.uniform(-0.6, 0.6, (N, M)) + 0.1j * rng.uniform(-1, 1, (N, M))) * rng.choice([0.2, 1.0, 2.5], (N, 1))
    psi = rng.standard_normal((N, n)) + 1j * rng.standard_normal((N, n))
    sp = vo.DenseBasisSplit(ctx, basis)
    x = torch.from_numpy(psi.view(np.float64).reshape(N, n, 2).copy()).cuda()
    y = torch.empty_like(x)
    sp.map_exp(sp.exp(coef), x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    got = y.cpu().numpy().reshape(N, n * 2).view(np.complex128)
    for i in range(N):
        ref = expm(np.einsum("m,mij->ij", coef[i], basis)) @ psi[i]
        assert np.abs(got[i] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (i, np.abs(got[i] - ref).max())
    with pytest.raises(vo.VecOdeError):  # (56, 2) is not compiled in: the error names the dense split
        sp56 = vo.DenseBasisSplit(ctx, np.zeros((2, 56, 56), dtype=complex))
        z = torch.zeros((1, 56, 2), dtype=torch.float64).cuda()
        sp56.map_exp(sp56.exp(np.ones((1, 2), dtype=complex)), z.data_ptr(), z.clone().data_ptr())


def test_map_exp_and_dense_exp_against_mpmath_50_digits(vo, ctx):
    """map_exp of the lazy split and the explicit U = exp(L) of the dense split (scaling and squaring) against
    exp(L) x summed as a plain Taylor series in 50-digit arithmetic (tests/_mp_expm.py), four systems of n = 16 with ||L||_1 from
    0.3 to 6 (no sub-step to six sub-steps / three squarings)."""
    import torch
    from _mp_expm import map_exp_mp
    n, N = 16, 4
    B0, B1, _, psi0 = _system(vo, n, N)
    coef = np.array([[0.1, 0.05 + 0.02j], [0.5, -0.2], [1.1, 0.3 - 0.05j], [2.4, 0.25]])
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    x = torch.from_numpy(psi0.view(np.float64).reshape(N, n, 2).copy()).cuda()
    y = torch.empty_like(x)
    sp.map_exp(sp.exp(coef), x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    got = y.cpu().numpy().reshape(N, n * 2).view(np.complex128)
    Ls = np.einsum("nm,mij->nij", coef, np.stack([B0, B1]))
    ref = np.stack([map_exp_mp(Ls[i], psi0[i]) for i in range(N)])
    assert np.abs(got - ref).max() <= 5e-15, np.abs(got - ref).max()
    ds = vo.DenseSplit(ctx, n, N)
    U = ds.to_host(ds.exp(ds.operator(Ls)))
    assert np.abs(np.einsum("nij,nj->ni", U, psi0) - ref).max() <= 2e-14
    y2 = torch.empty_like(x)
    ds.map_exp(ds.exp(ds.operator(Ls)), x.data_ptr(), y2.data_ptr())
    torch.cuda.synchronize()
    assert np.abs(y2.cpu().numpy().reshape(N, n * 2).view(np.complex128) - ref).max() <= 2e-14


@pytest.mark.parametrize("scheme,cls", [("midpoint", "MidpointExpLinearSolver"), ("cfm4", "ExpCFMSolver"), ("magnus42", "MagnusExpLinearSolver")])
@pytest.mark.parametrize("n", [16, 64])
def test_fixed_step_schemes_match_oracle(vo, ctx, oracle, scheme, cls, n):
    """Config 5 at oracle size: fixed step h = 0.1, 20 steps, every scheme; <= 1e-12 vs the oracle, unitarity preserved."""
    N = 40
    B0, B1, gp, psi0 = _system(vo, n, N)
    if scheme == "magnus42":
        basis, cs = vo.with_commutator_slot(B0, B1)
        sp = vo.DenseBasisSplit(ctx, basis, commutator_structure=cs)
    else:
        basis, cs = np.stack([B0, B1]), None
        sp = vo.DenseBasisSplit(ctx, basis)
    ref = oracle.exp_ensemble(scheme, basis, gp, psi0, 0.0, 2.0, 0.1, M_gen=2, cs=cs, no_adaptive=True, n_threads=8)
    s = getattr(vo, cls)(sp, gp, 0.0, 2.0, psi0, 0.1, M_gen=2).no_adaptive()
    st = s.run()
    assert st.kind == "Done" and st.counts["Step"] == int(ref["accepted"].sum()) and st.counts["Chkpt"] == N
    (tmin, tmax), psi = s.current()
    assert tmin == tmax == ref["t"][0]
    err = np.abs(psi - ref["psi"]).max()
    print(f"{scheme} n={n}: max |psi - oracle| = {err:.2e}")
    assert err <= 1e-12
    assert np.abs(np.linalg.norm(psi, axis=1) - 1.0).max() <= 1e-12


def test_cfm4_converges_to_fine_reference(vo, ctx):
    """Scheme order check independent of the oracle: CFM4 error falls ~16x when h halves (4th order)."""
    n, N = 16, 16
    B0, B1, gp, psi0 = _system(vo, n, N)
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))

    def solve(h):
        s = vo.ExpCFMSolver(sp, gp, 0.0, 1.6, psi0, h).no_adaptive()
        s.run()
        return s.current()[1]

    fine = solve(0.0125)
    e1, e2 = np.abs(solve(0.2) - fine).max(), np.abs(solve(0.1) - fine).max()
    print("cfm4 errors", e1, e2, "ratio", e1 / e2)
    assert 10.0 < e1 / e2 < 24.0


@pytest.mark.parametrize("scheme,cls", [("cfm4", "ExpCFMSolver"), ("magnus42", "MagnusExpLinearSolver")])
def test_adaptive_schemes_match_oracle(vo, ctx, oracle, scheme, cls):
    """Adaptive stepping with per-system step control: within rtol of the oracle at t_end, step counts close."""
    n, N, rtol = 16, 48, 1e-7
    B0, B1, gp, psi0 = _system(vo, n, N)
    if scheme == "magnus42":
        basis, cs = vo.with_commutator_slot(B0, B1)
        sp = vo.DenseBasisSplit(ctx, basis, commutator_structure=cs)
    else:
        basis, cs = np.stack([B0, B1]), None
        sp = vo.DenseBasisSplit(ctx, basis)
    ref = oracle.exp_ensemble(scheme, basis, gp, psi0, 0.0, 2.0, 0.05, M_gen=2, cs=cs, adaptive=True, no_adaptive=False, rtol=rtol, n_threads=8)
    s = getattr(vo, cls)(sp, gp, 0.0, 2.0, psi0, 0.05, M_gen=2).with_tolerance(rtol, rtol)
    st = s.run(adaptive=True)
    assert st.kind == "Done"
    psi = s.current()[1]
    stats = s.stats()
    print(f"{scheme}: accepted gpu/oracle {stats['accepted'].sum()}/{ref['accepted'].sum()} rejected {stats['rejected'].sum()}/{ref['rejected'].sum()}")
    assert np.abs(stats["t"] - 2.0).max() <= 1e-13
    assert np.abs(psi - ref["psi"]).max() <= 50 * rtol
    assert abs(int(stats["accepted"].sum()) - int(ref["accepted"].sum())) <= 0.02 * ref["accepted"].sum() + 2


def test_exp_error_behaviour(vo, ctx):
    B0, B1, gp, psi0 = _system(vo, 16, 4)
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    s = vo.ExpCFMSolver(sp, gp, 0.0, 1.0, psi0, 0.1).no_adaptive()
    with pytest.raises(vo.VecOdeError) as e:
        s.step_adaptive()
    assert e.value.code == vo._cabi.VO_ERR_NOT_ADAPTIVE
    with pytest.raises(vo.VecOdeError) as e:  # no Commutator supplied (neither a structure tensor nor the dense one): refused at the first step
        m = vo.MagnusExpLinearSolver(sp, gp, 0.0, 1.0, psi0, 0.1)
        m.step(), m.step()
    assert e.value.code == vo._cabi.VO_ERR_BAD_ARG
    with pytest.raises(vo.VecOdeError):
        s.with_tolerance(0.0, 1.0)
    first = vo.ExpCFMSolver(sp, gp, 0.0, 1.0, psi0, 0.1).step()
    assert first.counts["Chkpt"] == 4 and first.counts["Step"] == 0  # first call is the checkpoint at t0


def test_full_size_properties_schrodinger(vo, ctx):
    """Config 5 at full size (N = 1e5, n = 64, CFM4, h = 0.1): unitarity of every system after 10 steps, shard equivalence
    (two halves == whole, bit for bit), and agreement of a 16-system prefix with a standalone 16-system run."""
    N, n = 100_000, 64
    H0, H1 = vo.workloads.schrodinger_system(n)
    sp = vo.DenseBasisSplit(ctx, np.stack([-1j * H0, -1j * H1]))
    gp = vo.workloads.schrodinger_drive(N)
    psi0 = np.zeros((N, n), dtype=np.complex128)
    psi0[:, 0] = 1.0

    def run(lo, hi):
        s = vo.ExpCFMSolver(sp, gp[lo:hi], 0.0, 1.0, psi0[lo:hi], 0.1).no_adaptive()
        assert s.run().kind == "Done"
        return s.current()[1]

    full = run(0, N)
    assert np.abs(np.linalg.norm(full, axis=1) - 1.0).max() <= 1e-13
    half = 50_000
    assert np.array_equal(np.concatenate([run(0, half), run(half, N)]), full)
    assert np.array_equal(run(0, 16), full[:16])


def test_split_norm_and_commutator(vo, ctx):
    """NormedExponentialSplit::norm and Commutator::commutator of the shipped split (exp/mod.rs:37-54)."""
    n, N = 16, 50
    B0, B1, gp, psi0 = _system(vo, n, N)
    basis, cs = vo.with_commutator_slot(B0, B1)
    sp = vo.DenseBasisSplit(ctx, basis, commutator_structure=cs)
    s = vo.MagnusExpLinearSolver(sp, gp, 0.0, 0.5, 1.7 * psi0, 0.1, M_gen=2).no_adaptive()
    np.testing.assert_allclose(sp.norm(s.state_device_ptr, N), 1.7, rtol=1e-14)
    rng = np.random.default_rng(0)
    la = np.zeros((N, 3), complex); lb = np.zeros((N, 3), complex)
    la[:, :2] = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
    lb[:, :2] = rng.standard_normal((N, 2))
    c = sp.commutator(la, lb)
    for i in range(N):
        La, Lb = la[i, 0] * B0 + la[i, 1] * B1, lb[i, 0] * B0 + lb[i, 1] * B1
        ref = La @ Lb - Lb @ La
        got = c[i, 0] * basis[0] + c[i, 1] * basis[1] + c[i, 2] * basis[2]
        assert np.abs(got - ref).max() <= 1e-13


COS_BODY = "g[1] = p[0] * cos(p[1] * t + p[2]);"


@pytest.mark.parametrize("scheme,cls", [("cfm4", "ExpCFMSolver"), ("midpoint", "MidpointExpLinearSolver"), ("magnus42", "MagnusExpLinearSolver")])
def test_user_generator_restating_the_cosine_family_gives_the_builtin_bits(vo, ctx, scheme, cls):
    """vo_exp_set_generator: the generator closure (`FnMut(T) -> L`, exp/cfm.rs:54, exp/magnus.rs:12,32) compiled at run time
    into the same tensor-core kernel. The cosine drive written as a body must reproduce the compiled-in family bit for bit."""
    n, N = 64, 40
    B0, B1, gp, psi0 = _system(vo, n, N)
    if scheme == "magnus42":
        basis, cs = vo.with_commutator_slot(B0, B1)
        sp = vo.DenseBasisSplit(ctx, basis, commutator_structure=cs)
    else:
        sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    out = []
    for custom in (False, True):
        s = getattr(vo, cls)(sp, gp, 0.0, 1.0, psi0, 0.1, M_gen=2).no_adaptive()
        if custom:
            s.set_generator(COS_BODY)
        assert s.run().kind == "Done"
        out.append(s.current()[1])
    assert np.array_equal(out[0], out[1])


def test_user_generator_against_dense_expm_restatement_of_cfm4(vo, ctx):
    """A generator with no built-in counterpart, g(t) = a t exp(-b t) + c per system, against cfm_general (exp/cfm.rs:43-100)
    restated with dense matrix exponentials: nodes C_GAUSS_LEGENDRE_4, weights CFM_R4_J2_GL (dat/mod.rs:4, 71-74)."""
    from scipy.linalg import expm
    n, N, h, steps = 16, 20, 0.1, 10
    B0, B1, _, psi0 = _system(vo, n, N)
    rng = np.random.default_rng(2)
    gp = np.stack([rng.uniform(0.5, 2.0, N), rng.uniform(0.1, 1.0, N), rng.uniform(-0.5, 0.5, N)], axis=1)[:, None, :]
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    s = vo.ExpCFMSolver(sp, gp, 0.0, h * steps, psi0, h, M_gen=2).no_adaptive()
    s.set_generator("g[1] = p[0] * t * exp(-p[1] * t) + p[2];")
    assert s.run().kind == "Done"
    got = s.current()[1]
    c = (0.21132486540518711775, 0.78867513459481288225)
    a = ((0.53867513459481288225, -0.038675134594812882255), (-0.038675134594812882255, 0.53867513459481288225))
    for i in range(N):
        g = lambda t: gp[i, 0, 0] * t * np.exp(-gp[i, 0, 1] * t) + gp[i, 0, 2]
        x, t = psi0[i].copy(), 0.0
        for _ in range(steps):
            L = [B0 + g(t + cq * h) * B1 for cq in c]
            for row in a:  # cfm_exp: x <- exp(dt * sum_j alpha_ij L_j) x, row after row
                x = expm(h * (row[0] * L[0] + row[1] * L[1])) @ x
            t += h
        assert np.abs(got[i] - x).max() <= 1e-11, (i, np.abs(got[i] - x).max())
    with pytest.raises(vo.VecOdeError) as ei:
        s.set_generator("g[1] = undefined_symbol;")
    assert "generator_body(1)" in str(ei.value)


def test_grouping_systems_by_drive_amplitude_keeps_the_callers_order(vo, ctx, oracle):
    """group_similar=True reorders the systems on the device (tiles of similar ||L h|| share a Taylor plan with less waste);
    results come back in the caller's order and stay within 1e-12 of the oracle."""
    n, N = 16, 77
    B0, B1, gp, psi0 = _system(vo, n, N)
    basis = np.stack([B0, B1])
    sp = vo.DenseBasisSplit(ctx, basis)
    ref = oracle.exp_ensemble("cfm4", basis, gp, psi0, 0.0, 1.0, 0.1, M_gen=2, no_adaptive=True, n_threads=4)
    s = vo.ExpCFMSolver(sp, gp, 0.0, 1.0, psi0, 0.1, M_gen=2, group_similar=True).no_adaptive()
    assert s._perm is not None and not np.array_equal(s._perm, np.arange(N))
    assert s.run().kind == "Done"
    psi = s.current()[1]
    assert np.abs(psi - ref["psi"]).max() <= 1e-12
    st = s.stats()
    assert np.array_equal(st["accepted"], ref["accepted"])
    s.reset(psi0)
    assert s.run().kind == "Done" and np.array_equal(s.current()[1], psi)


# ---- round 2: cfm_general with run-time tables, split_cfm, dense per-system operators ---------------------------------------
GL6_NODES = [0.5 - np.sqrt(15.0) / 10.0, 0.5, 0.5 + np.sqrt(15.0) / 10.0]
GL6_WEIGHTS = [[5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0]]  # exp(dt sum_q w_q L(t + c_q dt)): the 2nd-order exponential, as the embedded row


@pytest.mark.parametrize("n", [16, 64])
def test_cfm_general_with_tables_matches_oracle(vo, ctx, oracle, n):
    """cfm_general (exp/cfm.rs:43-100) with the caller's nodes and weights: BLANES17_R4_J4 (dat/mod.rs:76-80) on the three
    6th-order Gauss-Legendre nodes, 4 exponentials per step; fixed steps <= 1e-12 against the oracle's cfm_general, and an
    adaptive run with a one-exponential embedded row follows the oracle's step sequence."""
    N = 24
    B0, B1, gp, psi0 = _system(vo, n, N)
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    alpha = vo.cfm_table("BLANES17_R4_J4")
    assert alpha.shape == (4, 3) and np.array_equal(vo.cfm_table("CFM_R4_J2_GL"), [[0.53867513459481288225, -0.038675134594812882255], [-0.038675134594812882255, 0.53867513459481288225]])
    s = vo.ExpCFMGeneralSolver(sp, gp, 0.0, 2.0, psi0, 0.1, GL6_NODES, alpha)
    st = s.run()
    assert st.kind == "Done"
    ref = oracle.exp_ensemble("cfm_table", np.stack([B0, B1]), gp, psi0, 0.0, 2.0, 0.1, no_adaptive=True, tables=(GL6_NODES, alpha, None))
    _, psi = s.current()
    assert np.abs(psi - ref["psi"]).max() <= 1e-12 and np.array_equal(s.stats()["accepted"], ref["accepted"])
    assert np.abs(np.linalg.norm(psi, axis=1) - 1.0).max() <= 1e-12
    # the same tables through CFM4's own path: ExpCFMSolver is cfm_general with C_GAUSS_LEGENDRE_4 / CFM_R4_J2_GL / CFM_R2_J1_GL
    a = vo.ExpCFMSolver(sp, gp, 0.0, 1.0, psi0, 0.1).no_adaptive()
    b = vo.ExpCFMGeneralSolver(sp, gp, 0.0, 1.0, psi0, 0.1, vo.cfm_table("C_GAUSS_LEGENDRE_4")[0], vo.cfm_table("CFM_R4_J2_GL"))
    a.run(), b.run()
    assert np.array_equal(a.current()[1].view(np.float64), b.current()[1].view(np.float64))
    # adaptive
    s = vo.ExpCFMGeneralSolver(sp, gp, 0.0, 2.0, psi0, 0.05, GL6_NODES, alpha, alph_err=GL6_WEIGHTS).with_tolerance(1e-6, 1e-6)
    st = s.run(adaptive=True)
    ref = oracle.exp_ensemble("cfm_table", np.stack([B0, B1]), gp, psi0, 0.0, 2.0, 0.05, adaptive=True, no_adaptive=False, rtol=1e-6,
                              tables=(GL6_NODES, alpha, GL6_WEIGHTS))
    stats = s.stats()
    assert st.kind == "Done" and np.all(np.abs(stats["t"] - 2.0) <= 1e-13)
    assert np.array_equal(stats["accepted"], ref["accepted"]) and np.array_equal(stats["rejected"], ref["rejected"])
    assert np.abs(s.current()[1] - ref["psi"]).max() <= 1e-10
    with pytest.raises(vo.VecOdeError) as e:
        vo.ExpCFMGeneralSolver(sp, gp, 0.0, 1.0, psi0, 0.1, GL6_NODES, alpha[:, :2])
    assert "Incompatible array dimensions" in e.value.msg


def test_split_cfm_solver_matches_restatements(vo, ctx):
    """split_cfm (exp/split_exp.rs:568-609): B(sigma_0) A(rho_0) B(sigma_1) A(rho_1) B(sigma_2) with commutator-free exponents
    on the two Gauss-Legendre nodes. Against the pure-Python restatement of the same operations (scaled Taylor series, <= 1e-12)
    and against dense matrix exponentials (scipy expm)."""
    from oracle import exp_oracle as eo
    from oracle import split_oracle as so
    n, N, h, steps = 16, 5, 0.07, 3
    B0, B1, gp, psi0 = _system(vo, n, N)
    basis = np.stack([B0, B1])
    sp = vo.DenseBasisSplit(ctx, basis)
    c = list(vo.cfm_table("C_GAUSS_LEGENDRE_4")[0])
    # a 2-stage BAB composition whose A-weights and B-weights each sum to one over the step (consistency), asymmetric in the nodes
    rho = [[0.30, 0.20], [0.20, 0.30]]
    sigma = [[0.15, 0.05], [0.35, 0.25], [0.05, 0.15]]
    s = vo.ExpSplitCFMSolver(sp, [0], gp, 0.0, steps * h, psi0, h, c, rho, sigma)
    st = s.run()
    assert st.kind == "Done" and np.all(s.stats()["accepted"] >= steps)
    got = s.current()[1]
    psp = eo.BasisSplit([[[(complex(z).real, complex(z).imag) for z in row] for row in B] for B in basis])
    for i in range(N):
        g = eo.gen_cos([tuple(r) for r in gp[i]], 2, 2)
        x, t = [(z.real, z.imag) for z in psi0[i]], 0.0
        y = psi0[i].copy()
        for k in range(int(s.stats()["accepted"][i])):
            dt = h if k < steps else steps * h - t  # the remainder step of the driver loop, if any
            x = eo.split_cfm(psp, [0], lambda ts: [g(tt) for tt in ts], t, x, dt, c, rho, sigma)
            y = so.split_cfm_step(lambda tt: B0, lambda tt: gp[i, 0, 0] * np.cos(gp[i, 0, 1] * tt + gp[i, 0, 2]) * B1, t, y, dt, c, rho, sigma)
            t += dt
        ref = np.array([complex(a, b) for a, b in x])
        assert np.abs(got[i] - ref).max() <= 1e-12 and np.abs(got[i] - y).max() <= 1e-11
    with pytest.raises(vo.VecOdeError):
        vo.ExpSplitCFMSolver(sp, [0], gp, 0.0, 1.0, psi0, h, c, rho, sigma[:2])
    with pytest.raises(vo.VecOdeError):
        s.step_adaptive()


@pytest.mark.parametrize("n", [8, 24, 64])
def test_dense_split_operator_algebra(vo, ctx, n):
    """ExponentialSplit / Commutator for general per-system dense operators (vo_split_dense_*): commutator against numpy,
    explicit U = exp(L) against scipy expm (<= 1e-12) with ||U U^dagger - I||_F <= 1e-13 for anti-Hermitian L, map_exp with the
    explicit U, multi_exp, lin_zero and LinearCombination on operator ensembles."""
    import torch
    from scipy.linalg import expm
    N = 7
    rng = np.random.default_rng(n)
    G = rng.standard_normal((N, n, n)) + 1j * rng.standard_normal((N, n, n))
    H = (G + np.conj(np.transpose(G, (0, 2, 1)))) / (2.0 * np.sqrt(n))
    scale = np.array([0.05, 0.3, 1.0, 2.5, 6.0, 0.7, 11.0])[:, None, None]   # ||L||_1 from << 1 to ~ 40: zero to six squarings
    La = -1j * H * scale
    Lb = rng.standard_normal((N, n, n)) + 1j * rng.standard_normal((N, n, n))
    sp = vo.DenseSplit(ctx, n, N)
    a, b = sp.operator(La), sp.operator(Lb)
    comm = sp.to_host(sp.commutator(a, b))
    ref = La @ Lb - Lb @ La
    assert np.abs(comm - ref).max() <= 1e-13 * np.abs(ref).max()
    U = sp.to_host(sp.exp(a))
    for i in range(N):
        R = expm(La[i])
        assert np.abs(U[i] - R).max() <= 1e-12, (i, np.abs(U[i] - R).max())
        assert np.linalg.norm(U[i] @ U[i].conj().T - np.eye(n)) <= 1e-13 * max(1.0, float(scale[i, 0, 0])), i
    psi = rng.standard_normal((N, n)) + 1j * rng.standard_normal((N, n))
    x = torch.from_numpy(psi.view(np.float64).copy()).cuda()
    y = torch.empty_like(x)
    u = sp.exp(a)
    sp.map_exp(u, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    got = y.cpu().numpy().view(np.complex128)
    assert np.abs(got - np.einsum("nij,nj->ni", U, psi)).max() <= 1e-13 * np.abs(psi).max() * n
    assert np.allclose(sp.norm(y.data_ptr()), np.linalg.norm(got, axis=1), rtol=1e-14)
    us = sp.multi_exp(a, [0.5, -1.0])
    assert np.abs(sp.to_host(us[0]) @ sp.to_host(us[0]) - U).max() <= 1e-12
    assert np.abs(sp.to_host(us[1]) @ U - np.eye(n)).max() <= 1e-11
    z = sp.lin_zero()
    assert not sp.to_host(z).any()
    vo.LinearCombination.add_scalar_mul(z, 2.0, a)      # LinearCombination on operators: z = 0 + 2 a
    vo.LinearCombination.delta(z, a)                    # z -= a
    assert np.array_equal(sp.to_host(z), La)
    basis = np.stack([La[0], Lb[0]])
    bs = vo.DenseBasisSplit(ctx, basis) if n in (16, 32, 64) else None
    if bs is not None:
        coef = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
        got_l = sp.to_host(sp.from_basis(bs, coef))
        assert np.abs(got_l - np.einsum("nm,mij->nij", coef, basis)).max() <= 1e-13 * np.abs(basis).max() * 4


def test_magnus_with_dense_commutator(vo, ctx, oracle):
    """magnus_42 (exp/magnus.rs:28-83) with commutator(l0, l1) formed densely per system. (i) On a two-matrix basis it must agree
    with the structure-tensor route (and so with the oracle) to 1e-12, fixed and adaptive. (ii) THREE generator matrices whose
    commutators leave their span — no structure tensor exists on that basis — against a dense restatement of magnus_42 with
    scipy expm."""
    from scipy.linalg import expm
    n, N, h = 16, 19, 0.1
    B0, B1, gp, psi0 = _system(vo, n, N)
    basis3, cs = vo.with_commutator_slot(B0, B1)
    a = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis3, commutator_structure=cs), gp, 0.0, 1.0, psi0, h, M_gen=2).no_adaptive()
    b = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1])), gp, 0.0, 1.0, psi0, h, dense_commutator=True).no_adaptive()
    c = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1])), gp, 0.0, 1.0, psi0, h, applied_commutator=True).no_adaptive()
    a.run(), b.run(), c.run()
    assert np.abs(a.current()[1] - b.current()[1]).max() <= 1e-12
    assert np.abs(a.current()[1] - c.current()[1]).max() <= 1e-12  # the commutator applied by products inside the Taylor series, never formed
    ref = oracle.exp_ensemble("magnus42", basis3, gp, psi0, 0.0, 1.0, h, M_gen=2, cs=cs, no_adaptive=True)
    assert np.abs(b.current()[1] - ref["psi"]).max() <= 1e-12 and np.abs(c.current()[1] - ref["psi"]).max() <= 1e-12
    a = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis3, commutator_structure=cs), gp, 0.0, 1.0, psi0, h, M_gen=2).with_tolerance(1e-7, 1e-7)
    b = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1])), gp, 0.0, 1.0, psi0, h, dense_commutator=True).with_tolerance(1e-7, 1e-7)
    c = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1])), gp, 0.0, 1.0, psi0, h, applied_commutator=True).with_tolerance(1e-7, 1e-7)
    a.run(adaptive=True), b.run(adaptive=True), c.run(adaptive=True)
    assert np.array_equal(a.stats()["accepted"], b.stats()["accepted"]) and np.array_equal(a.stats()["rejected"], b.stats()["rejected"])
    assert np.array_equal(a.stats()["accepted"], c.stats()["accepted"]) and np.array_equal(a.stats()["rejected"], c.stats()["rejected"])
    assert np.abs(a.current()[1] - b.current()[1]).max() <= 1e-10 and np.abs(a.current()[1] - c.current()[1]).max() <= 1e-10
    # (ii) three generators, not closed under commutation
    rng = np.random.default_rng(12)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B2 = -1j * (G + G.conj().T) / (2 * np.sqrt(n))
    gp3 = np.concatenate([gp, gp * np.array([0.6, 1.7, 1.0]) + np.array([0.0, 0.0, 0.4])], axis=1)  # [N][2][3]
    with pytest.raises(vo.VecOdeError):  # no Commutator at all on this split
        vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1, B2])), gp3, 0.0, 1.0, psi0, h).no_adaptive().step()
    s = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1, B2])), gp3, 0.0, 0.5, psi0, h, dense_commutator=True).no_adaptive()
    st = s.run()
    assert st.kind == "Done"
    got, steps = s.current()[1], s.stats()["accepted"]

    def L(i, t):
        return B0 + gp3[i, 0, 0] * np.cos(gp3[i, 0, 1] * t + gp3[i, 0, 2]) * B1 + gp3[i, 1, 0] * np.cos(gp3[i, 1, 1] * t + gp3[i, 1, 2]) * B2
    c_mid = 0.288675134594812882254574390251
    for i in range(N):
        x, t = psi0[i].copy(), 0.0
        for k in range(int(steps[i])):
            dt = h if k < 5 else 0.5 - t
            l0, l1 = L(i, t + dt / 2 - c_mid * dt), L(i, t + dt / 2 + c_mid * dt)
            om = (l0 + l1) * (dt / 2) + (l0 @ l1 - l1 @ l0) * (dt * dt * -0.144337567297406441127287195125)
            x, t = expm(om) @ x, t + dt
        assert np.abs(got[i] - x).max() <= 1e-12, (i, np.abs(got[i] - x).max())
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() <= 1e-12


def test_literal_magnus_norm_switch(vo, ctx, oracle):
    """MagnusExpLinearSolver::norm as the reference writes it (magnus.rs:274-276: the norm of adaptive_dat.dx, a clone of x0 that no
    step updates). With rtol <= ||x0|| = 1 every attempt is rejected until h reaches min_dt (flagged STUCK here; the reference loops
    forever); with rtol = 8 every attempt is accepted and h doubles (0.9 * 8^(1/3) = 1.8 -> clamp 2 is not reached: factor 1.8).
    Both must match the oracle's literal switch, and the default (off) stays the embedded-error controller."""
    n, N, h = 16, 9, 0.01
    B0, B1, gp, psi0 = _system(vo, n, N)
    basis3, cs = vo.with_commutator_slot(B0, B1)
    sp = vo.DenseBasisSplit(ctx, basis3, commutator_structure=cs)
    s = vo.MagnusExpLinearSolver(sp, gp, 0.0, 1.0, psi0, h, M_gen=2).with_tolerance(8.0, 8.0).literal_norm()
    st = s.run(adaptive=True)
    ref = oracle.exp_ensemble("magnus42", basis3, gp, psi0, 0.0, 1.0, h, M_gen=2, cs=cs, adaptive=True, no_adaptive=False, rtol=8.0, literal_norm=True)
    stats = s.stats()
    assert st.kind == "Done" and np.array_equal(stats["accepted"], ref["accepted"]) and not stats["rejected"].any() and not ref["rejected"].any()
    assert np.abs(stats["dx_norm"] - 1.0).max() <= 1e-15  # ||x0||, every attempt
    acc_literal = stats["accepted"].copy()
    assert np.abs(s.current()[1] - ref["psi"]).max() <= 1e-12
    # rtol below ||x0||: nothing but rejections
    s = vo.MagnusExpLinearSolver(sp, gp, 0.0, 1.0, psi0, h, M_gen=2).with_tolerance(1e-4, 1e-4).literal_norm()
    for _ in range(12):
        st = s.step_adaptive()
    stats = s.stats()
    ref = oracle.exp_ensemble("magnus42", basis3, gp, psi0, 0.0, 1.0, h, M_gen=2, cs=cs, adaptive=True, no_adaptive=False, rtol=1e-4, literal_norm=True, max_calls=12)
    assert not stats["accepted"].any() and np.array_equal(stats["rejected"], ref["rejected"]) and np.all(stats["rejected"] == 11)
    assert np.allclose(stats["h"], ref["h"], rtol=1e-15) and np.all(stats["t"] == 0.0)
    # the dense-commutator kernel honours the switch too
    d = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, np.stack([B0, B1])), gp, 0.0, 1.0, psi0, h, dense_commutator=True).with_tolerance(8.0, 8.0).literal_norm()
    d.run(adaptive=True)
    assert np.array_equal(d.stats()["accepted"], acc_literal) and not d.stats()["rejected"].any()
    cfm = vo.ExpCFMSolver(sp, gp, 0.0, 1.0, psi0, h, M_gen=2)
    assert vo._cabi.lib().vo_exp_set_literal_norm(cfm._h, 1) == vo._cabi.VO_ERR_STATE  # only the Magnus solver's norm() has the quirk


def test_exp_golden_fixtures_on_the_gpu(vo, ctx):
    """The committed known answers of the exponential integrators (tests/golden/exp_known_answers.json: both CPU restatements bit for
    bit, map_exp pinned to a 50-digit sum) against the kernels: <= 1e-13 on the states after the fixture's steps, 5e-15 on map_exp."""
    import importlib.util
    import json
    import os
    import torch
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location("make_golden_exp", os.path.join(here, "golden", "make_golden_exp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    G = json.load(open(os.path.join(here, "golden", "exp_known_answers.json")))
    s15 = np.sqrt(15.0) / 10.0
    for key, e in G.items():
        if key.startswith("_"):
            continue
        basis, gp, psi0 = mod.case(e["n"], e["M"], e["seed"])
        want = np.array(e["psi"]).reshape(psi0.shape[0], e["n"], 2).view(np.complex128)[..., 0]
        tf = e["h"] * e["steps"]
        if e["scheme"] == "magnus42":
            basis3, cs = vo.with_commutator_slot(basis[0], basis[1])
            s = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis3, commutator_structure=cs), gp, 0.0, 1.0e9, psi0, e["h"], M_gen=2).no_adaptive()
        elif e["scheme"] == "midpoint":
            s = vo.MidpointExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp, 0.0, 1.0e9, psi0, e["h"])
        elif e["scheme"] == "cfm_table":
            s = vo.ExpCFMGeneralSolver(vo.DenseBasisSplit(ctx, basis), gp, 0.0, 1.0e9, psi0, e["h"], [0.5 - s15, 0.5, 0.5 + s15], vo.cfm_table("BLANES17_R4_J4")).no_adaptive()
        else:
            s = vo.ExpCFMSolver(vo.DenseBasisSplit(ctx, basis), gp, 0.0, 1.0e9, psi0, e["h"]).no_adaptive()
        s.run(max_calls=e["steps"] + 1)
        got = s.current()[1]
        assert np.abs(got - want).max() <= 1e-13, (key, np.abs(got - want).max())
        # map_exp of the first system with the generator at t = 0 scaled by h, against the committed 50-digit answer
        M = basis.shape[0]
        coef = np.zeros((1, M), dtype=complex)
        coef[0, 0] = e["h"]
        for m in range(1, M):
            coef[0, m] = e["h"] * gp[0, m - 1, 0] * np.cos(gp[0, m - 1, 2])
        sp = vo.DenseBasisSplit(ctx, basis)
        x = torch.from_numpy(psi0[:1].view(np.float64).reshape(1, e["n"], 2).copy()).cuda()
        y = torch.empty_like(x)
        sp.map_exp(sp.exp(coef), x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        mp_ref = np.array(e["map_exp_mp"]).view(np.complex128)[:, 0]
        assert np.abs(y.cpu().numpy().reshape(e["n"] * 2).view(np.complex128) - mp_ref).max() <= 5e-15, key


@pytest.mark.parametrize("n", [16, 64])
def test_magnus_with_applied_commutator_on_generators_not_closed_under_commutation(vo, ctx, n):
    """Three generator matrices whose commutators leave their span, large steps (several sub-steps of the Taylor series): the
    commutator applied by products inside exp_step_kernel (vo_exp_set_applied_commutator) against the per-system dense commutator
    (vo_exp_set_dense_commutator) and against a dense restatement of magnus_42 with scipy expm; a run-time compiled generator too."""
    from scipy.linalg import expm
    N, h = 21, 0.35
    B0, B1, gp, psi0 = _system(vo, n, N)
    rng = np.random.default_rng(12)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B2 = -1j * (G + G.conj().T) / (2 * np.sqrt(n))
    gp3 = np.concatenate([gp, gp * np.array([0.6, 1.7, 1.0]) + np.array([0.0, 0.0, 0.4])], axis=1)
    basis = np.stack([B0, B1, B2])
    a = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp3, 0.0, 1.05, psi0, h, applied_commutator=True).no_adaptive()
    d = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp3, 0.0, 1.05, psi0, h, dense_commutator=True).no_adaptive()
    assert a.run().kind == "Done" and d.run().kind == "Done"
    got = a.current()[1]
    assert np.abs(got - d.current()[1]).max() <= 1e-12
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() <= 1e-12

    def L(i, t):
        return B0 + gp3[i, 0, 0] * np.cos(gp3[i, 0, 1] * t + gp3[i, 0, 2]) * B1 + gp3[i, 1, 0] * np.cos(gp3[i, 1, 1] * t + gp3[i, 1, 2]) * B2
    c_mid = 0.288675134594812882254574390251
    for i in range(0, N, 5):
        x, t = psi0[i].copy(), 0.0
        for dt in (h, h, h):
            l0, l1 = L(i, t + dt / 2 - c_mid * dt), L(i, t + dt / 2 + c_mid * dt)
            om = (l0 + l1) * (dt / 2) + (l0 @ l1 - l1 @ l0) * (dt * dt * -0.144337567297406441127287195125)
            x, t = expm(om) @ x, t + dt
        assert np.abs(got[i] - x).max() <= 1e-12, (i, np.abs(got[i] - x).max())
    # adaptive: same accept / reject decisions as the dense route
    a = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp3, 0.0, 1.0, psi0, 0.2, applied_commutator=True).with_tolerance(1e-6, 1e-6)
    d = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp3, 0.0, 1.0, psi0, 0.2, dense_commutator=True).with_tolerance(1e-6, 1e-6)
    a.run(adaptive=True), d.run(adaptive=True)
    assert np.array_equal(a.stats()["accepted"], d.stats()["accepted"]) and np.array_equal(a.stats()["rejected"], d.stats()["rejected"])
    assert np.abs(a.current()[1] - d.current()[1]).max() <= 1e-10
    if n == 16:  # with the user's generator compiled at run time
        body = "g[1] = p[0] * cos(p[1] * t + p[2]); g[2] = p[3] * cos(p[4] * t + p[5]);"
        u = vo.MagnusExpLinearSolver(vo.DenseBasisSplit(ctx, basis), gp3, 0.0, 1.05, psi0, h, applied_commutator=True).no_adaptive()
        u.set_generator(body)
        u.run()
        assert np.abs(u.current()[1] - got).max() <= 1e-13


@pytest.mark.parametrize("scheme,cls,kw", [("cfm4", "ExpCFMSolver", {}), ("magnus42", "MagnusExpLinearSolver", {"applied_commutator": True})])
def test_dynamic_grouping_changes_tiles_not_results(vo, ctx, scheme, cls, kw):
    """vo_exp_set_dynamic_grouping: the systems are sorted by the norm bound of their coming exponent before every event and the
    kernel forms its tiles in that order. Systems are independent, so states, times and counters must be those of the plain order
    (to rounding: a system may run a different Taylor degree in different company), in the caller's order, fixed-step and adaptive,
    with a ragged last tile and trajectories finishing at different times."""
    n, N, h = 16, 1003, 0.2
    B0, B1, gp, psi0 = _system(vo, n, N)
    gp = gp.copy()
    gp[:, 0, 0] *= 1.0 + 3.0 * (np.arange(N) % 7 == 0)  # a few strongly driven systems scattered around
    sp = vo.DenseBasisSplit(ctx, np.stack([B0, B1]))
    runs = []
    for dyn in (False, True):
        s = getattr(vo, cls)(sp, gp, 0.0, 1.0, psi0, h, **kw).no_adaptive()
        if dyn:
            s.dynamic_grouping()
        assert s.run().kind == "Done"
        a = getattr(vo, cls)(sp, gp, 0.0, 1.0, psi0, h, **kw).with_tolerance(1e-6, 1e-6)
        if dyn:
            a.dynamic_grouping()
        assert a.run(adaptive=True).kind == "Done"
        runs.append((s.current()[1], s.stats(), a.current()[1], a.stats()))
    assert np.abs(runs[0][0] - runs[1][0]).max() <= 1e-13 and np.array_equal(runs[0][1]["accepted"], runs[1][1]["accepted"])
    assert np.abs(runs[0][2] - runs[1][2]).max() <= 1e-9
    assert np.array_equal(runs[0][3]["accepted"], runs[1][3]["accepted"]) and np.array_equal(runs[0][3]["rejected"], runs[1][3]["rejected"])
    assert np.all(runs[1][3]["t"] == 1.0)
