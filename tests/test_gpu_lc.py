"""LinearCombination kernels (src/lc.rs:7-55) and norms against the CPU oracle — bit-exact in strict mode."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape)


@pytest.mark.parametrize("d,n", [(1, 1), (3, 1000), (2, 4097), (1, 1 << 20), (5, 33333)])
def test_lc_primitives_bit_exact(vo, ctx, oracle, d, n):
    LC = vo.LinearCombination
    a, b = _rand((n, d), 1), _rand((n, d), 2)
    k = 0.7310585786300049
    lib = oracle.lib()
    P = lambda x: x.ctypes.data_as(C.c_void_p)

    v = vo.Ensemble.from_host(ctx, a)
    LC.scale(v, k)
    ref = a.copy(); lib.orc_lc_scale(P(ref), C.c_double(k), C.c_int64(ref.size))
    assert np.array_equal(v.to_host(), ref)

    u = vo.Ensemble.from_host(ctx, b)
    t = vo.Ensemble(ctx, d, n)
    LC.scalar_multiply_to(u, k, t)
    ref = np.empty_like(b); lib.orc_lc_scalar_multiply_to(P(b), C.c_double(k), P(ref), C.c_int64(ref.size))
    assert np.array_equal(t.to_host(), ref)

    v = vo.Ensemble.from_host(ctx, a)
    LC.add_scalar_mul(v, k, u)
    ref = a.copy(); lib.orc_lc_add_scalar_mul(P(ref), C.c_double(k), P(b), C.c_int64(ref.size))
    assert np.array_equal(v.to_host(), ref)

    v = vo.Ensemble.from_host(ctx, a)
    LC.add_assign_ref(v, u)
    assert np.array_equal(v.to_host(), a + b)

    v = vo.Ensemble.from_host(ctx, a)
    LC.delta(v, u)
    assert np.array_equal(v.to_host(), a - b)


@pytest.mark.parametrize("nterms", [1, 2, 5, 6, 7, 8, 11, 16])
@pytest.mark.parametrize("d,n", [(3, 10007), (1, 65536)])
def test_linear_combination_fused_matches_chain(vo, ctx, oracle, nterms, d, n):
    """The one-pass n-term reducer must equal the reference's chain scalar_multiply_to + add_scalar_mul... (lc.rs:20-35),
    zero coefficients included."""
    LC = vo.LinearCombination
    vs = [_rand((n, d), 10 + j) for j in range(nterms)]
    ks = _rand(nterms, 99)
    if nterms > 2:
        ks[1] = 0.0
    ens = [vo.Ensemble.from_host(ctx, v) for v in vs]
    out = vo.Ensemble(ctx, d, n)
    LC.linear_combination(out, ens, ks)
    ref = ks[0] * vs[0]
    for j in range(1, nterms):
        ref = ref + (ks[j] * vs[j])
    assert np.array_equal(out.to_host(), ref)
    # fused stage argument (rk.rs:121-124)
    x0 = _rand((n, d), 7)
    LC.stage_combine(out, ens, ks, 0.0123, vo.Ensemble.from_host(ctx, x0))
    assert np.array_equal(out.to_host(), ref * 0.0123 + x0)


def test_linear_combination_errors(vo, ctx):
    LC = vo.LinearCombination
    v = vo.Ensemble(ctx, 2, 8)
    with pytest.raises(vo.VecOdeError) as e:
        LC.linear_combination(v, [], [])
    assert "cannot be empty" in e.value.msg  # lc.rs:21-23
    with pytest.raises(vo.VecOdeError):
        LC.add_assign_ref(v, vo.Ensemble(ctx, 3, 8))
    with pytest.raises(vo.VecOdeError):
        LC.linear_combination(v, [v], [1.0])  # aliasing


def test_layouts_roundtrip(vo, ctx):
    a = _rand((1234, 3), 3)
    e = vo.Ensemble.from_host(ctx, a, "aos")
    assert np.array_equal(e.to_host("aos"), a)
    assert np.array_equal(e.to_host("soa"), a.T)
    c = e.clone()
    assert np.array_equal(c.to_host(), a)


@pytest.mark.parametrize("kind", ["L2", "LINF", "L1", "HYPOT"])
def test_norm_small_d_bit_exact(vo, ctx, kind):
    d, n = 4, 5001
    a = _rand((n, d), 5)
    got = vo.Ensemble.from_host(ctx, a).norm(kind)
    if kind == "L2":
        acc = np.zeros(n)
        for c in range(d):
            acc = acc + a[:, c] * a[:, c]
        ref = np.sqrt(acc)
        assert np.array_equal(got, ref)
    elif kind == "LINF":
        assert np.array_equal(got, np.abs(a).max(axis=1))
    elif kind == "L1":
        acc = np.zeros(n)
        for c in range(d):
            acc = acc + np.abs(a[:, c])
        assert np.array_equal(got, acc)
    else:
        m0, m1 = np.hypot(a[:, 0], a[:, 1]), np.hypot(a[:, 2], a[:, 3])
        np.testing.assert_allclose(got, np.sqrt(m0 * m0 + m1 * m1), rtol=4e-16)


def test_norm_large_state(vo, ctx):
    d = (1 << 20) + 3
    a = _rand((1, d), 6)
    e = vo.Ensemble.from_host(ctx, a)
    np.testing.assert_allclose(e.norm("L2")[0], np.sqrt(np.sum(a * a)), rtol=1e-13)
    assert e.norm("LINF")[0] == np.abs(a).max()
    np.testing.assert_allclose(e.norm("L1")[0], np.abs(a).sum(), rtol=1e-13)


@pytest.mark.parametrize("d,n", [(1, 2), (1, 2002), (3, 4098), (1, 1 << 20), (2, 33334)])
def test_complex_lc_primitives_bit_exact(vo, ctx, oracle, d, n):
    """LinearCombination<Complex<f64>, V> (ndarray.rs:8-33 with complex elements): rows of interleaved (re, im) pairs, complex scalars,
    num-complex's product written out; strict arithmetic must give the bits of the C++ restatement."""
    LC = vo.ComplexLinearCombination
    a, b, c3 = _rand((d, n), 11), _rand((d, n), 12), _rand((d, n), 13)
    k = complex(0.7310585786300049, -1.2345678901234567)
    lib = oracle.lib()
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    nz = C.c_int64(a.size // 2)
    kr, ki = C.c_double(k.real), C.c_double(k.imag)

    v = vo.Ensemble.from_host(ctx, a, layout="soa")
    LC.scale(v, k)
    ref = a.copy(); lib.orc_lcz_scale(P(ref), kr, ki, nz)
    assert np.array_equal(v.to_host(layout="soa"), ref)
    z = (a.reshape(-1, 2)[:, 0] + 1j * a.reshape(-1, 2)[:, 1]) * k  # numpy's own complex product as a sanity bound
    assert np.allclose(ref.reshape(-1, 2)[:, 0] + 1j * ref.reshape(-1, 2)[:, 1], z, rtol=1e-15, atol=1e-15)

    u = vo.Ensemble.from_host(ctx, b, layout="soa")
    t = vo.Ensemble(ctx, d, n)
    LC.scalar_multiply_to(u, k, t)
    ref = np.empty_like(b); lib.orc_lcz_scalar_multiply_to(P(b), kr, ki, P(ref), nz)
    assert np.array_equal(t.to_host(layout="soa"), ref)

    v = vo.Ensemble.from_host(ctx, a, layout="soa")
    LC.add_scalar_mul(v, k, u)
    ref = a.copy(); lib.orc_lcz_add_scalar_mul(P(ref), kr, ki, P(b), nz)
    assert np.array_equal(v.to_host(layout="soa"), ref)

    # n-term reducer in one pass against the chain of the reference (lc.rs:20-35)
    w = vo.Ensemble.from_host(ctx, c3, layout="soa")
    ks = [k, complex(-0.25, 0.5), complex(0.0, 1.0)]
    LC.linear_combination(t, [v, u, w], ks)
    va = v.to_host(layout="soa")
    ptrs = (C.c_void_p * 3)(P(va), P(b), P(c3))
    kflat = np.array([[q.real, q.imag] for q in ks]).ravel()
    ref = np.empty_like(a); lib.orc_lcz_linear_combination(P(ref), ptrs, P(kflat), 3, nz)
    assert np.array_equal(t.to_host(layout="soa"), ref)

    # fast arithmetic (FMA): same values to rounding
    fast = vo.Context(0, arith="fast")
    v = vo.Ensemble.from_host(fast, a, layout="soa")
    LC.scale(v, k)
    ref = a.copy(); lib.orc_lcz_scale(P(ref), kr, ki, nz)
    assert np.allclose(v.to_host(layout="soa"), ref, rtol=0, atol=8e-16 * np.abs(a).max() * abs(k))
    fast.close()


def test_complex_lc_rejects_odd_rows(vo, ctx):
    v = vo.Ensemble.from_host(ctx, np.ones((1, 3)), layout="soa")
    with pytest.raises(Exception):
        vo.ComplexLinearCombination.scale(v, 1j)
