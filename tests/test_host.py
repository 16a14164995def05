"""Host-side logic that needs no GPU: synthetic workload generators, sharding, the import shim."""
import numpy as np

from oracle import vecode_oracle as po


def test_splitmix_matches_python_oracle(vo):
    w = vo.workloads
    idx = np.array([0, 1, 2, 12345, 2**40 + 7], dtype=np.uint64)
    assert [int(v) for v in w.splitmix64(42, idx)] == [po.splitmix64(42, int(i)) for i in idx]
    assert list(w.uniform01(7, idx)) == [po.uniform01(7, int(i)) for i in idx]


def test_lorenz_x0_is_shard_invariant(vo):
    w = vo.workloads
    full = w.lorenz_x0(1000)
    parts = []
    for r in range(3):
        lo, hi = w.shard_range(1000, r, 3)
        parts.append(w.lorenz_x0(hi - lo, first=lo))
    assert np.array_equal(np.concatenate(parts), full)
    assert np.all(np.abs(full - 1.0) <= 1e-3)


def test_vdp_mu_sweep_and_shards(vo):
    w = vo.workloads
    mu = w.vdp_mu(1001)
    assert mu[0] == 0.5 and mu[-1] == 20.0
    lo, hi = w.shard_range(1001, 1, 4)
    assert np.array_equal(w.vdp_mu(1001, hi - lo, lo), mu[lo:hi])


def test_shard_range_covers_everything(vo):
    for n, g in [(10**6, 8), (7, 8), (100001, 4), (1, 1)]:
        spans = [vo.workloads.shard_range(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))


def test_schrodinger_system_is_hermitian(vo):
    H0, H1 = vo.workloads.schrodinger_system(64)
    assert np.array_equal(H0, H0.conj().T) and np.allclose(H1, H1.conj().T, atol=0)
    gp = vo.workloads.schrodinger_drive(100)
    assert gp.shape == (100, 1, 3) and np.all(gp[:, 0, 0] >= 0.5) and np.all(gp[:, 0, 1] < 3.0)
