"""Synthetic inputs of BASELINE.json's configs. Counter-based splitmix64 so that every rank / the CPU
oracle / the GPU path generate identical values for a given (seed, index) with no shared state."""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed: int, index) -> np.ndarray:
    idx = np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, index) -> np.ndarray:
    return (splitmix64(seed, index) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def lorenz_x0(n: int, seed: int = 42, first: int = 0) -> np.ndarray:
    """Config 2: x0_i = (1,1,1) + 1e-3*(2u-1) per component; AoS [n][3]; trajectories first .. first+n-1."""
    idx = (np.arange(first, first + n, dtype=np.uint64)[:, None] * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :])
    return 1.0 + 1.0e-3 * (2.0 * uniform01(seed, idx) - 1.0)


LORENZ_PARAMS = (10.0, 28.0, 8.0 / 3.0)


def vdp_mu(n_total: int, n: int | None = None, first: int = 0) -> np.ndarray:
    """Config 3: mu_i = 0.5 + 19.5*i/(N-1) over the WHOLE ensemble of n_total; returns the slice [first, first+n)."""
    n = n_total if n is None else n
    i = np.arange(first, first + n, dtype=np.float64)
    return 0.5 + 19.5 * i / float(max(n_total - 1, 1))


def vdp_x0(n: int) -> np.ndarray:
    x = np.zeros((n, 2))
    x[:, 0] = 2.0
    return x


def heat_u0(d: int) -> np.ndarray:
    """Config 4: u0_j = sin(2 pi j/d) + 0.5 sin(14 pi j/d)."""
    j = np.arange(d, dtype=np.float64)
    return np.sin(2.0 * np.pi * j / d) + 0.5 * np.sin(14.0 * np.pi * j / d)


def heat_u0_at(j, d: int) -> np.ndarray:
    """Config 4's initial state at the grid indices `j` of a d-point grid (a slab of a domain-decomposed state)."""
    j = np.asarray(j, dtype=np.float64)
    return np.sin(2.0 * np.pi * j / d) + 0.5 * np.sin(14.0 * np.pi * j / d)


def schrodinger_system(n: int = 64, seed_h1: int = 7):
    """Config 5: H0 real symmetric tridiagonal (diag (k-31.5)*0.05, off-diag 0.5), H1 = (G+G^dagger)/(2 sqrt n)."""
    k = np.arange(n, dtype=np.float64)
    H0 = np.diag((k - (n - 1) / 2.0) * 0.05) + np.diag(np.full(n - 1, 0.5), 1) + np.diag(np.full(n - 1, 0.5), -1)
    idx = np.arange(n * n, dtype=np.uint64)
    G = (2.0 * uniform01(seed_h1, 2 * idx) - 1.0) + 1j * (2.0 * uniform01(seed_h1, 2 * idx + 1) - 1.0)
    G = G.reshape(n, n)
    H1 = (G + G.conj().T) / (2.0 * np.sqrt(n))
    return H0.astype(np.complex128), H1


def schrodinger_drive(n_total: int, n: int | None = None, first: int = 0, seed: int = 11) -> np.ndarray:
    """Config 5: H_i(t) = H0 + a_i cos(w_i t) H1, a_i = 0.5+u, w_i = 1+2u'. Returns gp [n][1][3] = (amp, omega, phase)."""
    n = n_total if n is None else n
    i = np.arange(first, first + n, dtype=np.uint64)
    gp = np.zeros((n, 1, 3))
    gp[:, 0, 0] = 0.5 + uniform01(seed, 2 * i)
    gp[:, 0, 1] = 1.0 + 2.0 * uniform01(seed, 2 * i + 1)
    return gp


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous ceil(N/G) trajectory ranges."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)
