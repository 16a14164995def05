"""50-digit reference for map_exp: exp(L) x as the plain Taylor series sum_k L^k x / k! in mpmath
arithmetic at 50 significant digits — no scaling, no squaring, no Pade, so nothing in common with the scaled series of the
oracle and of the kernels. Test infrastructure."""
import mpmath as mp
import numpy as np


def map_exp_mp(L: np.ndarray, x: np.ndarray, digits: int = 50) -> np.ndarray:
    n = L.shape[0]
    with mp.workdps(digits):
        Lm = [[mp.mpc(float(L[r, c].real), float(L[r, c].imag)) for c in range(n)] for r in range(n)]
        term = [mp.mpc(float(z.real), float(z.imag)) for z in x]
        acc = list(term)
        tol = mp.mpf(10) ** (-(digits - 2))
        for k in range(1, 400):
            term = [mp.fsum(Lm[r][c] * term[c] for c in range(n)) / k for r in range(n)]
            acc = [a + t for a, t in zip(acc, term)]
            if max(abs(t) for t in term) < tol:
                break
        else:
            raise RuntimeError("Taylor series did not converge")
        return np.array([complex(a) for a in acc])
