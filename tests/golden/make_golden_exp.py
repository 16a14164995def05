"""Regenerates tests/golden/exp_known_answers.json: known answers of the exponential integrators (src/exp) from the two independent
restatements (oracle/vecode_oracle.cpp and oracle/exp_oracle.py), which must agree BIT FOR BIT before anything is written, and whose
map_exp is checked against a 50-digit mpmath sum on the way. Inputs come from a counter-based generator so that the fixture holds the
seeds, not the matrices. Run from the repository root:  python tests/golden/make_golden_exp.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import exp_oracle as eo  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402


def case(n, M, seed, N=3):
    """The inputs of a fixture entry (also used by the tests that read the fixture)."""
    rng = np.random.default_rng(seed)
    H = []
    for _ in range(M):
        g = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        H.append((g + g.conj().T) / (2.0 * np.sqrt(n)))
    basis = np.stack([-1j * h for h in H])
    gp = np.stack([[0.5 + rng.random(), 1.0 + 2.0 * rng.random(), rng.random()] for _ in range(M - 1)])[None].repeat(N, axis=0)
    gp = gp * (1.0 + 0.1 * np.arange(N))[:, None, None]
    psi0 = rng.standard_normal((N, n)) + 1j * rng.standard_normal((N, n))
    psi0 /= np.linalg.norm(psi0, axis=1)[:, None]
    return basis, gp, psi0


def py_split(basis):
    return eo.BasisSplit([[[(complex(z).real, complex(z).imag) for z in row] for row in B] for B in basis])


def main():
    from _mp_expm import map_exp_mp
    out = {"_doc": "psi: final states [system][component][re, im] after `steps` fixed steps of size h from t = 0; inputs: tests/golden/make_golden_exp.py::case(n, M, seed)"}
    s15 = np.sqrt(15.0) / 10.0
    for scheme, n, M, seed, h, steps in [("midpoint", 16, 2, 21, 0.2, 5), ("cfm4", 16, 2, 22, 0.25, 5), ("magnus42", 16, 2, 23, 0.2, 5),
                                         ("cfm_table", 16, 2, 24, 0.3, 4), ("cfm4", 32, 3, 25, 0.15, 3)]:
        basis, gp, psi0 = case(n, M, seed)
        M_gen, cs, tables = M, None, None
        if scheme == "magnus42":
            comm = basis[0] @ basis[1] - basis[1] @ basis[0]
            basis = np.concatenate([basis, comm[None]])
            cs = np.zeros((3, 3, 3))
            cs[0, 1, 2], cs[1, 0, 2] = 1.0, -1.0
        if scheme == "cfm_table":
            tables = ([0.5 - s15, 0.5, 0.5 + s15], eo.BLANES17_R4_J4, None)
        ref = ol.exp_ensemble(scheme, basis, gp, psi0, 0.0, 1.0e9, h, M_gen=M_gen, cs=cs, no_adaptive=True, max_calls=steps + 1, tables=tables)
        sp = py_split(basis)
        for i in range(psi0.shape[0]):
            x = eo.solve_fixed(scheme, sp, [tuple(r) for r in gp[i]], M_gen, [(z.real, z.imag) for z in psi0[i]], 0.0, h, steps,
                               cs=None if cs is None else cs.tolist(), tables=tables)
            got = np.array([complex(a, b) for a, b in x])
            assert np.array_equal(got.view(np.float64), ref["psi"][i].view(np.float64)), (scheme, n, i)  # C++ == pure Python, bit for bit
        # the one piece that is this project's own: map_exp against a 50-digit sum (first system, generator at t = 0 scaled by h)
        g0 = eo.gen_cos([tuple(r) for r in gp[0]], M_gen, basis.shape[0])(0.0)
        L = sum(complex(*g0[m]) * basis[m] for m in range(basis.shape[0])) * h
        mp_ref = map_exp_mp(L, psi0[0])
        me = np.array([complex(a, b) for a, b in sp.map_exp([(c[0] * h, c[1] * h) for c in g0], [(z.real, z.imag) for z in psi0[0]])])
        assert np.abs(me - mp_ref).max() <= 2e-15
        out[f"{scheme}_n{n}_M{M}"] = dict(scheme=scheme, n=n, M=M, seed=seed, h=h, steps=steps,
                                          psi=[[[float(z.real), float(z.imag)] for z in row] for row in ref["psi"]],
                                          map_exp_mp=[[float(z.real), float(z.imag)] for z in mp_ref])
    path = os.path.join(ROOT, "tests", "golden", "exp_known_answers.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
