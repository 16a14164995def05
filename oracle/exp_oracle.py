"""Pure-Python restatement of hmunozb/vec-ode's exponential integrators (src/exp) on the shared-basis dense split.

TEST INFRASTRUCTURE ONLY — the second, independent oracle of the exponential path. It must agree BIT FOR BIT with the C++
restatement (oracle/vecode_oracle.cpp: BasisSplit, cfm_exp, orc_exp_ensemble); tests/test_oracle.py checks that. Nothing in
the product imports it.

PARITY UNPINNED against the compiled crate: no Rust toolchain here, and the reference contains no implementation of
`ExponentialSplit` at all (src/exp/mod.rs:11-54 are trait declarations), so what is pinned to the reference is the SCHEME —
nodes, weights and the order of the LinearCombination calls (cited per function) — while `exp` / `map_exp` are this
project's shared-basis split: L = sum_m coef[m] B_m, exp lazy, map_exp = scaled Taylor series.

Complex numbers are (re, im) pairs of Python floats and every product is written out as num-complex does it
(`(a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re)`), so no library complex multiply (and no FMA) can change a bit.
"""
from __future__ import annotations

import math

# src/dat/mod.rs:4, 67-80 (same literals)
C_GAUSS_LEGENDRE_4 = [0.21132486540518711775, 0.78867513459481288225]
CFM_R2_J1_GL = [[0.5, 0.5]]
CFM_R4_J2_GL = [[0.53867513459481288225, -0.038675134594812882255], [-0.038675134594812882255, 0.53867513459481288225]]
BLANES17_R4_J4 = [[0.2463347584748155, -0.0469610812011527, 0.0119511881315244],
                  [0.0622500005170514, 0.2691833034233750, -0.0427581693456134],
                  [-0.0427581693456134, 0.2691833034233750, 0.0622500005170514],
                  [0.0119511881315244, -0.0469610812011527, 0.2463347584748155]]


def cmul(a, b):
    return (a[0] * b[0] - a[1] * b[1], a[0] * b[1] + a[1] * b[0])


def cadd(a, b):
    return (a[0] + b[0], a[1] + b[1])


def csub(a, b):
    return (a[0] - b[0], a[1] - b[1])


class BasisSplit:
    """ExponentialSplit (src/exp/mod.rs:11-35) for L = sum_m coef[m] B_m: `basis` is [M][n][n] of (re, im) pairs."""

    def __init__(self, basis, taylor_deg: int = 0):
        self.B, self.M, self.n = basis, len(basis), len(basis[0])
        self.taylor_deg = taylor_deg
        self.b_norm1 = []
        for m in range(self.M):  # induced 1-norm: largest column sum of |entries|
            best = 0.0
            for c in range(self.n):
                col = 0.0
                for r in range(self.n):
                    col += math.hypot(self.B[m][r][c][0], self.B[m][r][c][1])
                best = max(best, col)
            self.b_norm1.append(best)

    def theta(self, coef):
        th = 0.0
        for m in range(self.M):
            th += math.hypot(coef[m][0], coef[m][1]) * self.b_norm1[m]
        return th

    @staticmethod
    def plan(theta):
        """sub-steps so that theta/sq <= 1, then the smallest degree with (theta/sq)^k/k! <= 2^-53"""
        s = int(math.ceil(theta)) if theta > 1.0 else 1
        th = theta / s
        term, k = 1.0, 0
        while k < 60:
            k += 1
            term = term * th / k
            if term <= 1.1102230246251565e-16:
                break
        return s, k

    def apply_L(self, coef, x):
        n = self.n
        y = [(0.0, 0.0)] * n
        for m in range(self.M):
            cm, Bm = coef[m], self.B[m]
            for r in range(n):
                a = (0.0, 0.0)
                row = Bm[r]
                for c in range(n):
                    a = cadd(a, cmul(row[c], x[c]))
                y[r] = cadd(y[r], cmul(cm, a))
        return y

    def map_exp(self, coef, x):
        """map_exp(&exp(L), &x) (src/exp/mod.rs:23-25)"""
        sq, deg = self.plan(self.theta(coef))
        if self.taylor_deg > 0:
            deg = self.taylor_deg
        inv = 1.0 / sq
        cs = [(c[0] * inv, c[1] * inv) for c in coef]
        cur = list(x)
        for _ in range(sq):
            acc, term = list(cur), list(cur)
            for k in range(1, deg + 1):
                w = self.apply_L(cs, term)
                ik = 1.0 / k
                for r in range(self.n):
                    term[r] = (w[r][0] * ik, w[r][1] * ik)
                    acc[r] = cadd(acc[r], term[r])
            cur = acc
        return cur


def cfm_exp(sp, x0, dt, m, a):
    """src/exp/cfm.rs:20-40: k = a[0]*m[0]; k += a[i]*m[i]; k *= dt; x1 = map_exp(exp(k), x0). Operators are coefficient vectors."""
    M = len(m[0])
    k = [cmul((a[0], 0.0), m[0][q]) for q in range(M)]            # scalar_multiply_to, :31
    for i in range(1, len(a)):                                    # :33-36
        ai = (a[i], 0.0)
        k = [cadd(k[q], cmul(ai, m[i][q])) for q in range(M)]
    cdt = (dt, 0.0)
    k = [cmul(e, cdt) for e in k]                                  # scale, :37
    return sp.map_exp(k, x0)                                      # :38-39


def cfm_general(sp, f, t, x0, dt, c, alpha, alpha_err=None):
    """src/exp/cfm.rs:43-100. `f(t_arr)` returns one coefficient vector per node. Returns (xf, x_err | None)."""
    if any(len(row) != len(c) for row in alpha):
        raise ValueError("split_cfm: Incompatible array dimensions")   # :63
    t_arr = [t + ci * dt for ci in c]                                   # :70
    va = f(t_arr)                                                       # :72
    x = cfm_exp(sp, x0, dt, va, alpha[0])                               # :74-75
    for i in range(1, len(alpha)):                                      # :76-80
        x = cfm_exp(sp, x, dt, va, alpha[i])
    xf, xe = x, None
    if alpha_err is not None:                                           # :83-97
        if len(alpha_err) > len(alpha) or any(len(row) != len(c) for row in alpha_err):
            raise ValueError("split_cfm: Incompatible array dimensions for alph_err")
        e = cfm_exp(sp, x0, dt, va, alpha_err[0])
        for i in range(1, len(alpha_err)):
            e = cfm_exp(sp, e, dt, va, alpha_err[i])
        xe = [csub(e[r], xf[r]) for r in range(len(xf))]
    return xf, xe


def midpoint(sp, f, t, x0, dt):
    """src/exp/magnus.rs:10-26: xf = exp(dt * L(t + dt/2)) x0"""
    l = f(t + dt * 0.5)
    cdt = (dt, 0.0)
    return sp.map_exp([cmul(e, cdt) for e in l], x0)


def magnus_42(sp, f, t, x0, dt, cs, want_err=True):
    """src/exp/magnus.rs:28-83 with the commutator expanded on the basis through the structure tensor cs[a][b][c]."""
    M = sp.M
    c_mid = 0.288675134594812882254574390251
    b1 = dt * 0.5
    b2 = dt * dt * -0.144337567297406441127287195125
    mid_t = t + b1
    l0, l1 = f(mid_t - c_mid * dt), f(mid_t + c_mid * dt)               # :42-52
    w2 = [(0.0, 0.0)] * M
    for a in range(M):                                                   # commutator(l0, l1), :55
        for b in range(M):
            ab = cmul(l0[a], l1[b])
            for c in range(M):
                sc = cs[a][b][c]
                if sc != 0.0:
                    w2[c] = cadd(w2[c], (ab[0] * sc, ab[1] * sc))
    w2 = [cmul(e, (b2, 0.0)) for e in w2]                                # :56
    w1 = [cadd(l0[q], l1[q]) for q in range(M)]                          # :59-60
    w1 = [cmul(e, (b1, 0.0)) for e in w1]                                # :61
    w = [cadd(w1[q], w2[q]) for q in range(M)]                           # :65-66
    xf = sp.map_exp(w, x0)                                               # :72, 75
    xe = None
    if want_err:                                                         # :76-79
        u1 = sp.map_exp(w1, x0)
        xe = [csub(u1[r], xf[r]) for r in range(len(xf))]
    return xf, xe


def gen_cos(gp_i, M_gen, M):
    """The generator family of the configs: L(t) = B_0 + sum_m amp_m cos(omega_m t + phase_m) B_m, padded with zeros to M."""
    def one(t):
        coef = [(0.0, 0.0)] * M
        coef[0] = (1.0, 0.0)
        for m in range(1, M_gen):
            g = gp_i[m - 1]
            coef[m] = (g[0] * math.cos(g[1] * t + g[2]), 0.0)
        return coef
    return one


def solve_fixed(scheme, sp, gp_i, M_gen, psi0, t0, h, n_steps, cs=None, tables=None):
    """n_steps fixed steps of size h from t0 (t accumulated like ODEData::advance, ode.rs:184-188); no remainder logic."""
    g = gen_cos(gp_i, M_gen, sp.M)
    x, t = list(psi0), t0
    for _ in range(n_steps):
        if scheme == "midpoint":
            x = midpoint(sp, g, t, x, h)
        elif scheme == "cfm4":
            x, _ = cfm_general(sp, lambda ts: [g(tt) for tt in ts], t, x, h, C_GAUSS_LEGENDRE_4, CFM_R4_J2_GL)
        elif scheme == "cfm_table":
            c, alpha, _ = tables
            x, _ = cfm_general(sp, lambda ts: [g(tt) for tt in ts], t, x, h, c, alpha)
        else:
            x, _ = magnus_42(sp, g, t, x, h, cs, want_err=False)
        t += h
    return x


def split_cfm(sp, a_idx, f, t, x0, dt, c, rho, sigma):
    """src/exp/split_exp.rs:568-609 with the two user splits SpA, SpB as the index sets `a_idx` / the rest of ONE shared basis:
    an operator of split A is a coefficient vector that vanishes outside a_idx (and likewise for B), so f's pair (va, vb) is
    the generator's coefficient vector masked to either side. B(sigma_0) A(rho_0) ... A(rho_{s-1}) B(sigma_s) (:601-608)."""
    if any(len(r) != len(c) for r in rho) or any(len(r) != len(c) for r in sigma) or len(sigma) != len(rho) + 1:
        raise ValueError("split_cfm: Incompatible array dimensions")   # :587-592
    t_arr = [t + ci * dt for ci in c]                                   # :597
    full = f(t_arr)                                                     # :599
    in_a = [m in a_idx for m in range(sp.M)]
    va = [[coef[m] if in_a[m] else (0.0, 0.0) for m in range(sp.M)] for coef in full]
    vb = [[(0.0, 0.0) if in_a[m] else coef[m] for m in range(sp.M)] for coef in full]
    x = list(x0)
    for i in range(len(rho)):                                           # :601-606
        x = cfm_exp(sp, x, dt, vb, sigma[i])
        x = cfm_exp(sp, x, dt, va, rho[i])
    return cfm_exp(sp, x, dt, vb, sigma[len(rho)])                      # :607-608
