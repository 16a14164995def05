"""CPU restatement of the composite exponential splits of hmunozb/vec-ode (src/exp/split_exp.rs) with dense matrix
exponentials from scipy — TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

PARITY UNPINNED against the compiled crate (no Rust toolchain; the reference has no test for src/exp). Each function
restates the order of `sp_a.map_exp` / `sp_b.map_exp` calls of the cited lines with explicit matrices
A = sum_m la[m] B_a[m], B = sum_m lb[m] B_b[m]; tables are the literals of src/dat/mod.rs:30-62.
"""
import numpy as np
from scipy.linalg import expm

RKN_O4_A = [0.209515106613362, -0.143851773179818, 0.434336666566456]                                    # dat/mod.rs:34-36
RKN_O4_B = [0.0792036964311957, 0.353172906049774, -0.0420650803577195, 0.21937695575349958]              # dat/mod.rs:38-40
TJ_O4_A = [0.32439640402017118298 + 0.13458627249080669679j, 0.35120719195965763405 - 0.26917254498161339358j]  # :46-49
TJ_O4_B = [0.16219820201008559149 + 0.06729313624540334839j, 0.33780179798991440851 - 0.06729313624540334839j]  # :51-54
SEMI_COMPLEX_O4_B = [0.1 - 1j / 30.0, 4.0 / 15.0 + 2j / 15.0, 4.0 / 15.0 - 1j / 5.0]                          # :59-62


def commutative(A, B, x):   # split_exp.rs:165-167: sp_b.map_exp(u.1, sp_a.map_exp(u.0, x))
    return expm(B) @ (expm(A) @ x)


def strang(A, B, x):        # :246-257: lb scaled by 1/2; y = A(B(x)); B(y)
    ub = expm(0.5 * B)
    return ub @ (expm(A) @ (ub @ x))


def semi_complex_o4(A, B, x):  # :343-382
    ua = expm(0.25 * A)
    ub = [expm(k * B) for k in SEMI_COMPLEX_O4_B]
    y1 = ua @ (ub[0] @ x)
    y2 = ua @ (ub[1] @ y1)
    y3 = ua @ (ub[2] @ y2)
    y4 = ua @ (ub[1] @ y3)
    return ub[0] @ y4


def triple_jump(A, B, x):   # :426-445
    ua, ub = [expm(k * A) for k in TJ_O4_A], [expm(k * B) for k in TJ_O4_B]
    y0 = ua[0] @ (ub[0] @ x)
    y1 = ua[1] @ (ub[1] @ y0)
    y2 = ua[0] @ (ub[1] @ y1)
    return ub[0] @ y2


def rknr4(A, B, x):         # :500-515
    ua, ub = [expm(k * A) for k in RKN_O4_A], [expm(k * B) for k in RKN_O4_B]
    y0 = ua[0] @ (ub[0] @ x)
    y1 = ua[1] @ (ub[1] @ y0)
    y2 = ua[2] @ (ub[2] @ y1)
    y3 = ua[2] @ (ub[3] @ y2)
    y4 = ua[1] @ (ub[2] @ y3)
    y5 = ua[0] @ (ub[1] @ y4)
    return ub[0] @ y5


def split_exp_midpoint_step(A_of_t, B_of_t, t, x, dt):
    """split_exp_midpoint (:520-562), literally: (la, lb) = f(t); UA0 = exp(dt/2 la); UB0 = exp(dt/2 lb); A B A."""
    ua, ub = expm(0.5 * dt * A_of_t(t)), expm(0.5 * dt * B_of_t(t))
    return ua @ (ub @ (ua @ x))


def split_cfm_step(A_of_t, B_of_t, t, x, dt, c, rho, sigma):
    """split_cfm (:568-609) with dense matrix exponentials: B(sigma_0) A(rho_0) ... A(rho_{s-1}) B(sigma_s), each exponent
    dt * sum_q w_q L(t + c_q dt) (cfm_exp, src/exp/cfm.rs:20-40)."""
    ts = [t + ci * dt for ci in c]
    As, Bs = [A_of_t(tt) for tt in ts], [B_of_t(tt) for tt in ts]

    def comb(mats, w):
        return dt * sum(wq * m for wq, m in zip(w, mats))
    y = x
    for i in range(len(rho)):
        y = expm(comb(Bs, sigma[i])) @ y
        y = expm(comb(As, rho[i])) @ y
    return expm(comb(Bs, sigma[len(rho)])) @ y
