"""ctypes front-end for oracle/libvecode_oracle.so (the C++ CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvecode_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vecode_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class SolveCfg(C.Structure):
    _fields_ = [("rhs_kind", C.c_int32), ("d", C.c_int32), ("s", C.c_int32), ("has_berr", C.c_int32),
                ("adaptive", C.c_int32), ("no_adaptive", C.c_int32), ("norm_kind", C.c_int32), ("n_tlist", C.c_int32),
                ("t0", C.c_double), ("tf", C.c_double), ("h0", C.c_double),
                ("rtol", C.c_double), ("atol", C.c_double), ("min_dt", C.c_double), ("max_dt", C.c_double),
                ("order", C.c_double), ("alpha", C.c_double), ("max_calls", C.c_int64)]


class SolveOut(C.Structure):
    _fields_ = [("t", C.c_double), ("h", C.c_double), ("prev_h", C.c_double), ("dx_norm", C.c_double),
                ("n_accept", C.c_int64), ("n_reject", C.c_int64), ("n_calls", C.c_int64), ("state", C.c_int32)]


class ExpCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("M", C.c_int32), ("scheme", C.c_int32), ("adaptive", C.c_int32),
                ("no_adaptive", C.c_int32), ("taylor_deg", C.c_int32),
                ("t0", C.c_double), ("tf", C.c_double), ("h0", C.c_double), ("rtol", C.c_double),
                ("min_dt", C.c_double), ("max_dt", C.c_double), ("order", C.c_double), ("alpha", C.c_double),
                ("max_calls", C.c_int64),
                ("n_nodes", C.c_int32), ("n_rows", C.c_int32), ("n_rows_err", C.c_int32), ("literal_norm", C.c_int32),
                ("c", C.c_void_p), ("alpha_tab", C.c_void_p), ("alpha_err_tab", C.c_void_p)]


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_hardware_threads.restype = C.c_int32
        _lib.orc_rhs_n_params.restype = C.c_int32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def builtin_tableau(which: int):
    """0 RKF45_REF (reference literal), 1 RK4, 2 DOPRI5. Returns (ac, b, b_err|None, s)."""
    ac = np.zeros(64)
    b = np.zeros(8)
    be = np.zeros(8)
    s = C.c_int32()
    he = C.c_int32()
    lib().orc_builtin_tableau(C.c_int32(which), _p(ac), _p(b), _p(be), C.byref(s), C.byref(he))
    n = s.value
    return ac[:n * n].copy(), b[:n].copy(), (be[:n].copy() if he.value else None), n


TABLEAU_ID = {"RKF45_REF": 0, "RK4": 1, "DOPRI5": 2}
RHS_ID = {"DIAG_LINEAR": 0, "HARMONIC2D": 1, "LORENZ63": 2, "VDP": 3, "HEAT1D": 4}
NORM_ID = {"L2": 0, "LINF": 1, "L1": 2, "HYPOT": 3}


def make_cfg(rhs, d, tableau, t0, tf, h0, adaptive=False, no_adaptive=False, norm="L2", rtol=1e-4, atol=1e-6,
             min_dt=1e-6, max_dt=1.0, order=3.0, alpha=0.9, max_calls=0, t_list=None):
    ac, b, be, s = tableau
    cfg = SolveCfg(RHS_ID[rhs] if isinstance(rhs, str) else rhs, d, s, 0 if be is None else 1, int(adaptive),
                   int(no_adaptive), NORM_ID[norm] if isinstance(norm, str) else norm,
                   0 if t_list is None else len(t_list), t0, tf, h0, rtol, atol, min_dt, max_dt, order, alpha, max_calls)
    return cfg


def rk_solve(rhs, params, tableau, t0, tf, x0, h0, trace_cap=0, t_list=None, **kw):
    """One trajectory through the reference's driver loop. Returns (x, SolveOut, trace|None)."""
    x = np.array(x0, dtype=np.float64).copy()
    ac, b, be, s = tableau
    cfg = make_cfg(rhs, x.size, tableau, t0, tf, h0, t_list=t_list, **kw)
    out = SolveOut()
    params = np.ascontiguousarray(params, dtype=np.float64)
    tl = None if t_list is None else np.ascontiguousarray(t_list, dtype=np.float64)
    trace = np.zeros((trace_cap, 4)) if trace_cap else None
    lib().orc_rk_solve(C.byref(cfg), _p(ac), _p(b), _p(be), _p(params), _p(tl), _p(x), C.byref(out), _p(trace),
                       C.c_int64(trace_cap))
    if trace is not None:
        trace = trace[:min(trace_cap, out.n_calls)]
    return x, out, trace


def rk_solve_c64(params, tableau, t0, tf, z0, h0, **kw):
    z = np.array(z0, dtype=np.complex128).copy()
    x = z.view(np.float64)
    ac, b, be, s = tableau
    cfg = make_cfg("DIAG_LINEAR", z.size, tableau, t0, tf, h0, **kw)
    out = SolveOut()
    params = np.ascontiguousarray(params, dtype=np.float64)
    lib().orc_rk_solve_c64(C.byref(cfg), _p(ac), _p(b), _p(be), _p(params), None, _p(x), C.byref(out))
    return z, out


def rk_ensemble(rhs, params, tableau, t0, tf, x0, h0, n_threads=1, h0_arr=None, t_list=None, **kw):
    """x0: AoS [N][d]; params: AoS [N][n_params]. Returns dict of per-trajectory results."""
    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    N, d = x.shape
    ac, b, be, s = tableau
    cfg = make_cfg(rhs, d, tableau, t0, tf, h0, t_list=t_list, **kw)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(N, -1)
    t = np.zeros(N)
    h = np.zeros(N)
    acc = np.zeros(N, dtype=np.int64)
    rej = np.zeros(N, dtype=np.int64)
    st = np.zeros(N, dtype=np.int32)
    tl = None if t_list is None else np.ascontiguousarray(t_list, dtype=np.float64)
    h0a = None if h0_arr is None else np.ascontiguousarray(h0_arr, dtype=np.float64)
    lib().orc_rk_ensemble(C.byref(cfg), _p(ac), _p(b), _p(be), _p(params), C.c_int32(params.shape[1]), _p(tl),
                          C.c_int64(N), _p(x), _p(h0a), _p(t), _p(h), _p(acc), _p(rej), _p(st), C.c_int32(n_threads))
    return dict(x=x, t=t, h=h, accepted=acc, rejected=rej, state=st)


def rk_step(rhs, params, tableau, t, dt, x0, want_err=True):
    """One bare rk_step. Returns (xf, x_err|None, K[s][d])."""
    ac, b, be, s = tableau
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    d = x0.size
    xf = np.zeros(d)
    xe = np.zeros(d) if (want_err and be is not None) else None
    K = np.zeros((s, d))
    params = np.ascontiguousarray(params, dtype=np.float64)
    lib().orc_rk_step(C.c_int32(RHS_ID[rhs]), C.c_int32(d), _p(params), _p(ac), _p(b), _p(be), C.c_int32(s),
                      C.c_double(t), C.c_double(dt), _p(x0), _p(xf), _p(xe), _p(K))
    return xf, xe, K


def exp_ensemble(scheme, basis, gp, psi0, t0, tf, h0, M_gen=None, cs=None, adaptive=False, no_adaptive=True,
                 taylor_deg=0, rtol=1e-4, min_dt=1e-6, max_dt=1.0, order=3.0, alpha=0.9, max_calls=0, n_threads=1, tables=None, literal_norm=False):
    """basis: complex [M][n][n] (already -i*H_m); gp: [N][M_gen-1][3]; psi0: complex [N][n].
    scheme "cfm_table": cfm_general (cfm.rs:43-100) with tables = (c[k], alpha[s][k], alpha_err[s_err][k] | None)."""
    basis = np.ascontiguousarray(basis, dtype=np.complex128)
    M, n, _ = basis.shape
    M_gen = M if M_gen is None else M_gen
    psi = np.ascontiguousarray(psi0, dtype=np.complex128).copy()
    N = psi.shape[0]
    gp = np.ascontiguousarray(gp, dtype=np.float64).reshape(N, max(M_gen - 1, 0), 3)
    cfg = ExpCfg(n, M, {"midpoint": 0, "cfm4": 1, "magnus42": 2, "cfm_table": 3}[scheme], int(adaptive), int(no_adaptive), taylor_deg,
                 t0, tf, h0, rtol, min_dt, max_dt, order, alpha, max_calls)
    cfg.literal_norm = int(literal_norm)  # magnus.rs:274-276 as written
    keep = []
    if scheme == "cfm_table":
        c_t, a_t, e_t = tables
        c_a = np.ascontiguousarray(c_t, dtype=np.float64)
        a_a = np.ascontiguousarray(a_t, dtype=np.float64).reshape(-1, c_a.size)
        e_a = None if e_t is None else np.ascontiguousarray(e_t, dtype=np.float64).reshape(-1, c_a.size)
        keep = [c_a, a_a, e_a]
        cfg.n_nodes, cfg.n_rows, cfg.n_rows_err = c_a.size, a_a.shape[0], 0 if e_a is None else e_a.shape[0]
        cfg.c, cfg.alpha_tab, cfg.alpha_err_tab = c_a.ctypes.data, a_a.ctypes.data, None if e_a is None else e_a.ctypes.data
    t = np.zeros(N)
    h = np.zeros(N)
    acc = np.zeros(N, dtype=np.int64)
    rej = np.zeros(N, dtype=np.int64)
    csa = None if cs is None else np.ascontiguousarray(cs, dtype=np.float64)
    lib().orc_exp_ensemble(C.byref(cfg), _p(basis.view(np.float64)), _p(gp), C.c_int32(M_gen), _p(csa), C.c_int64(N),
                           _p(psi.view(np.float64)), _p(t), _p(h), _p(acc), _p(rej), C.c_int32(n_threads))
    return dict(psi=psi, t=t, h=h, accepted=acc, rejected=rej)


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())
