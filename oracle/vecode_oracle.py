"""Pure-Python restatement of hmunozb/vec-ode's Runge-Kutta stepping path.

TEST INFRASTRUCTURE ONLY — the second, independent oracle. It must agree BIT FOR BIT with
oracle/vecode_oracle.cpp (tests/test_oracle.py checks that); nothing in the product imports it.

PARITY UNPINNED: no Rust toolchain here and the crate's tests assert nothing, so this follows the
reference SOURCE (cited per function, paths relative to /root/reference), not its compiled output.
Python floats are IEEE binary64 and CPython never contracts a*b+c into an FMA, which is the same
arithmetic rustc emits by default.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.pow.restype = ctypes.c_double
_libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
_libm.hypot.restype = ctypes.c_double
_libm.hypot.argtypes = [ctypes.c_double, ctypes.c_double]

EPS = 2.220446049250313e-16  # f64::EPSILON

# --- src/dat/mod.rs:9-27, as division expressions (typo `2526` at :19 kept) ------------------------
RK45_AC = [
    0., 0., 0., 0., 0., 0.,
    1. / 4., 1. / 4., 0., 0., 0., 0.,
    3.0 / 32., 9.0 / 32., 3. / 8., 0., 0., 0.,
    1932. / 2197., -7200. / 2197., 7296. / 2197., 12. / 13., 0., 0.,
    439. / 216., -8., 3680. / 513., -845. / 4104., 1.0, 0.,
    -8. / 27., 2., -3544. / 2526., 1859. / 4104., -11. / 40., 1.0 / 2.0]
RK45_B = [16. / 135., 0., 6656. / 12825., 28561. / 56430., -9. / 50., 2. / 55.]
RK45_BERR = [25. / 216., 0., 1408. / 2565., 2197. / 4104., -1. / 5., 0.]

RK4_AC = [0., 0., 0., 0., 1. / 2., 1. / 2., 0., 0., 0., 1. / 2., 1. / 2., 0., 0., 0., 1., 1.]
RK4_B = [1. / 6., 1. / 3., 1. / 3., 1. / 6.]

DOPRI5_AC = [
    0., 0., 0., 0., 0., 0., 0.,
    1. / 5., 1. / 5., 0., 0., 0., 0., 0.,
    3. / 40., 9. / 40., 3. / 10., 0., 0., 0., 0.,
    44. / 45., -56. / 15., 32. / 9., 4. / 5., 0., 0., 0.,
    19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 8. / 9., 0., 0.,
    9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 1., 0.,
    35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 1.]
DOPRI5_B = [5179. / 57600., 0., 7571. / 16695., 393. / 640., -92097. / 339200., 187. / 2100., 1. / 40.]
DOPRI5_BERR = [35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 0.]

TABLEAUX = {
    "RKF45_REF": (RK45_AC, RK45_B, RK45_BERR, 6),
    "RK4": (RK4_AC, RK4_B, None, 4),
    "DOPRI5": (DOPRI5_AC, DOPRI5_B, DOPRI5_BERR, 7),
}


# --- right-hand sides (operation order identical to oracle/vecode_oracle.cpp `Rhs`) ----------------
def rhs_diag_linear(p):
    def f(t, x, dx):
        for c in range(len(x)):
            dx[c] = p[c] * x[c]
    return f


def rhs_harmonic2d(p):
    def f(t, x, dx):
        dx[0] = x[1]
        dx[1] = -(p[0] * x[0])
    return f


def rhs_lorenz63(p):
    def f(t, x, dx):
        dx[0] = p[0] * (x[1] - x[0])
        dx[1] = x[0] * (p[1] - x[2]) - x[1]
        dx[2] = x[0] * x[1] - p[2] * x[2]
    return f


def rhs_vdp(p):
    def f(t, x, dx):
        dx[0] = x[1]
        dx[1] = (p[0] * (1.0 - x[0] * x[0])) * x[1] - x[0]
    return f


def rhs_heat1d(p):
    def f(t, x, dx):
        n = len(x)
        for j in range(n):
            l = x[n - 1 if j == 0 else j - 1]
            r = x[0 if j + 1 == n else j + 1]
            dx[j] = p[0] * ((l + r) - 2.0 * x[j])
    return f


RHS = {0: rhs_diag_linear, 1: rhs_harmonic2d, 2: rhs_lorenz63, 3: rhs_vdp, 4: rhs_heat1d}


# --- LinearCombination: src/lc.rs:10-54 with src/impls/ndarray.rs:14-32 arithmetic ------------------
def lc_scale(v, k):
    for i in range(len(v)):
        v[i] = v[i] * k


def lc_scalar_multiply_to(v, k, target):
    for i in range(len(v)):
        target[i] = k * v[i]


def lc_add_scalar_mul(v, k, u):
    for i in range(len(v)):
        v[i] = v[i] + (k * u[i])


def lc_add_assign_ref(v, u):
    for i in range(len(v)):
        v[i] = v[i] + u[i]


def lc_delta(v, y):
    for i in range(len(v)):
        v[i] = v[i] - y[i]


def lc_linear_combination(v, v_arr, k_arr):
    if not v_arr or not k_arr:
        raise ValueError("linear_combination: slices cannot be empty")  # lc.rs:21-23
    lc_scalar_multiply_to(v_arr[0], k_arr[0], v)
    for vi, k in zip(v_arr[1:], k_arr[1:]):
        lc_add_scalar_mul(v, k, vi)


# --- the same with A = Complex<f64> (ndarray.rs:8-33 is generic over the element type): vectors of (re, im) tuples, k = (re, im).
# num-complex 0.4 (third-party; restated): Mul = (a.re b.re - a.im b.im, a.re b.im + a.im b.re), MulAssign the same two sums.
def _zmul(a, b):
    return (a[0] * b[0] - a[1] * b[1], a[0] * b[1] + a[1] * b[0])


def lcz_scale(v, k):
    for i in range(len(v)):
        re, im = v[i]
        v[i] = (re * k[0] - im * k[1], im * k[0] + re * k[1])  # MulAssign


def lcz_scalar_multiply_to(v, k, target):
    for i in range(len(v)):
        target[i] = _zmul(k, v[i])


def lcz_add_scalar_mul(v, k, u):
    for i in range(len(v)):
        p = _zmul(k, u[i])
        v[i] = (v[i][0] + p[0], v[i][1] + p[1])


def lcz_linear_combination(v, v_arr, k_arr):
    if not v_arr or not k_arr:
        raise ValueError("linear_combination: slices cannot be empty")  # lc.rs:21-23
    lcz_scalar_multiply_to(v_arr[0], k_arr[0], v)
    for vi, k in zip(v_arr[1:], k_arr[1:]):
        lcz_add_scalar_mul(v, k, vi)


# --- rk_step: src/base/rk.rs:90-155 ------------------------------------------------------------------
def rk_step(f, t, x0, xf, x_err, dt, ac, b, b_err, s, K):
    """Returns (xf, x_err) because the reference swaps the two buffers (rk.rs:142)."""
    f(t, x0, K[0])
    for i in range(1, s):
        row = ac[i * s:(i + 1) * s]
        ti = t + row[i] * dt
        lc_linear_combination(xf, K[:i], row[:i])
        lc_scale(xf, dt)
        lc_add_assign_ref(xf, x0)
        f(ti, xf, K[i])
    lc_linear_combination(xf, K[:s], b)
    lc_scale(xf, dt)
    lc_add_assign_ref(xf, x0)
    if b_err is not None and x_err is not None:
        x_err, xf = xf, x_err
        lc_linear_combination(xf, K[:s], b_err)
        lc_scale(xf, dt)
        lc_add_assign_ref(xf, x0)
        lc_delta(x_err, xf)
    return xf, x_err


# --- state machine: src/base/ode.rs -------------------------------------------------------------------
STEP, CHKPT, REJECT, END, ERR = range(5)
OK, DONE, SERR = range(3)


def relative_eq(a, b, eps=EPS, max_rel=EPS):
    """approx 0.5 RelativeEq for f64 (dependency; semantics restated)."""
    if a == b:
        return True
    if math.isinf(a) or math.isinf(b):
        return False
    ad = abs(a - b)
    if ad <= eps:
        return True
    largest = abs(b) if abs(b) > abs(a) else abs(a)
    return ad <= largest * max_rel


def check_step(t0, tf, dt):  # ode.rs:389-399
    rem = tf - t0
    if relative_eq(rem, 0.0):
        return None
    return rem if rem < dt else dt


def _rmax(a, b):  # Rust f64::max: NaN-dropping
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _rmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def norm_of(v, kind):
    if callable(kind):  # the user's `Normed` impl (ode.rs:9-11): any function of the error vector
        return kind(v)
    if kind == 0:  # L2
        acc = 0.0
        for e in v:
            acc = acc + e * e
        return math.sqrt(acc)
    if kind == 1:  # Linf
        acc = 0.0
        for e in v:
            acc = _rmax(acc, abs(e))
        return acc
    if kind == 2:  # L1
        acc = 0.0
        for e in v:
            acc = acc + abs(e)
        return acc
    if kind == 3:  # hypot (complex scalar, rk.rs:209-214)
        if len(v) == 2:
            return _libm.hypot(v[0], v[1])
        acc = 0.0
        for i in range(0, len(v) - 1, 2):
            m = _libm.hypot(v[i], v[i + 1])
            acc = acc + m * m
        return math.sqrt(acc)
    raise ValueError(kind)


class RKSolver:
    """RK45Solver (src/base/rk.rs:158-320) with a runtime tableau."""

    def __init__(self, f, tableau, t0, tf, x0, h):
        ac, b, b_err, s = tableau
        self.f, self.ac, self.b, self.b_err, self.s = f, ac, b, b_err, s
        # ODEData::new, ode.rs:141-150
        self.t0, self.tf, self.t = t0, tf, t0
        self.x = list(x0)
        self.next_x = list(x0)
        self.t_list = [t0, tf]
        self.tgt_t = 0
        self.next_dt = h
        self.h = h
        self.prev_h = h
        # ODEAdaptiveData::new_with_defaults(order 3).with_alpha(0.9), rk.rs:258-260, ode.rs:114-128
        self.atol, self.rtol, self.dx_norm = 1.0e-6, 1.0e-4, 0.0
        self.alpha, self.min_dt, self.max_dt, self.pow = 0.9, 1.0e-6, 1.0, 1.0 / 3.0
        self.x_err = list(x0)  # Some(x0.clone()), rk.rs:249
        self.K = [list(x0) for _ in range(s + 1)]  # rk.rs:255-256
        self.norm_kind = 0
        self.n_accept = self.n_reject = self.n_calls = 0

    # builders ---------------------------------------------------------------------------------------
    def no_adaptive(self):  # rk.rs:233-237
        self.x_err = None
        return self

    def with_tolerance(self, atol, rtol):  # ode.rs:298-306
        if atol <= 0.0 or rtol <= 0.0:
            raise ValueError(f"Invalid tolerances: atol={atol}, rtol={rtol}")
        self.atol, self.rtol = atol, rtol
        return self

    def with_step_range(self, dt_min, dt_max):  # ode.rs:267-285
        if dt_min <= 0.0 or dt_max <= 0.0 or dt_max <= dt_min:
            raise ValueError(f"Invalid step range: ({dt_min}, {dt_max})")
        self.min_dt, self.max_dt = dt_min, dt_max
        self.h = self.prev_h = math.sqrt(dt_min * dt_max)
        return self

    def with_init_step(self, h):  # ode.rs:287-296
        if h < self.min_dt or h > self.max_dt:
            raise ValueError(f"Step {h} is not inside the range ({self.min_dt}, {self.max_dt})")
        self.h = self.prev_h = h
        return self

    # stepping ---------------------------------------------------------------------------------------
    def step_size(self):  # ode.rs:165-181
        if self.tgt_t >= len(self.t_list):
            return END, 0.0
        dt = check_step(self.t, self.t_list[self.tgt_t], self.h)
        if dt is not None:
            return STEP, dt
        if self.tgt_t >= len(self.t_list) - 1:
            return END, 0.0
        return CHKPT, 0.0

    def try_step(self, dt):  # rk.rs:287-293
        self.next_x, self.x_err = rk_step(self.f, self.t, self.x, self.next_x, self.x_err, dt, self.ac, self.b,
                                          self.b_err, self.s, self.K)

    def _apply(self, ev, adaptive):  # ode.rs:402-428
        if ev == STEP:
            self.x, self.next_x = self.next_x, self.x  # advance, ode.rs:184-188
            self.t += self.next_dt
            self.n_accept += 1
            return OK
        if ev == CHKPT:
            self.tgt_t += 1
            self.h = self.prev_h
            return OK
        if ev == REJECT:
            self.n_reject += 1
            return OK if adaptive else SERR
        if ev == END:
            self.tgt_t += 1
            self.h = self.prev_h
            return DONE
        return SERR

    def step(self):  # ode.rs:249-253
        self.n_calls += 1
        ev, dt = self.step_size()
        if ev == STEP:
            self.next_dt = dt
            self.try_step(dt)
        self.last_event = ev
        return self._apply(ev, False)

    def step_adaptive(self):  # ode.rs:311-341
        self.n_calls += 1
        if self.x_err is None:
            raise RuntimeError("adaptive step validation failed")  # ode.rs:312
        ev, dt = self.step_size()
        h = self.h
        if ev == STEP:
            self.next_dt = dt
            self.try_step(dt)
            self.dx_norm = norm_of(self.x_err, self.norm_kind)
            f = self.rtol / self.dx_norm if self.dx_norm != 0.0 else math.copysign(math.inf, self.rtol)
            mul = self.alpha * _libm.pow(f, self.pow)
            fp_lim = _rmin(_rmax(mul, 0.3), 2.0)
            new_h = _rmin(_rmax(fp_lim * h, self.min_dt), self.max_dt)
            self.prev_h, self.h = self.h, new_h
            if f <= 1.0:
                ev = REJECT
        self.last_event = ev
        return self._apply(ev, True)

    def run(self, adaptive=False, max_calls=0):
        st = OK
        while st == OK and (max_calls <= 0 or self.n_calls < max_calls):
            st = self.step_adaptive() if adaptive else self.step()
        return st


# --- synthetic inputs: counter-based splitmix64 ---------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(seed: int, index: int) -> int:
    z = (seed + (index + 1) * 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def uniform01(seed: int, index: int) -> float:
    return (splitmix64(seed, index) >> 11) * (1.0 / 9007199254740992.0)
