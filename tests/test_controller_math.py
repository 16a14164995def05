"""The two reformulations of the step-size controller's arithmetic (handle_step_adaptive, src/base/ode.rs:311-334), emulated
operation for operation on the host: the GPU kernels (vec-ode_b200/csrc/rk_small.cuh: pow_third_cr, controller_l2_fast) use
exactly these operations, with the SFU seed replaced here by a float32 seed perturbed by its worst-case error."""
import math
from decimal import Decimal, getcontext

import numpy as np

getcontext().prec = 60
P = 1.0 / 3.0  # `order.recip()` with order = 3 (rk.rs:258-260, ode.rs:120): the double nearest to 1/3
DELTA = float(Decimal(P) - Decimal(1) / Decimal(3))


def _fma(a, b, c):
    return float(Decimal(float(a)) * Decimal(float(b)) + Decimal(float(c)))


def pow_third_cr(f, seed_err):
    """STRICT: powf(f, 1/3) correctly rounded — seed, third-order step, double-double Newton step, exponent offset."""
    z = float(np.float32(np.exp2(np.float32(np.float32(np.log2(np.float32(f))) * np.float32(-1.0 / 3.0))))) * (1.0 + seed_err)
    r = _fma(-f, (z * z) * z, 1.0)
    z = _fma(z * r, _fma(r, 2.0 / 9.0, 1.0 / 3.0), z)
    zz = z * z
    y = f * zz
    y2 = y * y
    y2l = _fma(y, y, -y2)
    y3 = y2 * y
    y3l = _fma(y2, y, -y3) + y2l * y
    res = (f - y3) - y3l
    c = res * (zz * (1.0 / 3.0))
    c = _fma(y * -1.850371707708594e-17, float(np.float32(np.log2(np.float32(f)))) * 0.6931471805599453, c)
    return y + c


def test_exponent_offset_constant():
    assert DELTA == -1.850371707708594e-17


def test_strict_pow_third_is_correctly_rounded_and_matches_libm():
    rng = np.random.default_rng(1)
    n, wrong, libm_differs = 4000, 0, 0
    for _ in range(n):
        f = float(10.0 ** rng.uniform(-6, 6))
        exact = float((Decimal(f).ln() * Decimal(P)).exp())  # Decimal -> float rounds to nearest
        got = pow_third_cr(f, rng.uniform(-1, 1) * 2.0 ** -20)
        wrong += got != exact
        libm_differs += math.pow(f, P) != exact
    assert wrong == 0
    assert libm_differs <= 0.01 * n  # the reference's powf (libm) is itself correctly rounded except for rare arguments


def test_fast_controller_factor_within_an_ulp_or_two():
    """FAST: alpha * (rtol/dx_norm)^(1/3) as alpha * g^(-1/6), g = (dx_norm/rtol)^2: SFU seed and ONE third-order step."""
    rng = np.random.default_rng(2)
    worst = 0.0
    for _ in range(4000):
        g = float(10.0 ** rng.uniform(-2.2, 2.9))  # the range where the factor is not clamped to [0.3, 2]
        y = float(np.float32(np.exp2(np.float32(np.float32(np.log2(np.float32(g))) * np.float32(-0.16666667))))) * (1.0 + rng.uniform(-1, 1) * 2.0 ** -20)
        y2 = y * y
        r = max(_fma(-g, (y2 * y2) * y2, 1.0), -1.0)
        y = _fma(y * r, _fma(r, 7.0 / 72.0, 1.0 / 6.0), y)
        exact = (Decimal(g).ln() * (Decimal(-1) / Decimal(6))).exp()
        worst = max(worst, abs(float((Decimal(y) - exact) / exact)) / 2.0 ** -53)
    assert worst <= 2.0, worst
