"""The library's own bounds check (VECODE_GUARD=1; include/vecode_b200.h: vo_guard_enabled / vo_guard_check).

compute-sanitizer is closed on the GPU pool, so memory safety of the kernels is checked by the library itself: with the switch on,
every device block it allocates lies between two 4 KiB guard zones, and tests/conftest.py asserts after EVERY GPU test that no zone
was written. This file checks the checker: the switch is read from the environment at load time, so the self-test runs the library in
a child process with the switch on, damages a block on purpose (a wrapped view one element longer than the block it sits on) and
expects exactly that to be reported; a clean run of every ragged ensemble size must report nothing.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, {root!r})
import vecode_b200 as vo
from vecode_b200 import _cabi
lib = _cabi.lib()
assert lib.vo_guard_enabled() == 1
ctx = vo.Context(0, arith="strict")
live = C.c_int64(0)

# 1. clean: ragged ensemble sizes through the LinearCombination kernels and both RK paths
for n in (1, 3, 127, 128, 129, 255, 257, 1001, 4097):
    a = vo.Ensemble.from_host(ctx, np.random.default_rng(n).standard_normal((n, 3)))
    b = a.clone()
    vo.LinearCombination.scale(a, 0.5)
    vo.LinearCombination.add_scalar_mul(a, 2.0, b)
    f = vo.Rhs(ctx, "LORENZ63", 3, [10.0, 28.0, 8.0 / 3.0])
    s = vo.RK45Solver(f, 0.0, 0.05, a, 0.01, vo.ButcherTableu.builtin("RK4")).no_adaptive()
    s.run()
    g = vo.Rhs(ctx, "VDP", 2, [1.5])
    x = vo.Ensemble.from_host(ctx, np.tile([2.0, 0.0], (n, 1)))
    s2 = vo.RK45Solver(g, 0.0, 0.5, x, 0.01, vo.ButcherTableu.builtin("DOPRI5")).with_tolerance(1e-9, 1e-6)
    s2.run(adaptive=True)
    del s, s2
ctx.sync()
clean = int(lib.vo_guard_check(C.byref(live)))
print("clean", clean, "live", live.value)

# 2. damaged on purpose: a view one element longer than the block it sits on, scaled in place
n = 1000
e = vo.Ensemble.from_host(ctx, np.ones((n, 1)))
h = C.c_void_p()
_cabi.check(lib.vo_ens_wrap(ctx._h, C.c_void_p(e.device_ptr), 1, n + 1, C.byref(h)), ctx._h)
_cabi.check(lib.vo_lc_scale(h, C.c_double(2.0)), ctx._h)
ctx.sync()
lib.vo_ens_destroy(h)
dirty = int(lib.vo_guard_check(None))
again = int(lib.vo_guard_check(None))
print("dirty", dirty, "again", again)
"""


@pytest.mark.gpu
@pytest.mark.guard_selftest
def test_guard_zones_report_an_overrun_and_nothing_else():
    env = dict(os.environ, VECODE_GUARD="1")
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict()
    for line in r.stdout.splitlines():
        w = line.split()
        if w and w[0] in ("clean", "dirty"):
            out[w[0]] = (int(w[1]), int(w[3]))
    assert out["clean"][0] == 0 and out["clean"][1] > 0, r.stdout + r.stderr      # blocks were live and none was damaged
    assert out["dirty"] == (1, 1), r.stdout + r.stderr                             # exactly the block written past its end, counted once
    # the element behind the block held the pattern 0xA5A5...; doubling it as a double changes its exponent byte only
    assert "written outside a 8000-byte device block; first at end+" in r.stderr, r.stderr


@pytest.mark.gpu
def test_guard_switch_is_off_by_default(vo):
    from vecode_b200 import _cabi
    if os.environ.get("VECODE_GUARD", "0") not in ("", "0"):
        pytest.skip("the suite itself runs under VECODE_GUARD=1")
    lib = _cabi.lib()
    assert lib.vo_guard_enabled() == 0 and lib.vo_guard_check(None) == 0
