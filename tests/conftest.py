import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "guard_selftest: damages a guard zone on purpose (exempt from the per-test guard assertion)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test infrastructure): C++ via ctypes."""
    from oracle import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def vo():
    import vecode_b200
    return vecode_b200


@pytest.fixture()
def ctx(vo):
    c = vo.Context(0, arith="strict")
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _guard_zones(request):
    """VECODE_GUARD=1 (the library's own bounds check, include/vecode_b200.h: vo_guard_check): after every GPU test no device block
    of the library may have been written outside its bounds. A no-op without the switch and for CPU tests."""
    yield
    if os.environ.get("VECODE_GUARD", "0") in ("", "0") or request.node.get_closest_marker("gpu") is None:
        return
    if request.node.get_closest_marker("guard_selftest") is not None:
        return
    from vecode_b200 import _cabi
    lib = _cabi.lib()
    assert lib.vo_guard_enabled() == 1
    before = getattr(_guard_zones, "seen", 0)
    now = int(lib.vo_guard_check(None))
    _guard_zones.seen = now
    assert now == before, f"{now - before} device block(s) written out of bounds during {request.node.nodeid} (see stderr)"
