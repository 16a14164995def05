import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
