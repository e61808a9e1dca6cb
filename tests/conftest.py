import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (oracle/): the checker, never the thing under test."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def cb():
    import cubature_b200
    cubature_b200._ffi.lib()  # raises if libcubature_cuda.so is not built: no fallback
    return cubature_b200
