import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure): built on demand from oracle/*.cpp."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.load()


@pytest.fixture(scope="session")
def gpu_lib():
    """libflipb200.so, the product. No fallback: a missing library or device fails the test."""
    from zeno_b200 import abi
    lib = abi.load_library()
    assert lib.flipb200_device_count() > 0, "no CUDA device visible"
    return lib
