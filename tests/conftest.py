import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the C-ABI library; GPU tests fail loudly if that is impossible."""
    from a3t_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()
