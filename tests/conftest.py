"""pytest configuration: the ``gpu`` marker (tests that need a B200) and import path setup."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from ml_inference_optimizer_b200 import build as b
    from ml_inference_optimizer_b200 import _lib

    b.build()
    return _lib.load()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
