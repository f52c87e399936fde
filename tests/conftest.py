import os
import sys

import numpy as np
import pytest

# importing pnumpy_b200 must not try to grab a GPU in the CPU-only runs; GPU tests call pn.init()
os.environ.setdefault("PNUMPY_B200_AUTOINIT", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="function")
def rng():
    # the reference's fixture seed (tests/conftest.py:21-30)
    return np.random.default_rng(1234)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
