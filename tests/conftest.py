import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "plspm-python_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sat():
    return np.load(os.path.join(GOLDEN, "satisfaction.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def syn():
    return np.load(os.path.join(GOLDEN, "synthetic.npz"), allow_pickle=False)
