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


def _gpu_available() -> bool:
    try:
        from plspm_b200 import engine
        return engine.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device (or without the built library) skips the gpu-marked tests
    instead of failing them; with `-m gpu` on the GPU box nothing is skipped, and a missing library is an error
    there (tests/test_gpu_* import the engine and fail loudly)."""
    if os.environ.get("PLSPM_REQUIRE_GPU") or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device / libplspm_b200.so: gpu tests need the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def sat():
    return np.load(os.path.join(GOLDEN, "satisfaction.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def syn():
    return np.load(os.path.join(GOLDEN, "synthetic.npz"), allow_pickle=False)
