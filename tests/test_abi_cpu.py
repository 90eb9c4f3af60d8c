"""The C-ABI library builds, loads and exports every symbol include/plspm_b200.h declares; and the
product path fails LOUDLY without a GPU (no CPU fallback).  CPU only -- no compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plspm_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from plspm_b200 import engine
    assert os.path.exists(engine.LIB_PATH)
    return engine.LIB_PATH


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(plspm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for required in ("plspm_model_create", "plspm_data_create", "plspm_fit", "plspm_bootstrap",
                     "plspm_bootstrap_host", "plspm_resample_indices", "plspm_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    from plspm_b200 import engine
    assert sorted(engine.EXPORTS) == declared_functions()


def test_library_contains_sm100a_code_with_tma_and_fp64_fma(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass          # cp.async.bulk (TMA engine) in the Gram kernel
    assert "SYNCS" in sass           # mbarrier
    assert sass.count("DFMA") > 64   # fp64 register-tile accumulation


def test_no_gpu_means_loud_failure(lib_path):
    from plspm_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    path = np.array([[0, 0], [1, 0]], dtype=np.int8)
    with pytest.raises(engine.EngineError):
        engine.Model([2, 2], [0, 0], path, True)


def test_invalid_models_are_rejected_before_touching_the_device(lib_path):
    from plspm_b200 import engine
    with pytest.raises(engine.EngineError):  # upper-triangular path
        engine.Model([2, 2], [0, 0], np.array([[0, 1], [0, 0]], dtype=np.int8), True)
    with pytest.raises(engine.EngineError):  # empty block
        engine.Model([2, 0], [0, 0], np.array([[0, 0], [1, 0]], dtype=np.int8), True)
    with pytest.raises(ValueError):
        engine.Model([2, 2], [0], np.array([[0, 0], [1, 0]], dtype=np.int8), True)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "plspm-python_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") or f.endswith((".cu", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "plspm_emul" not in text, f
