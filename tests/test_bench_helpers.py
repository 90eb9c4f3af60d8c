"""bench.py helpers that only run on the GPU box otherwise: launch accounting and the bounded CPU sample."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_library_launches_only_on_ab_switches(monkeypatch):
    # default routes: both contractions are this repo's tcgen05 kernels -> no library GEMM whatever the profile says
    prof = {"counts": (3.0, 5), "solve": (5.6, 10), "colsum": (1.7, 10), "cross": (10.3, 5), "gram_i8": (10.9, 5), "finalize": (0.9, 5)}
    monkeypatch.delenv("PLSPM_GRAM", raising=False)
    monkeypatch.delenv("PLSPM_VOTE", raising=False)
    assert bench.library_launches(prof) == 0
    # A/B switches bring cuBLAS back and the count says so
    monkeypatch.setenv("PLSPM_VOTE", "cublas")
    assert bench.library_launches({"cross": (10.0, 125)}) == 125
    monkeypatch.setenv("PLSPM_GRAM", "cublas")
    assert bench.library_launches({"gram_i8": (10.0, 10), "colsum": (1.0, 15)}) > 0


def test_kernel_names_cover_every_stage():
    for st in ("counts", "gram", "reduce", "solve", "scores", "upload", "colsum", "cross", "scoregen", "conv", "gram_i8", "finalize"):
        assert st in bench.KERNEL_NAMES


def test_cpu_workload_subsamples_large_inputs():
    small = dict(X=np.zeros((1000, 20)), blocks=[20])
    w, scale, note = bench.cpu_workload(small)
    assert w is small and scale == 1.0 and note == ""
    big = dict(X=np.zeros((200_000, 256)), blocks=[256])
    w, scale, note = bench.cpu_workload(big)
    n_sub = int(bench.CPU_MAX_ELEMENTS // 256)
    assert w["X"].shape == (n_sub, 256) and abs(scale - n_sub / 200_000) < 1e-15 and "EXTRAPOLATED" in note


def test_workload_table_is_consistent():
    for name, (N, L, K, mode, scheme, reps, desc) in bench.WORKLOADS.items():
        assert scheme in ("centroid", "factorial", "path") and mode in (0, 1) and reps > 0 and desc
