"""bench.py helpers that only run on the GPU box otherwise: launch accounting and the bounded CPU sample."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_split_launches_separates_library_gemms():
    # c3, 5 steps: counts 5, solve 10, colsum 3 per step (counts8 + GEMM + combine), 25 vote chunks per step
    prof = {"counts": (3.0, 5), "solve": (5.6, 10), "colsum": (1.7, 15), "cross": (10.3, 125), "scoregen": (17.5, 125),
            "gram_i8": (10.9, 10), "gram": (0.0, 0)}
    assert bench.split_launches(prof, 5, 100_000, 5056) == (155, 135)
    # small data: fp64 kernels only, no library launches
    prof = {"counts": (0.05, 5), "gram": (0.3, 5), "reduce": (0.05, 5), "solve": (0.3, 5), "colsum": (0.3, 5)}
    assert bench.split_launches(prof, 5, 250, 1000) == (25, 0)
    # c5, streamed planes: generator + GEMM + combine per chunk
    prof = {"counts": (6, 2), "solve": (40, 4), "colsum": (17, 18), "cross": (200, 490), "scoregen": (300, 490), "gram_i8": (580, 126)}
    assert bench.split_launches(prof, 2, 1_000_000, 40704) == (590, 540)


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
