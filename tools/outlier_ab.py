"""A/B of the second-moment routes on data with one gross outlier (run per route: PLSPM_GRAM=<route>)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "plspm-python_b200"))
from oracle import plspm_oracle as orc
from plspm_b200 import engine
from plspm_b200.synth import make_synthetic
N, L, K = 8192, 5, 4
for mag in (1e3, 1e4, 1e5, 1e6):
    X, path = make_synthetic(N, L, K, 31)
    X[1234, 6] = mag
    X[:, 13] *= 1.0e-5
    model = engine.Model([K] * L, [0] * L, path, True)
    data = engine.Data(model, X)
    rows, status, iters = engine.bootstrap(model, data, "centroid", 0, 4, seed=11)
    worst = 0.0
    for b in range(4):
        ref, it, st = orc.replicate_row(X, orc.philox_indices(11, b, N), [K] * L, [0] * L, path, "centroid", True)
        worst = max(worst, float(np.max(np.abs(rows[b] - ref) / np.maximum(np.abs(ref), 1e-3))))
    print(os.environ.get("PLSPM_GRAM", "mma"), "outlier %.0e" % mag, "max rel diff vs oracle %.3e" % worst, "iters", iters.tolist(), flush=True)
