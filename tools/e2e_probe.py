"""Where the end-to-end time of plspm_bootstrap_host goes: upload vs bootstrap vs teardown (c3 shape)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plspm-python_b200"))
from plspm_b200 import engine  # noqa: E402
from plspm_b200.synth import make_synthetic  # noqa: E402

engine.set_device(0)
N, L, K, B = 100000, 32, 8, 1536
X, path = make_synthetic(N, L, K, seed=0)
model = engine.Model([K] * L, [0] * L, path, True)
Xp = engine.pinned_empty(X.shape)
Xp[...] = X
out = engine.pinned_empty((B, model.n_out))
for r in range(6):
    t0 = time.perf_counter()
    data = engine.Data(model, Xp)
    t1 = time.perf_counter()
    engine.profile_reset()
    rows, status, iters = engine.bootstrap(model, data, "centroid", r * B, B, seed=0)
    t2 = time.perf_counter()
    prof = engine.profile_get()
    data.close()
    t3 = time.perf_counter()
    engine.bootstrap_host(model, Xp, "centroid", r * B, B, seed=0, out=out)
    t4 = time.perf_counter()
    print("create %.2f ms | bootstrap %.2f ms (device stages %.2f) | destroy %.2f ms | bootstrap_host %.2f ms -> %.0f fits/s" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), sum(v[0] for v in prof.values()), 1e3 * (t3 - t2), 1e3 * (t4 - t3), B / (t4 - t3)),
        flush=True)
