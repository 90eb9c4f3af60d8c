"""Quick GPU probe: bootstrap throughput + per-stage device time for one synthetic config."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plspm-python_b200"))
from plspm_b200 import engine  # noqa: E402
from plspm_b200.synth import make_synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=100000)
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--B", type=int, default=1184)
ap.add_argument("--scheme", default="centroid")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--policy", type=int, default=0)
ap.add_argument("--numeric", action="store_true", help="non-metric estimator for numeric scales")
a = ap.parse_args()

engine.set_device(0)
t0 = time.time()
X, path = make_synthetic(a.N, a.L, a.K, seed=0)
print("gen %.1fs" % (time.time() - t0), flush=True)
model = engine.Model([a.K] * a.L, [a.mode] * a.L, path, True, a.policy, numeric=a.numeric)
t0 = time.time()
data = engine.Data(model, X)
print("upload %.3fs  tiles=%d tile_groups=%d full=%s" % (time.time() - t0, model.n_tiles, model.n_tile_groups, model.full_tiles), flush=True)
t0 = time.time()
fit_model = engine.Model([a.K] * a.L, [a.mode] * a.L, path, True, engine.TILES_FULL, numeric=True) if a.numeric else model
f = engine.fit(fit_model, data, a.scheme)
print("fit %.4fs iters=%d status=%d" % (time.time() - t0, f["iterations"], f["status"]), flush=True)
for r in range(a.reps):
    engine.profile_reset()
    t0 = time.time()
    rows, status, iters = engine.bootstrap(model, data, a.scheme, 0, a.B, seed=r)
    dt = time.time() - t0
    prof = engine.profile_get()
    nit = float(iters.mean())
    gb = (iters.astype(np.float64) + 2).sum() * a.N * a.L * a.K * 8 / 1e9
    print("B=%d %.4fs %.1f fits/s ok=%d mean_iters=%.2f alg %.1f GB/s | %s" % (
        a.B, dt, a.B / dt, int((status == 0).sum()), nit, gb / dt,
        " ".join("%s=%.2fms/%d" % (k, v[0], v[1]) for k, v in prof.items() if v[1])), flush=True)
