"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): one fit and a handful of
bootstrap replicates for a full and a sparse tile set, with and without the fast sign vote."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plspm-python_b200"))
from plspm_b200 import engine  # noqa: E402
from plspm_b200.synth import make_synthetic  # noqa: E402

engine.set_device(0)
for (N, L, K, policy) in ((700, 5, 3, 1), (900, 7, 8, 2), (4200, 6, 16, 2)):
    X, path = make_synthetic(N, L, K, seed=1, reverse_blocks=(1,))
    model = engine.Model([K] * L, [0, 1] * (L // 2) + [0] * (L % 2), path, True, policy)
    data = engine.Data(model, X)
    f = engine.fit(model, data, "path")
    rows, status, iters = engine.bootstrap(model, data, "path", 0, 5, seed=2)
    print(N, L, K, policy, "fit iters", f["iterations"], "boot status", status.tolist(), "finite", bool(np.isfinite(rows).all()))
    data.close()
    model.close()
print("redo", engine.redo_count())
