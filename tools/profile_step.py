"""A few bootstrap batches of a bench workload on resident data (the command ncu wraps; see profiles/README.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "plspm-python_b200"))
import numpy as np
from plspm_b200 import engine
from bench import load_workload
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = load_workload(name)
reps = reps or w["reps"]
engine.set_device(0)
model = engine.Model(w["blocks"], w["modes"], w["path"], w["scaled"], numeric=bool(w.get("numeric")))
data = engine.Data(model, w["X"])
for s in range(steps):
    rows, status, iters = engine.bootstrap(model, data, w["scheme"], s * reps, reps, seed=0)
print("ok", int((status == 0).sum()), float(iters.mean()), {k: v for k, v in engine.profile_get().items() if v[1]})
