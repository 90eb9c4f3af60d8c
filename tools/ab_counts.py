import os, sys, subprocess, json
ROOT="/root/repo"
code = r'''
import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/plspm-python_b200")
import numpy as np
from plspm_b200 import engine
from bench import load_workload
w = load_workload(sys.argv[1]); reps = int(sys.argv[2])
engine.set_device(0)
model = engine.Model(w["blocks"], w["modes"], w["path"], w["scaled"], numeric=bool(w.get("numeric")))
data = engine.Data(model, w["X"])
engine.profile_reset()
rows, status, iters = engine.bootstrap(model, data, w["scheme"], 7, reps, seed=3)
rows2, status2, iters2 = engine.bootstrap(model, data, w["scheme"], 7 + reps, reps, seed=3)
np.save(sys.argv[3], rows)
print(json.dumps({k: v for k, v in engine.profile_get().items() if v[1]}), int((status==0).sum()), float(iters.mean()))
'''
open("/root/repo/gpurun_out/t_fused_child.py","w").write(code)
outs = {}
for name, env in (("fused", {}), ("split", {"PLSPM_COUNTS": "split"})):
    for wl, reps in (("c3", 1284), ("c3", 1536), ("c2", 300)):
        e = dict(os.environ); e.update(env)
        f = "/root/repo/gpurun_out/rows_%s_%s_%d.npy" % (name, wl, reps)
        r = subprocess.run([sys.executable, "/root/repo/gpurun_out/t_fused_child.py", wl, str(reps), f], env=e, capture_output=True, text=True)
        print(name, wl, reps, r.stdout.strip()[-600:], r.stderr.strip()[-300:])
import numpy as np
for wl, reps in (("c3", 1284), ("c3", 1536), ("c2", 300)):
    a = np.load("/root/repo/gpurun_out/rows_fused_%s_%d.npy" % (wl, reps)); b = np.load("/root/repo/gpurun_out/rows_split_%s_%d.npy" % (wl, reps))
    print(wl, reps, "identical" if np.array_equal(a, b) else "max diff %g" % np.abs(a - b).max())
