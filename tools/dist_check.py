"""torchrun smoke for the NCCL path of plspm.bootstrap.Bootstrap: every rank must end up with all
replicates, in global order, equal to the oracle on the same Philox indices.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "plspm-python_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import plspm.config as c  # noqa: E402
from oracle import plspm_oracle as orc  # noqa: E402
from plspm.mode import Mode  # noqa: E402
from plspm.plspm import Plspm  # noqa: E402
from plspm_b200 import engine  # noqa: E402
from plspm_b200.synth import make_synthetic  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
engine.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
N, L, K, B = 3000, 5, 4, 37
X, path = make_synthetic(N, L, K, seed=4)
lvs = ["lv%d" % j for j in range(L)]
cols = ["x%d_%d" % (j, k) for j in range(L) for k in range(K)]
df = pd.DataFrame(X, columns=cols)
cfg = c.Config(pd.DataFrame(path.astype(int), index=lvs, columns=lvs), scaled=True)
for j, lv in enumerate(lvs):
    cfg.add_lv(lv, Mode.A, *[c.MV(m) for m in cols[j * K:(j + 1) * K]])
calc = Plspm(df, cfg, bootstrap=True, bootstrap_iterations=B, processes=1, bootstrap_seed=5)
w = calc.bootstrap().samples()["weights"].loc[:, cols].values
idx = np.stack([orc.philox_indices(5, b, N) for b in range(B)])
rows, iters, status = orc.bootstrap(X, idx, [K] * L, [0] * L, path, "centroid", True)
np.testing.assert_allclose(w, rows[:, :L * K], rtol=1e-6)
st, it = calc.bootstrap().replicate_status()
np.testing.assert_array_equal(it, iters)
print("rank %d/%d: %d replicates gathered over NCCL match the oracle" % (rank, world, B), flush=True)
dist.barrier()
dist.destroy_process_group()
