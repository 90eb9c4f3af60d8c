#!/usr/bin/env python3
"""Times the REFERENCE ITSELF (plspm-python, staged under baseline/_ref by stage_reference.py) on the host
cores, through its own public API, in a process of its own (its package is also called `plspm`).

    python baseline/ref_runner.py --workload c3 --rows 30000 --replicates 16 --processes 16 --steps 3 --warmup 1

Setup (untimed, reported as `single_fit_s`): `Plspm(data, config, scheme)` -- the reference's single fit, which is
BASELINE config 3's figure.  Each timed step is the reference's own bootstrap, `plspm.bootstrap.Bootstrap(config,
data, inner_model, outer_model, calculator, replicates, processes)` (bootstrap.py:83-119: fork `processes`
workers x replicates // processes fits each, Queue gather with its 1-s poll).  `--rows` restricts the data to its
first rows (the caller scales the throughput by rows / N and labels it EXTRAPOLATED: every O(N) step of the
reference -- DataFrame.dot, .corr() -- is linear in the rows).  Prints one JSON object.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(HERE, "_ref")
sys.path.insert(0, REFDIR)                                   # the reference `plspm` + the statsmodels stand-in
sys.path.append(os.path.join(ROOT, "plspm-python_b200"))     # plspm_b200.synth only (after the reference!)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402


def build(workload: str, rows: int):
    import plspm.config as c
    from plspm.mode import Mode
    from plspm.scale import Scale
    from plspm.scheme import Scheme
    assert os.path.realpath(c.__file__).startswith(os.path.realpath(REFDIR)), "not the staged reference: " + c.__file__
    sys.path.insert(0, ROOT)
    from bench import WORKLOADS  # the table of workloads (N, L, K, mode, scheme, ...)
    N, L, K, mode, scheme, _, _ = WORKLOADS[workload]
    if workload == "c2":
        g = np.load(os.path.join(ROOT, "tests", "golden", "satisfaction.npz"))
        X, pathm, blocks = g["X"], g["path"], [int(v) for v in g["block_sizes"]]
        lvs, mvs, scaled = [str(v) for v in g["lvs"]], [str(v) for v in g["mvs"]], False
    else:
        from plspm_b200.synth import make_synthetic
        X, pathm = make_synthetic(N, L, K, seed=0)
        blocks, scaled = [K] * L, True
        lvs, mvs = ["lv%d" % i for i in range(L)], ["x%d" % i for i in range(L * K)]
    if rows and rows < X.shape[0]:
        X = X[:rows]
    data = pd.DataFrame(X, columns=mvs)
    path = pd.DataFrame(np.asarray(pathm, dtype=np.int64), index=lvs, columns=lvs)
    numeric = workload.endswith("n")
    config = c.Config(path, default_scale=Scale.NUM) if numeric else c.Config(path, scaled=scaled)
    o = 0
    for lv, k in zip(lvs, blocks):
        config.add_lv(lv, Mode.B if mode else Mode.A, *[c.MV(m) for m in mvs[o:o + k]])
        o += k
    sch = {"centroid": Scheme.CENTROID, "factorial": Scheme.FACTORIAL, "path": Scheme.PATH}[scheme]
    return data, config, sch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--replicates", type=int, default=0)
    ap.add_argument("--processes", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    args = ap.parse_args()
    try:  # one BLAS thread per worker process: the workers are the parallelism (bootstrap.py:93-96)
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from plspm.bootstrap import Bootstrap
    from plspm.plspm import Plspm
    data, config, scheme = build(args.workload, args.rows)
    t = time.perf_counter()
    calc = Plspm(data, config, scheme, 100, 1e-6)
    single = time.perf_counter() - t
    out = {"workload": args.workload, "rows": int(data.shape[0]), "single_fit_s": single, "step_s": [],
           "replicates": args.replicates, "processes": args.processes}
    if args.replicates:
        # the objects Plspm.__init__ hands to Bootstrap (plspm.py:76-81)
        im, om = calc._Plspm__inner_model, calc._Plspm__outer_model
        from plspm.weights import WeightsCalculatorFactory
        filtered = config.filter(data)
        n = filtered.shape[0]
        calculator = WeightsCalculatorFactory(config, 100, 1e-6, float(np.sqrt(n / (n - 1))), scheme)
        for s in range(args.warmup + args.steps):
            t = time.perf_counter()
            b = Bootstrap(config, filtered, im, om, calculator, args.replicates, args.processes)
            dt = time.perf_counter() - t
            assert np.isfinite(b.weights()["mean"].to_numpy(dtype=float)).all(), "reference bootstrap produced no rows"
            if s >= args.warmup:
                out["step_s"].append(dt)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
