#!/usr/bin/env python3
"""bench.py -- bootstrap PLS-PM fits/s of the B200 engine (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5|c3f|c3n] [--impl reference]

A *step* is one bootstrap batch on every rank: `replicates_per_gpu_per_step` independent PLS-PM fits
to convergence (tol 1e-6, max 100 iterations) on resamples of the HBM-resident observation matrix,
drawn from the Philox stream of their GLOBAL replicate ids; for N > 1 the run ends with ONE all-gather
of the per-replicate rows of all K steps (inside the timed region).  Weak scaling: per-GPU work is fixed.

  value     fits/s over all ranks, inputs resident in HBM, result rows left on the device
  e2e       the same step through plspm_bootstrap_host(): X starts in pinned HOST memory, is
            uploaded inside the timed region, and the rows come back to host memory
  roofline  the dominant kernel against its bound: tensor work / CUDA-event time vs the measured dense
            peak; plus roofline.hbm: measured DRAM traffic fraction (frac_hw) and SURVEY 8(d)'s figure
  parity_checked
            four replicate rows of the timed run compared with the CPU oracle at 1e-6
  cpu_baseline / --impl reference
            the reference itself (baseline/_ref, staged by baseline/stage_reference.py) through its own
            bootstrap API in a subprocess on the box's host cores, on a bounded sample; the oracle port
            only as a labelled fallback when the staged reference is missing
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "plspm-python_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (N, L, K, mode, scheme, replicates per GPU per step, description)
    # (c3 / c4: 3072 = the batch the library forms by itself under its 3 GB workspace budget; a step = one batch)
    "c3": (100_000, 32, 8, 0, "centroid", 3072,
           "synthetic N=100k, 32 LVs x 8 MVs (P=256), Mode A, centroid, bootstrap (north-star headline config)"),
    "c3f": (100_000, 32, 8, 0, "factorial", 3072, "synthetic N=100k, 32 LVs x 8 MVs, Mode A, factorial, bootstrap"),
    "c4": (100_000, 32, 8, 1, "path", 3072, "synthetic N=100k, 32 LVs x 8 MVs, Mode B, path scheme, bootstrap"),
    "c5": (1_000_000, 64, 16, 0, "centroid", 512, "synthetic N=1M, 64 LVs x 16 MVs (P=1024), Mode A, centroid, bootstrap"),
    "c3n": (100_000, 32, 8, 0, "centroid", 1536,
            "synthetic N=100k, 32 LVs x 8 MVs, Scale.NUM (non-metric estimator), Mode A, centroid, bootstrap"),
    "c3s": (100_000, 32, 8, 0, "factorial", 1, "synthetic N=100k, 32 LVs x 8 MVs, Mode A, factorial scheme, SINGLE fit (BASELINE config 3)"),
    "c2": (250, 6, 0, 0, "centroid", 1000, "satisfaction 250x27, 6 LVs, Mode A, centroid, 1000 resamples (latency-bound)"),
}
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def load_workload(name):
    from plspm_b200.synth import make_synthetic
    N, L, K, mode, scheme, reps, desc = WORKLOADS[name]
    if name == "c2":
        g = np.load(os.path.join(ROOT, "tests", "golden", "satisfaction.npz"))
        return dict(X=np.ascontiguousarray(g["X"]), path=g["path"], blocks=[int(v) for v in g["block_sizes"]],
                    modes=[0] * 6, scheme=scheme, reps=reps, desc=desc, scaled=False)
    X, path = make_synthetic(N, L, K, seed=0)
    return dict(X=X, path=path, blocks=[K] * L, modes=[mode] * L, scheme=scheme, reps=reps, desc=desc, scaled=True,
                numeric=name.endswith("n"))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_one(rep):
    from oracle import plspm_oracle as orc
    w = _CPU["w"]
    idx = orc.philox_indices(0, rep, w["X"].shape[0])
    t = time.perf_counter()
    if w.get("numeric"):
        from oracle import plspm_oracle_nonmetric as onm
        res = onm.fit_num(w["X"][idx], w["blocks"], w["modes"], w["path"], w["scheme"])
        return time.perf_counter() - t, int(res["iterations"]), 0
    row, iters, status = orc.replicate_row(w["X"], idx, w["blocks"], w["modes"], w["path"], w["scheme"], w["scaled"])
    return time.perf_counter() - t, int(iters), int(status)


def cpu_fits_per_sec(w, n_fits, cores):
    """n_fits replicates of the same workload, one worker process per core (fork: X is shared)."""
    import multiprocessing as mp
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    _CPU["w"] = w
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_one, range(n_fits), chunksize=1)
    dt = time.perf_counter() - t0
    if limiter is not None:
        limiter.unregister() if hasattr(limiter, "unregister") else None
    assert all(r[2] == 0 for r in res)
    return n_fits / dt, dt, float(np.mean([r[1] for r in res]))


CPU_MAX_ELEMENTS = 3.0e7  # rows x columns of one CPU fit; larger workloads are timed on a row subsample


def cpu_workload(w):
    """(workload for the CPU arm, throughput scale, note).  The oracle's cost is linear in the rows, so a workload
    too large to time within the budget (c5: minutes per fit) is timed on its first rows and scaled, labelled."""
    N, P = w["X"].shape
    if N * P <= CPU_MAX_ELEMENTS:
        return w, 1.0, ""
    n_sub = int(CPU_MAX_ELEMENTS // P)
    return dict(w, X=np.ascontiguousarray(w["X"][:n_sub])), n_sub / N, (
        "; EXTRAPOLATED: timed on the first %d of %d rows and scaled by %d/%d (cost is linear in the rows)" % (n_sub, N, n_sub, N))


def cpu_sample_size(w, cores, budget_s):
    """Sample size for ~budget_s seconds: time one fit first."""
    _CPU["w"] = w
    t1 = _cpu_one(0)[0]
    per_round = max(t1, 1e-3)
    rounds = max(1, int(budget_s / per_round))
    return int(min(rounds * cores, 4096)), t1


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference ITSELF (baseline/_ref, staged by baseline/stage_reference.py) in a subprocess
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "STAGING.txt"))


def _ref_runner(workload, rows, replicates, processes, steps, warmup, timeout_s):
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py"), "--workload", workload, "--rows", str(int(rows)),
           "--replicates", str(int(replicates)), "--processes", str(int(processes)), "--steps", str(int(steps)),
           "--warmup", str(int(warmup))]
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference runner failed: " + r.stderr[-800:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def reference_fits_per_sec(workload, w, cores, step_budget_s, steps=1, warmup=0):
    """Times the reference's own bootstrap (plspm.bootstrap.Bootstrap, `cores` forked workers) on a bounded sample:
    as many rows of the workload as fit the per-step budget (two small single fits calibrate the reference's
    cost a + b * rows), one or more replicates per worker.  Returns (fits/s scaled to the full row count,
    description of the sample, seconds per step, reference single-fit seconds at the sampled rows)."""
    N, P = w["X"].shape
    n1 = int(min(N, max(250, 300_000 // P)))
    c1 = _ref_runner(workload, n1, 0, 1, 1, 0, 900)["single_fit_s"]
    if n1 >= N:
        a, b = c1, 0.0
    else:
        n2 = int(min(N, 3 * n1))
        c2 = _ref_runner(workload, n2, 0, 1, 1, 0, 900)["single_fit_s"]
        b = max((c2 - c1) / max(n2 - n1, 1), 1e-9)
        a = max(c1 - b * n1, 0.0)
    fit_budget = step_budget_s / 1.5  # workers contend for memory bandwidth; the parent polls once per second
    rows = N if a + b * N <= fit_budget else int(max(n1, min(N, (fit_budget - a) / b)))
    per_worker = int(max(1, fit_budget // max(a + b * rows, 1e-3)))
    reps = per_worker * cores
    res = _ref_runner(workload, rows, reps, cores, steps, warmup, 3600)
    total = sum(res["step_s"])
    scale = rows / N
    value = reps * len(res["step_s"]) / total * scale
    sample = ("plspm.bootstrap.Bootstrap of the reference (v0.5.7, statsmodels stand-in, pandas-3 one-liner): %d replicates "
              "per step on %d forked workers, %d of %d rows" % (reps, cores, rows, N))
    if rows < N:
        sample += ("; EXTRAPOLATED to %d rows by rows/N = %.4f (the reference's O(N) steps are linear in the rows; its "
                   "fixed per-fit overhead makes this flatter the reference)" % (N, scale))
    return value, sample, total / max(len(res["step_s"]), 1), res["single_fit_s"], rows


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w_full = load_workload(args.workload)
    cores = os.cpu_count() or 1
    n_steps = max(args.steps + args.warmup, 1)
    if reference_available():
        budget = min(60.0, max(8.0, 200.0 / n_steps))
        t0 = time.perf_counter()
        value, sample, step_s, single_s, rows = reference_fits_per_sec(args.workload, w_full, cores, budget, args.steps,
                                                                       args.warmup)
        kind, n_fits = "reference", 0
        note = ("the reference itself (baseline/_ref) through its own public API in a subprocess; reference single fit on "
                "%d rows: %.2f s; whole arm %.0f s" % (rows, single_s, time.perf_counter() - t0))
        ms_per_step = 1e3 * step_s
    else:  # labelled fallback: the oracle port (the staged reference did not travel)
        w, cpu_scale, cpu_note = cpu_workload(w_full)
        n_fits, t1 = cpu_sample_size(w, cores, budget_s=max(3.0, 60.0 / n_steps))
        for _ in range(args.warmup):
            cpu_fits_per_sec(w, max(cores, n_fits // 4), cores)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            cpu_fits_per_sec(w, n_fits, cores)
            total += n_fits
        el = time.perf_counter() - t0
        value = total / el * cpu_scale
        kind = "port"
        sample = "%d bootstrap fits per step on %d worker processes (oracle port, 1 BLAS thread each), same data/config%s" % (
            n_fits, cores, cpu_note)
        note = "baseline/_ref is missing: oracle port of the reference algorithm; single fit %.3f s on one core" % t1
        ms_per_step = 1e3 * el / args.steps
    line = {
        "impl": "reference", "metric": "bootstrap_fits_per_sec", "value": value, "unit": "fits/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, w_full, w_full["reps"]),
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": note,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, w, reps):
    N, P = w["X"].shape
    return {"workload": w["desc"], "name": name, "N": int(N), "P": int(P), "L": len(w["blocks"]),
            "mode": "B" if w["modes"][0] else "A", "scheme": w["scheme"], "scaled": bool(w["scaled"]), "numeric_scales": bool(w.get("numeric")),
            "tol": 1e-6, "max_iter": 100, "replicates_per_gpu_per_step": int(reps),
            "l2": "inputs larger than L2, no flush needed: per step the kernels stream the operands derived from X (%.0f MB: "
                  "the pre-scaled transposed copy and the fp16 images of the sign vote; X itself is %.1f MB fp64), %.0f MB of "
                  "int8 multiplicity images of the step's own resamples and the Gram partials (L2 is 126 MB)"
                  % (N * P * 8 / 1e6 + N * P * 2 * 1.5 / 1e6, N * P * 8 / 1e6, 2.0 * reps * N / 1e6)}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.split(",") for r in open(self.file.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.file.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit()]
        out["sm_mhz"] = float(np.median(sm)) if sm else None
        out["sm_max_mhz"] = float(rows[0][1]) if rows[0][1].strip().replace(".", "").isdigit() else None
        try:
            out["power_w_median"] = float(np.median([float(r[2]) for r in rows]))
        except Exception:
            pass
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for k, nme in enumerate(names):
            if any("Active" in r[3 + k] and "Not" not in r[3 + k] for r in rows):
                out["reasons"].append(nme)
        out["samples"] = len(rows)
        return out


def library_launches(prof):
    """Library (cuBLAS) GEMM launches inside the timed region.  Zero on the default routes (both contractions are
    this repo's tcgen05 kernels); only the A/B switches PLSPM_GRAM=cublas / PLSPM_VOTE=cublas bring the library back."""
    n = lambda k: int(prof.get(k, (0.0, 0))[1])
    lib = 0
    if os.environ.get("PLSPM_GRAM") == "cublas":
        lib += n("gram_i8") // 2 + max(0, n("colsum") - 1) // 2
    if os.environ.get("PLSPM_VOTE") == "cublas":
        lib += n("cross")
    return lib


def tensor_peaks():
    """Measured dense tensor-core peaks for the roofline denominators: bf16 from MEASURED_PEAKS.json (driver-written),
    int8 / fp16 from profiles/measured_tc_peaks.json (tools/tc_peak.py, library GEMMs on this pool's B200s)."""
    out = {}
    for fn in (os.path.join(ROOT, "MEASURED_PEAKS.json"), os.path.join(ROOT, "profiles", "measured_tc_peaks.json")):
        try:
            out.update(json.load(open(fn)))
        except Exception:
            pass
    return out


def ncu_metric(stage, metric, tag="r02"):
    """One metric of the committed ncu --set full capture of a stage's kernel (profiles/<stage>_<tag>_details.csv), or None."""
    try:
        for line in open(os.path.join(ROOT, "profiles", "%s_%s_details.csv" % (stage, tag))):
            f = line.rstrip("\n").split(",")
            if f[0] == metric:
                return float(f[-1])
    except Exception:
        pass
    return None


KERNEL_NAMES = {
    "gram": "gram_kernel<false> (fp64 weighted Gram tiles)",
    "gram_i8": "gram_mma_kernel (tcgen05 kind::i8: multiplicities x on-the-fly 48-bit digits of the pair products)",
    "finalize": "gram_finalize_kernel (int64 recombination -> fp64 moments)",
    "cross": "vote_mma_kernel (tcgen05 kind::f16: score MMA -> TMEM -> multiplicity scaling -> vote MMA)",
    "colsum": "counts8_image_kernel + vote_c8_image_kernel (int8 multiplicity tile images)",
    "scoregen": "scoregen_kernel", "conv": "conv_kernel", "solve": "solve_kernel", "counts": "counts_kernel",
    "reduce": "reduce_chunks_kernel", "scores": "scores_kernel", "upload": "upload kernels"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from plspm_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    engine.set_device(local)
    dev = torch.device("cuda", local)

    w = load_workload(args.workload)
    X = w["X"]
    N, P = X.shape
    reps = args.replicates or w["reps"]
    model = engine.Model(w["blocks"], w["modes"], w["path"], w["scaled"], numeric=bool(w.get("numeric")))
    data = engine.Data(model, X)
    n_out = model.n_out
    if args.workload.endswith("s"):
        return run_single_fit(args, w, model, data, engine)
    # every timed step writes its rows into its own slice; ONE all-gather of all slices ends the run (north_star)
    rows_dev = torch.empty((args.steps, reps, n_out), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, args.steps, reps, n_out), dtype=torch.float64, device=dev) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    step_id = [0]

    def step(slot):
        # global replicate ids: step-major, then rank: never reused, independent of world size
        begin = (step_id[0] * world + rank) * reps
        step_id[0] += 1
        _, status, iters = engine.bootstrap(model, data, w["scheme"], begin, reps, seed=0,
                                            out_device_ptr=rows_dev[slot].data_ptr())
        return begin, status, iters

    def finish_run():
        if world > 1:
            dist.all_gather_into_tensor(gathered, rows_dev)
        torch.cuda.synchronize()

    # nvidia-smi needs ~100 ms per sample and the timed region can be shorter than that: the sampler runs from
    # the warm-up through the timed steps and a tail of identical (untimed) steps, all under the same load
    sampler = ClockSampler(local) if rank == 0 else None
    if args.per_step_calls:
        for _ in range(args.warmup):
            step(0)
    elif args.warmup > 0:
        # the same call shape as the timed region (one call, W pipelined batches): the doubled workspace and the
        # page-locked staging of the pipeline are allocated here, not inside the timed region
        wbuf = torch.empty((args.warmup, reps, n_out), dtype=torch.float64, device=dev)
        engine.bootstrap(model, data, w["scheme"], (step_id[0] * world + rank * args.warmup) * reps, args.warmup * reps, seed=0,
                         out_device_ptr=wbuf.data_ptr())
        step_id[0] += args.warmup
        del wbuf
    finish_run()
    barrier()
    engine.profile_reset()
    t0 = time.perf_counter()
    iters_all, begins, bad = [], [], 0
    if args.per_step_calls:  # A/B: one library call per step (the device idles while the host turns around)
        for k in range(args.steps):
            begin, status, iters = step(k)
            begins.append(begin)
            iters_all.append(iters.astype(np.float64))
            bad += int((status != 0).sum())
    else:
        # ONE library call for the K steps, as a user bootstrapping K x reps replicates calls it: the library runs them as
        # K batches of `reps` and enqueues batch k + 1 before it waits for batch k (same kernels, same work per step)
        begin = (step_id[0] * world + rank * args.steps) * reps
        step_id[0] += args.steps
        _, status, iters = engine.bootstrap(model, data, w["scheme"], begin, args.steps * reps, seed=0,
                                            out_device_ptr=rows_dev.data_ptr())
        begins = [begin + k * reps for k in range(args.steps)]
        iters_all = [iters[k * reps:(k + 1) * reps].astype(np.float64) for k in range(args.steps)]
        bad += int((status != 0).sum())
    finish_run()
    barrier()
    elapsed = time.perf_counter() - t0
    prof = engine.profile_get()
    el_t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el_t, op=dist.ReduceOp.MAX)
    elapsed_max = float(el_t.item())
    rows_first = rows_dev[0].cpu().numpy() if rank == 0 else None
    rows_last = rows_dev[args.steps - 1].cpu().numpy() if rank == 0 else None
    # (the same count on every rank)
    tail_steps = int(min(200, 0.6 / max(elapsed_max / args.steps, 1e-4) + 1))
    for _ in range(tail_steps):
        step(0)
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps + %d identical untimed steps (same load)" % tail_steps
    value = world * reps * args.steps / elapsed_max
    iters_cat = np.concatenate(iters_all)

    # ---- end to end: host X (pinned) -> upload -> bootstrap -> rows back on the host ---------
    Xp = engine.pinned_empty(X.shape)
    Xp[...] = X
    out_host = engine.pinned_empty((reps, n_out))
    warm_ms = []
    for s in range(2):  # warm-up: the buffer pool is populated here
        ts = time.perf_counter()
        engine.bootstrap_host(model, Xp, w["scheme"], s * reps, reps, seed=1, out=out_host)
        warm_ms.append(1e3 * (time.perf_counter() - ts))
    # enough steps for ~2 s of end-to-end work (3 .. 12): a single host hiccup must not decide the figure
    warm_t = torch.tensor([warm_ms[-1]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(warm_t, op=dist.ReduceOp.MAX)  # the same step count on every rank
    e2e_steps = int(max(3, min(12, round(2000.0 / max(float(warm_t.item()), 1.0)))))
    barrier()
    e2e_each = []
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ts = time.perf_counter()
        engine.bootstrap_host(model, Xp, w["scheme"], (10_000 + s * world + rank) * reps, reps, seed=0, out=out_host)
        e2e_each.append(1e3 * (time.perf_counter() - ts))
    barrier()
    e2e_el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_el, op=dist.ReduceOp.MAX)
    e2e_value = world * reps * e2e_steps / float(e2e_el.item())

    if rank == 0:
        # ---- parity self-check: four rows of the TIMED run against the CPU oracle (1e-6, same iteration counts) -----
        parity_checked, parity_failed = 0, []
        if not args.no_parity and not w.get("numeric"):
            from oracle import plspm_oracle as orc
            Xo = X if N * P <= 6.0e7 else None  # (c5: an oracle fit takes minutes; parity at that size is a -m gpu test)
            picks = [(0, 0, rows_first), (0, reps - 1, rows_first), (args.steps - 1, 0, rows_last), (args.steps - 1, reps - 1, rows_last)]
            for k, b, rows in (picks if Xo is not None else []):
                idx = orc.philox_indices(0, begins[k] + b, N)
                ref, it, st = orc.replicate_row(Xo, idx, w["blocks"], w["modes"], w["path"], w["scheme"], w["scaled"])
                ok = st == 0 and int(iters_all[k][b]) == it and np.allclose(rows[b], ref, rtol=1e-6, atol=1e-9)
                parity_checked += 1
                if not ok:
                    parity_failed.append(int(begins[k] + b))
        total_ms = sum(v[0] for v in prof.values())
        stages = {k: v[0] / args.steps for k, v in prof.items() if v[1]}
        top = max((k for k in prof if prof[k][1]), key=lambda k: prof[k][0])
        top_ms, top_n = prof[top]
        peak_hbm, peak_src = hbm_peak()
        peaks = tensor_peaks()
        # algorithmic HBM bytes of the reference dataflow, (n_iter + 2) N P 8 per fit
        alg_bytes = float((iters_cat + 2.0).sum()) * N * P * 8.0
        stream_ms = sum(prof.get(k, (0.0, 0))[0] for k in ("gram", "gram_i8", "cross", "colsum", "conv", "finalize", "scoregen"))
        survey_gbs = alg_bytes / 1e9 / (stream_ms / 1e3) if stream_ms > 0 else 0.0
        # measured DRAM traffic of the step's kernels (ncu, per launch; profiles/kernel_traffic.json, tools/ncu_summary.py)
        traffic_tab = {}
        try:
            tab = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
            # c4 runs the same kernels on the same data as c3 (only the solver's mode / scheme differ): its capture serves both
            traffic_tab = tab.get(args.workload) or (tab.get("c3", {}) if args.workload == "c4" else {})
        except Exception:
            pass
        scale = reps / float(traffic_tab.get("_replicates_per_launch", reps))
        step_traffic = sum(float(traffic_tab[k]) * scale * prof[k][1] / args.steps for k in traffic_tab if k in prof and prof[k][1])
        frac_hw = (step_traffic / 1e9) / (total_ms / args.steps / 1e3) / peak_hbm if step_traffic and total_ms else None
        # tensor work of the two tcgen05 kernels (useful operations only: padding rows / columns not counted)
        gram_ms, vote_ms = prof.get("gram_i8", (0.0, 0))[0], prof.get("cross", (0.0, 0))[0]
        n_moments = model.n_pair_columns + P
        gram_ops = 2.0 * 6 * n_moments * reps * N * args.steps           # u8 digits x s8 multiplicities, 2 ops per MAC
        vote_ops = 2.0 * reps * len(w["blocks"]) * N * P * args.steps    # fp16 vote MMA (the score MMA adds 1/16 of it)
        # int8 denominator: the NOMINAL dense 8-bit tensor peak (B200_PROFILING.md: 4.5 P op/s).  MEASURED_PEAKS.json has bf16
        # only, and the library int8 GEMM measured by tools/tc_peak.py (2.7 P sustained / 3.3 P burst) is SLOWER than this
        # kernel, so it cannot serve as a peak; it is reported beside it (int8_library_tops).
        i8_lib = float(peaks.get("int8_tops", 2.0 * float(peaks.get("bf16_tflops_sustained", 1402.0))))
        i8_peak = 4500.0
        f16_peak = float(peaks.get("fp16_tflops", peaks.get("bf16_tflops_sustained", 1402.0)))
        mma_route = model.n_pair_columns > 0 and os.environ.get("PLSPM_GRAM") not in ("cublas", "fp64") and gram_ms > 0
        if top == "cross" and not model.full_tiles:
            roof = {"bound": "tensor", "achieved": vote_ops / (vote_ms / 1e3) / 1e12, "peak": f16_peak, "unit": "TFLOP/s",
                    "peak_source": "measured fp16 library GEMM (profiles/measured_tc_peaks.json), else MEASURED_PEAKS.json bf16 sustained"}
        elif top == "gram_i8" and mma_route:
            roof = {"bound": "tensor", "achieved": gram_ops / (gram_ms / 1e3) / 1e12, "peak": i8_peak, "unit": "TFLOP/s",
                    "peak_source": "NOMINAL dense 8-bit tensor peak, T op/s (B200_PROFILING.md); the measured library int8 GEMM (int8_library_tops, tools/tc_peak.py) is slower than this kernel"}
        else:
            roof = {"bound": "hbm", "achieved": survey_gbs, "peak": peak_hbm, "unit": "GB/s", "peak_source": peak_src}
        roof["frac"] = roof["achieved"] / roof["peak"] if roof["peak"] else None
        roof["traffic"] = float(traffic_tab[top]) * scale if top in traffic_tab else None
        roof.update({
            "kernel": KERNEL_NAMES.get(top, top), "kernel_stage": top, "kernel_share_of_step": top_ms / max(total_ms, 1e-9),
            "kernel_launch_ms": top_ms / max(top_n, 1), "kernel_launches": top_n,
            "gram_int8_tops": gram_ops / (gram_ms / 1e3) / 1e12 if mma_route else None,
            "vote_fp16_tflops": vote_ops / (vote_ms / 1e3) / 1e12 if vote_ms > 0 and not model.full_tiles else None,
            "int8_peak_tops": i8_peak, "int8_library_tops": i8_lib, "fp16_peak_tflops": f16_peak,
            "tensor_pipe_active_pct_ncu": ncu_metric(top, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "hbm": {"frac_hw": frac_hw, "dram_bytes_per_step": step_traffic or None, "peak": peak_hbm, "peak_source": peak_src,
                    "survey_8d_achieved_gbs": survey_gbs, "survey_8d_frac": survey_gbs / peak_hbm,
                    "algorithmic_bytes_per_step": alg_bytes / args.steps,
                    "note": "frac_hw = measured DRAM bytes of the step's kernels (ncu, profiles/kernel_traffic.json) / step time / "
                            "measured copy bandwidth: the step is tensor-bound, not HBM-bound.  survey_8d_* is SURVEY 8(d)'s "
                            "prescribed figure, (n_iter+2)*N*P*8 bytes per fit over the streaming stages; the engine never re-reads "
                            "X per iteration or per replicate (the iteration runs on second moments of the whole batch), so that "
                            "figure exceeds the peak by construction and is NOT a hardware fraction"},
            "tile_set": "full" if model.full_tiles else "sparse",
            "sign_vote": ("n/a (non-metric estimator)" if w.get("numeric") else "n/a (full tile set)" if model.full_tiles else
                          "fused tcgen05 vote kernel with error bound, %d replicates redone exactly" % engine.redo_count()),
        })
        lib = library_launches(prof)
        line = {
            "metric": "bootstrap_fits_per_sec", "value": value, "unit": "fits/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, w, reps),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(N * P * 8),
                    "d2h_bytes_per_step": int(reps * (n_out * 8 + 8)), "steps": e2e_steps,
                    "ms_each_step": [round(v, 2) for v in e2e_each]},
            "gpu_launches": sum(int(v[1]) for v in prof.values()) - lib, "library_gemm_launches": lib,
            "parity_checked": parity_checked, "parity_failed": parity_failed,
            "roofline": roof,
            "stages_ms_per_step": stages,
            "timing": "value = replicates / wall clock around the K steps + the one all-gather (barrier + synchronize on both "
                      "sides, max over ranks).  %s.  CUDA-event sum of the step's kernels on the library's stream: %.3f ms per "
                      "step" % ("one library call per step, each ending in a stream synchronize" if args.per_step_calls else
                                "The K steps are ONE library call of K x %d replicates = K batches; the library enqueues "
                                "batch k + 1 before it waits for batch k" % reps, total_ms / args.steps),
            "calls": "per step" if args.per_step_calls else "one call, K pipelined batches",
            "mean_iterations": float(iters_cat.mean()), "failed_replicates": bad,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            if reference_available():
                fps, sample, step_s, single_s, rows = reference_fits_per_sec(args.workload, w, cores, args.cpu_seconds)
                line["cpu_baseline"] = {"value": fps, "unit": "fits/s", "cores": cores, "kind": "reference",
                                        "sample": sample + "; one step of %.1f s" % step_s,
                                        "reference_single_fit_s": single_s, "reference_single_fit_rows": rows,
                                        # quirk Q1: the reference runs its weight iteration twice per fit
                                        # (estimator.py:36,43,52); `value` is the reference as it is, this is the same
                                        # figure with a whole duplicate fit discounted (an upper bound: treatment and the
                                        # inner model run once)
                                        "value_without_q1_duplicate": 2.0 * fps}
            else:
                w_cpu, cpu_scale, cpu_note = cpu_workload(w)
                n_fits, t1 = cpu_sample_size(w_cpu, cores, budget_s=args.cpu_seconds)
                fps, dt, mean_it = cpu_fits_per_sec(w_cpu, n_fits, cores)
                line["cpu_baseline"] = {"value": fps * cpu_scale, "unit": "fits/s", "cores": cores, "kind": "port",
                                        "sample": "%d bootstrap fits of the same workload in %.1f s on %d worker processes "
                                                  "(oracle port, 1 BLAS thread each; baseline/_ref missing)%s" % (
                                                      n_fits, dt, cores, cpu_note)}
        print(json.dumps(line), flush=True)
        if parity_failed:
            print("PARITY FAILED for global replicates %s" % parity_failed, file=sys.stderr, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_single_fit(args, w, model, data, engine):
    """c3s: BASELINE config 3 taken literally -- ONE fit on the resident data (factorial scheme), latency in ms.
    value = device + call latency of plspm_fit without the N x L score download; e2e = upload + fit + scores to the host."""
    import torch
    X = w["X"]
    N, P = X.shape
    for _ in range(args.warmup):
        engine.fit(model, data, w["scheme"], want_scores=False)
    torch.cuda.synchronize()
    engine.profile_reset()
    lat = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        got = engine.fit(model, data, w["scheme"], want_scores=False)
        lat.append(1e3 * (time.perf_counter() - ts))
    prof = engine.profile_get()
    Xp = engine.pinned_empty(X.shape)
    Xp[...] = X
    e2e = []
    for s in range(max(3, args.steps)):
        ts = time.perf_counter()
        d2 = engine.Data(model, Xp)
        res = engine.fit(model, d2, w["scheme"], want_scores=True)
        d2.close()
        e2e.append(1e3 * (time.perf_counter() - ts))
    parity = 0
    if not args.no_parity:
        from oracle import plspm_oracle as orc
        ref = orc.fit(X, w["blocks"], w["modes"], w["path"], w["scheme"], w["scaled"])
        ok = res["iterations"] == ref["iterations"] and all(
            np.allclose(res[k], ref[k], rtol=1e-6, atol=1e-8) for k in ("weights", "scores", "path_coefficients"))
        parity = 1 if ok else -1
    peak_hbm, peak_src = hbm_peak()
    stages = {k: v[0] / args.steps for k, v in prof.items() if v[1]}
    dev_ms = sum(stages.values())
    alg = (got["iterations"] + 2) * N * P * 8.0
    line = {
        "metric": "single_fit_latency_ms", "value": float(np.median(lat)), "unit": "ms", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(lat)), "higher_is_better": False, "scaling": "replicas only",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, w, 1),
        "e2e": {"value": float(np.median(e2e)), "unit": "ms", "h2d_bytes_per_step": int(N * P * 8),
                "d2h_bytes_per_step": int(N * len(w["blocks"]) * 8), "ms_each_step": [round(v, 2) for v in e2e]},
        "gpu_launches": sum(int(v[1]) for v in prof.values()), "library_gemm_launches": 0, "parity_checked": parity,
        "roofline": {"bound": "hbm", "achieved": (N * P * 8.0 * 2) / 1e9 / (dev_ms / 1e3), "peak": peak_hbm, "unit": "GB/s",
                     "frac": (N * P * 8.0 * 2) / 1e9 / (dev_ms / 1e3) / peak_hbm, "traffic": None, "peak_source": peak_src,
                     "note": "two passes over the resident fp64 matrix (moments, then cross moments for the sign vote) are the "
                             "compulsory traffic of one fit in the covariance formulation; SURVEY 8(d) figure: %.1f GB/s" % (
                                 alg / 1e9 / (dev_ms / 1e3))},
        "stages_ms_per_step": stages, "iterations": int(got["iterations"]),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--replicates", type=int, default=0, help="replicates per GPU per step (default: per workload)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--per-step-calls", action="store_true", help="one library call per timed step instead of one call for all K steps")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of four timed replicate rows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
