#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What it does
  * copies /root/reference/plspm to a temp dir and changes ONE expression that
    pandas 3 rejects (inner_model.py:75 `path.loc[dv,]` -> `path.loc[dv]`;
    SURVEY.md §8(c) "Accommodation 2"); nothing else is touched;
  * puts oracle/_shim (statsmodels stand-in) on sys.path;
  * runs the reference's own public API (`plspm.plspm.Plspm`,
    `Estimator.estimate`, `InnerModel`) on the satisfaction data set and on
    small synthetic models, for every scheme x mode x scaled combination, and
    stores inputs + outputs as .npz;
  * stores the R-generated golden CSV values of the reference's own tests
    (tests/data/satisfaction.*.csv) next to them as known-answer vectors;
  * replays the per-replicate body of BootstrapProcess.run
    (bootstrap.py:54-66) with INJECTED resample indices so bootstrap parity
    can be checked replicate by replicate (the reference itself is unseeded).

The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PLSPM_REFERENCE", "/root/reference")


def _stage_reference():
    tmp = tempfile.mkdtemp(prefix="plspm_ref_")
    shutil.copytree(os.path.join(REF, "plspm"), os.path.join(tmp, "plspm"))
    p = os.path.join(tmp, "plspm", "inner_model.py")
    src = open(p).read()
    assert "path.loc[dv,][path.loc[dv,] == 1]" in src
    open(p, "w").write(src.replace("path.loc[dv,][path.loc[dv,] == 1]", "path.loc[dv][path.loc[dv] == 1]"))
    sys.path.insert(0, tmp)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
    return tmp


_stage_reference()
sys.path.append(os.path.join(ROOT, "plspm-python_b200"))  # after the staged reference: only plspm_b200.synth is used

import plspm.config as c  # noqa: E402  (the REFERENCE package, from the staged copy)
import plspm.weights as ref_weights  # noqa: E402
import plspm.inner_model as ref_im  # noqa: E402
from plspm.estimator import Estimator  # noqa: E402
from plspm.mode import Mode  # noqa: E402
from plspm.plspm import Plspm  # noqa: E402
from plspm.scheme import Scheme  # noqa: E402

assert os.path.realpath(c.__file__).startswith(os.path.realpath(tempfile.gettempdir()))

# count iterate() calls (two identical calculate() runs per fit, estimator.py:39,52)
_ITER_CALLS = [0]
_orig_iterate = ref_weights._MetricWeights.iterate


def _counting_iterate(self, scheme):
    _ITER_CALLS[0] += 1
    return _orig_iterate(self, scheme)


ref_weights._MetricWeights.iterate = _counting_iterate

SCHEMES = {"centroid": Scheme.CENTROID, "factorial": Scheme.FACTORIAL, "path": Scheme.PATH}
SAT_LVS = ["IMAG", "EXPE", "QUAL", "VAL", "SAT", "LOY"]
SAT_PREFIX = {"IMAG": "imag", "EXPE": "expe", "QUAL": "qual", "VAL": "val", "SAT": "sat", "LOY": "loy"}


def satisfaction_structure():
    s = c.Structure()
    s.add_path(["IMAG"], ["EXPE", "SAT", "LOY"])
    s.add_path(["EXPE"], ["QUAL", "VAL", "SAT"])
    s.add_path(["QUAL"], ["VAL", "SAT"])
    s.add_path(["VAL"], ["SAT"])
    s.add_path(["SAT"], ["LOY"])
    return s.path()


def run_reference(data: pd.DataFrame, path: pd.DataFrame, blocks: dict, modes: dict, scheme, scaled: bool,
                  tol=1e-6, iterations=100):
    """Runs Plspm() and returns plain arrays, everything in path-LV / ODM order."""
    config = c.Config(path, scaled=scaled)
    for lv in blocks:
        config.add_lv(lv, modes[lv], *[c.MV(m) for m in blocks[lv]])
    _ITER_CALLS[0] = 0
    calc = Plspm(data, config, scheme, iterations, tol)
    n_iter = _ITER_CALLS[0] // 2
    lvs = list(path)
    mvs = [m for lv in lvs for m in blocks[lv]]
    om = calc.outer_model()
    eff = calc.effects()
    return dict(
        lvs=np.array(lvs), mvs=np.array(mvs),
        weights=om.loc[mvs, "weight"].values.astype(np.float64),
        loadings=om.loc[mvs, "loading"].values.astype(np.float64),
        communality=om.loc[mvs, "communality"].values.astype(np.float64),
        redundancy=om.loc[mvs, "redundancy"].values.astype(np.float64),
        scores=calc.scores().loc[:, lvs].values.astype(np.float64),
        path_coefficients=calc.path_coefficients().loc[lvs, lvs].values.astype(np.float64),
        r_squared=calc.inner_summary().loc[lvs, "r_squared"].values.astype(np.float64),
        crossloadings=calc.crossloadings().loc[mvs, lvs].values.astype(np.float64),
        effects_from=np.array([str(v) for v in eff["from"]]), effects_to=np.array([str(v) for v in eff["to"]]),
        effects_direct=eff["direct"].values.astype(np.float64),
        effects_indirect=eff["indirect"].values.astype(np.float64),
        effects_total=eff["total"].values.astype(np.float64),
        inner_model_index=np.array(list(calc.inner_model().index)),
        inner_model=calc.inner_model().loc[:, ["estimate", "std error", "t", "p>|t|"]].values.astype(np.float64),
        iterations=np.int64(n_iter),
    )


def run_reference_replicates(data, path, blocks, modes, scheme, scaled, idx, tol=1e-6, iterations=100):
    """Per-replicate body of BootstrapProcess.run (bootstrap.py:54-66) with injected indices."""
    config = c.Config(path, scaled=scaled)
    for lv in blocks:
        config.add_lv(lv, modes[lv], *[c.MV(m) for m in blocks[lv]])
    filtered = config.filter(data)
    n = filtered.shape[0]
    correction = np.sqrt(n / (n - 1))
    calculator = ref_weights.WeightsCalculatorFactory(config, iterations, tol, correction, scheme)
    estimator = Estimator(config)
    lvs = list(path)
    mvs = [m for lv in lvs for m in blocks[lv]]
    odm = config.odm(config.path())
    B = idx.shape[0]
    L, P = len(lvs), len(mvs)
    out = dict(weights=np.full((B, P), np.nan), loadings=np.full((B, P), np.nan), r_squared=np.full((B, L), np.nan),
               path_coefficients=np.full((B, L, L), np.nan), total_effects=np.full((B, L, L), np.nan),
               iterations=np.zeros(B, dtype=np.int64), ok=np.zeros(B, dtype=np.int8))
    for b in range(B):
        try:
            _ITER_CALLS[0] = 0
            fd, sc, w = estimator.estimate(calculator, filtered.iloc[idx[b], :])
            im = ref_im.InnerModel(config.path(), sc)
            out["iterations"][b] = _ITER_CALLS[0] // 2
            out["weights"][b] = w.loc[mvs, "weight"].values
            out["r_squared"][b] = im.r_squared().loc[lvs].values
            pc = im.path_coefficients().loc[lvs, lvs].values.astype(np.float64)
            out["path_coefficients"][b] = pc
            tot = np.zeros((L, L))
            eff = im.effects()
            for f, t, v in zip(eff["from"], eff["to"], eff["total"]):
                tot[lvs.index(t), lvs.index(f)] = v
            out["total_effects"][b] = tot
            load = (sc.apply(lambda s: fd.corrwith(s)) * odm).sum(axis=1)
            out["loadings"][b] = load.loc[mvs].values
            out["ok"][b] = 1
        except Exception as e:  # mirrors bootstrap.py:67-68 (replicate dropped)
            print("replicate", b, "failed in reference:", repr(e))
    return out


def flatten(prefix, d):
    return {prefix + "/" + k: v for k, v in d.items()}


def main():
    from plspm_b200.synth import make_synthetic
    tdata = os.path.join(REF, "tests", "data")
    sat = pd.read_csv(os.path.join(tdata, "satisfaction.csv"), index_col=0)
    path = satisfaction_structure()
    lvs = list(path)
    blocks = {lv: [m for m in sat.columns if m.startswith(SAT_PREFIX[lv])] for lv in lvs}
    mvs = [m for lv in lvs for m in blocks[lv]]

    # ---- satisfaction: inputs + R golden CSVs + reference outputs ----------------------------
    store = {"X": sat.loc[:, mvs].values.astype(np.float64), "mvs": np.array(mvs), "lvs": np.array(lvs),
             "path": path.loc[lvs, lvs].values.astype(np.int8),
             "block_sizes": np.array([len(blocks[lv]) for lv in lvs], dtype=np.int32)}
    # R golden vectors (reference tests/test_regression_metric.py:43-94)
    store["R/scores"] = pd.read_csv(os.path.join(tdata, "satisfaction.scores.csv")).loc[:, lvs].values
    for tag, fn in (("centroid", "satisfaction.outer-model.csv"), ("path", "satisfaction.outer-model-path.csv"),
                    ("factorial", "satisfaction.outer-model-factorial.csv")):
        om = pd.read_csv(os.path.join(tdata, fn), index_col=0).loc[mvs]
        for col in ("weight", "loading", "communality", "redundancy"):
            store["R/%s/%s" % (tag, col)] = om[col].values.astype(np.float64)
    eff = pd.read_csv(os.path.join(tdata, "satisfaction.effects.csv"), index_col=0)
    store["R/effects_from"] = np.array([str(v) for v in eff["from"]])
    store["R/effects_to"] = np.array([str(v) for v in eff["to"]])
    for col in ("direct", "indirect", "total"):
        store["R/effects_" + col] = eff[col].values.astype(np.float64)
    cl = pd.read_csv(os.path.join(tdata, "satisfaction.crossloadings.csv"), index_col=0).loc[mvs, lvs]
    store["R/crossloadings"] = cl.values.astype(np.float64)
    isum = pd.read_csv(os.path.join(tdata, "satisfaction.inner-summary.csv"), index_col=0).loc[lvs]
    for col in ("r_squared", "block_communality", "mean_redundancy", "ave"):
        store["R/inner_summary/" + col] = isum[col].values.astype(np.float64)
    imod = pd.read_csv(os.path.join(tdata, "satisfaction.inner-model.csv"), index_col=0)
    store["R/inner_model_SAT_from"] = np.array(list(imod.index))
    store["R/inner_model_SAT"] = imod.values.astype(np.float64)
    isb = pd.read_csv(os.path.join(tdata, "satisfaction.modeb.inner-summary.csv"), index_col=0).loc[lvs]
    isb.columns = [col.lower() for col in isb.columns]
    for col in ("r_squared", "block_communality", "mean_redundancy"):
        store["R/modeb/inner_summary/" + col] = isb[col].values.astype(np.float64)
    store["R/gof"] = np.float64(0.609741624338411)  # test_regression_metric.py:80
    uni = pd.read_csv(os.path.join(tdata, "satisfaction_unidim.csv"), index_col=0).loc[lvs]
    for col in ("mvs", "cronbach_alpha", "dillon_goldstein_rho", "eig_1st", "eig_2nd"):
        store["R/unidim/" + col] = uni[col].values.astype(np.float64)
    store["index"] = np.array([str(v) for v in sat.index])
    for tag in ("weights", "loadings", "paths", "rsquared", "total_effects"):
        bt = pd.read_csv(os.path.join(tdata, "satisfaction_boot_%s.csv" % tag), index_col=0)
        store["R/boot/%s/index" % tag] = np.array(list(bt.index))
        store["R/boot/%s/columns" % tag] = np.array(list(bt.columns))
        store["R/boot/%s/values" % tag] = bt.values.astype(np.float64)

    for sname, scheme in SCHEMES.items():
        for mname, mode in (("A", Mode.A), ("B", Mode.B)):
            for scaled in (False, True):
                tag = "ref/%s/%s/%s" % (sname, mname, "scaled" if scaled else "unscaled")
                res = run_reference(sat, path, blocks, {lv: mode for lv in lvs}, scheme, scaled)
                print(tag, "iterations", int(res["iterations"]))
                store.update(flatten(tag, res))
    # mixed modes (A for reflective, B for two blocks)
    mixed = {lv: (Mode.B if lv in ("IMAG", "VAL") else Mode.A) for lv in lvs}
    store.update(flatten("ref/path/mixed/scaled", run_reference(sat, path, blocks, mixed, Scheme.PATH, True)))
    store["mixed_modes"] = np.array([1 if mixed[lv] == Mode.B else 0 for lv in lvs], dtype=np.int8)

    # bootstrap replicates with injected indices (C2 parity)
    idx_all = np.random.default_rng(1234).integers(0, 250, (1000, 250), dtype=np.int32)
    nrep = 48
    for sname, mname, scaled in (("centroid", "A", False), ("path", "B", True), ("factorial", "A", True)):
        mode = Mode.A if mname == "A" else Mode.B
        tag = "boot/%s/%s/%s" % (sname, mname, "scaled" if scaled else "unscaled")
        r = run_reference_replicates(sat, path, blocks, {lv: mode for lv in lvs}, SCHEMES[sname], scaled,
                                     idx_all[:nrep])
        print(tag, "ok", int(r["ok"].sum()), "iters", np.bincount(r["iterations"]))
        store.update(flatten(tag, r))
    store["boot/n_replicates"] = np.int64(nrep)
    assert not [k for k, v in store.items() if np.asarray(v).dtype == object]
    np.savez_compressed(os.path.join(HERE, "satisfaction.npz"), **store)

    # ---- small synthetic models --------------------------------------------------------------
    syn = {}
    cases = [
        ("syn_a", dict(N=400, L=5, K=3, seed=11, reverse=()), "centroid", "A", True),
        ("syn_b", dict(N=400, L=5, K=3, seed=11, reverse=(1, 3)), "centroid", "A", False),
        ("syn_c", dict(N=300, L=4, K=4, seed=5, reverse=(2,)), "factorial", "A", True),
        ("syn_d", dict(N=600, L=6, K=2, seed=7, reverse=()), "path", "B", True),
        ("syn_e", dict(N=350, L=7, K=5, seed=3, reverse=(0,)), "path", "A", False),
        ("syn_f", dict(N=2000, L=9, K=8, seed=9, reverse=(4,)), "factorial", "B", False),
        ("syn_g", dict(N=500, L=3, K=11, seed=21, reverse=()), "centroid", "B", True),
        ("syn_h", dict(N=200, L=2, K=3, seed=2, reverse=()), "centroid", "A", True),
    ]
    for name, g, sname, mname, scaled in cases:
        X, pm = make_synthetic(g["N"], g["L"], g["K"], g["seed"], reverse_blocks=g["reverse"])
        names = ["lv%02d" % j for j in range(g["L"])]
        cols = ["x%02d_%02d" % (j, k) for j in range(g["L"]) for k in range(g["K"])]
        df = pd.DataFrame(X, columns=cols)
        pdf = pd.DataFrame(pm.astype(int), index=names, columns=names)
        blk = {names[j]: cols[j * g["K"]:(j + 1) * g["K"]] for j in range(g["L"])}
        mode = Mode.A if mname == "A" else Mode.B
        res = run_reference(df, pdf, blk, {n: mode for n in names}, SCHEMES[sname], scaled)
        print(name, sname, mname, scaled, "iterations", int(res["iterations"]))
        syn.update(flatten(name, res))
        syn[name + "/gen"] = np.array([g["N"], g["L"], g["K"], g["seed"]], dtype=np.int64)
        syn[name + "/reverse"] = np.array(g["reverse"], dtype=np.int64)
        syn[name + "/scheme"] = np.array(sname)
        syn[name + "/mode"] = np.array(mname)
        syn[name + "/scaled"] = np.int8(scaled)
        if name in ("syn_b", "syn_d"):
            idx = np.random.default_rng(77).integers(0, g["N"], (6, g["N"]), dtype=np.int32)
            r = run_reference_replicates(df, pdf, blk, {n: mode for n in names}, SCHEMES[sname], scaled, idx)
            syn.update(flatten(name + "/boot", r))
            syn[name + "/boot/idx"] = idx
    # ---- higher-order constructs: stage-1 path expansion (estimator.py:60-74).  The reference cannot run
    # the two-stage estimation itself on METRIC data (stage 2 raises "matrices are not aligned" at
    # weights.py:30), so only the path expansion is pinned here.
    st = c.Structure()
    st.add_path(["Expectation", "Quality"], ["Satisfaction"])
    st.add_path(["Satisfaction"], ["Complaints", "Loyalty"])
    cfg = c.Config(st.path())
    cfg.add_higher_order("Satisfaction", Mode.A, ["Image", "Value"])
    fs = Estimator(cfg).hoc_path_first_stage(cfg)
    syn["hoc/first_stage_lvs"] = np.array([str(v) for v in fs.index])
    syn["hoc/first_stage_path"] = fs.values.astype(np.int8)
    # ---- nonmetric path with numeric scales (groundwork for f3): russa + mobi, Scale.NUM --------------
    from plspm.scale import Scale
    nm = {}
    russa = pd.read_csv(os.path.join(tdata, "russa.csv"), index_col=0)
    st = c.Structure()
    st.add_path(["AGRI", "IND"], ["POLINS"])
    rpath = st.path()
    rblocks = {"AGRI": ["gini", "rent", "farm"], "IND": ["gnpr", "labo"], "POLINS": ["ecks", "death", "demo", "inst"]}
    rlvs = list(rpath)
    rmvs = [m for lv in rlvs for m in rblocks[lv]]
    nm["russa/X"] = russa.loc[:, rmvs].values.astype(np.float64)
    nm["russa/path"] = rpath.loc[rlvs, rlvs].values.astype(np.int8)
    nm["russa/block_sizes"] = np.array([len(rblocks[lv]) for lv in rlvs], dtype=np.int32)
    nm["russa/lvs"] = np.array(rlvs)
    nm["russa/mvs"] = np.array(rmvs)
    for sname, scheme in SCHEMES.items():
        for mname, mode in (("A", Mode.A), ("B", Mode.B)):
            cfg = c.Config(rpath, default_scale=Scale.NUM)
            for lv in rlvs:
                cfg.add_lv(lv, mode, *[c.MV(m) for m in rblocks[lv]])
            calc = Plspm(russa, cfg, scheme, 100, 1e-7)
            tag = "russa/%s/%s/" % (sname, mname)
            om = calc.outer_model()
            nm[tag + "weights"] = om.loc[rmvs, "weight"].values.astype(np.float64)
            nm[tag + "loadings"] = om.loc[rmvs, "loading"].values.astype(np.float64)
            nm[tag + "scores"] = calc.scores().loc[:, rlvs].values.astype(np.float64)
            nm[tag + "path_coefficients"] = calc.path_coefficients().loc[rlvs, rlvs].values.astype(np.float64)
            nm[tag + "crossloadings"] = calc.crossloadings().loc[rmvs, rlvs].values.astype(np.float64)
    for tag, fn in (("centroid", "russa.outer_model.csv"), ("path", "russa.outer_model_path.csv"),
                    ("factorial", "russa.outer_model_factorial.csv")):
        om = pd.read_csv(os.path.join(tdata, fn), index_col=0).loc[rmvs]
        nm["R/russa/%s/weight" % tag] = om["weight"].values.astype(np.float64)
        nm["R/russa/%s/loading" % tag] = om["loading"].values.astype(np.float64)
    nm["R/russa/scores"] = pd.read_csv(os.path.join(tdata, "russa.scores.csv"), index_col=0).loc[:, rlvs].values
    mobi = pd.read_csv(os.path.join(tdata, "mobi.csv"), index_col=0)
    st = c.Structure()
    st.add_path(["Expectation", "Quality"], ["Loyalty"])
    st.add_path(["Image"], ["Expectation"])
    st.add_path(["Complaints"], ["Loyalty"])
    mpath = st.path()
    mlvs = list(mpath)
    prefix = {"Expectation": "CUEX", "Quality": "PERQ", "Loyalty": "CUSL", "Image": "IMAG", "Complaints": "CUSCO"}
    mmode = {"Expectation": Mode.A, "Quality": Mode.B, "Loyalty": Mode.A, "Image": Mode.A, "Complaints": Mode.A}
    mblocks = {lv: [m for m in mobi.columns if m.startswith(prefix[lv])] for lv in mlvs}
    mmvs = [m for lv in mlvs for m in mblocks[lv]]
    cfg = c.Config(mpath, default_scale=Scale.NUM)
    for lv in mlvs:
        cfg.add_lv(lv, mmode[lv], *[c.MV(m) for m in mblocks[lv]])
    calc = Plspm(mobi, cfg, Scheme.PATH, 100, 1e-8)
    nm["mobi/X"] = mobi.loc[:, mmvs].values.astype(np.float64)
    nm["mobi/path"] = mpath.loc[mlvs, mlvs].values.astype(np.int8)
    nm["mobi/block_sizes"] = np.array([len(mblocks[lv]) for lv in mlvs], dtype=np.int32)
    nm["mobi/modes"] = np.array([0 if mmode[lv] == Mode.A else 1 for lv in mlvs], dtype=np.int8)
    nm["mobi/weights"] = calc.outer_model().loc[mmvs, "weight"].values.astype(np.float64)
    nm["mobi/loadings"] = calc.outer_model().loc[mmvs, "loading"].values.astype(np.float64)
    nm["mobi/path_coefficients"] = calc.path_coefficients().loc[mlvs, mlvs].values.astype(np.float64)
    som = pd.read_csv(os.path.join(tdata, "seminr-mobi-basic-outer-model.csv"), index_col=0).loc[mmvs]
    nm["R/mobi/weight"] = som["weight"].values.astype(np.float64)
    nm["R/mobi/loading"] = som["loading"].values.astype(np.float64)
    assert not [k for k, v in nm.items() if np.asarray(v).dtype == object]
    np.savez_compressed(os.path.join(HERE, "nonmetric.npz"), **nm)
    syn["cases"] = np.array([cs[0] for cs in cases])
    assert not [k for k, v in syn.items() if np.asarray(v).dtype == object]
    np.savez_compressed(os.path.join(HERE, "synthetic.npz"), **syn)
    print("wrote fixtures to", HERE)


def make_hoc():
    """Higher-order construct, two-stage approach (estimator.py:41-52): the reference's own test case
    (tests/test_regression_seminr.py:49-74) plus a Mode-B / centroid variant, with the R (seminr) values."""
    from plspm.scale import Scale
    tdata = os.path.join(REF, "tests", "data")
    mobi = pd.read_csv(os.path.join(tdata, "mobi.csv"), index_col=0)
    prefix = {"Expectation": "CUEX", "Quality": "PERQ", "Loyalty": "CUSL", "Image": "IMAG", "Complaints": "CUSCO",
              "Value": "PERV"}
    order = ["Expectation", "Quality", "Loyalty", "Image", "Complaints", "Value"]
    out = {}
    mvs_all = [m for lv in order for m in mobi.columns if m.startswith(prefix[lv])]
    out["mobi/X"] = mobi.loc[:, mvs_all].values.astype(np.float64)
    out["mobi/mvs"] = np.array(mvs_all)
    for tag, scheme, hoc_mode, tol in (("path", Scheme.PATH, Mode.A, 1e-8), ("centroid_b", Scheme.CENTROID, Mode.B, 1e-7)):
        st = c.Structure()
        st.add_path(["Expectation", "Quality"], ["Satisfaction"])
        st.add_path(["Satisfaction"], ["Complaints", "Loyalty"])
        cfg = c.Config(st.path(), default_scale=Scale.NUM)
        cfg.add_higher_order("Satisfaction", hoc_mode, ["Image", "Value"])
        for lv in order:
            cfg.add_lv_with_columns_named(lv, Mode.B if lv == "Quality" else Mode.A, mobi, prefix[lv])
        calc = Plspm(mobi, cfg, scheme, 100, tol)
        lvs = list(calc.path_coefficients().index)
        om = calc.outer_model()
        out[tag + "/lvs"] = np.array([str(v) for v in lvs])
        out[tag + "/outer_index"] = np.array([str(v) for v in om.index])
        out[tag + "/weights"] = om["weight"].values.astype(np.float64)
        out[tag + "/loadings"] = om["loading"].values.astype(np.float64)
        out[tag + "/communality"] = om["communality"].values.astype(np.float64)
        out[tag + "/path_coefficients"] = calc.path_coefficients().loc[lvs, lvs].values.astype(np.float64)
        out[tag + "/scores"] = calc.scores().loc[:, lvs].values.astype(np.float64)
        out[tag + "/r_squared"] = calc.inner_summary().loc[lvs, "r_squared"].values.astype(np.float64)
    r_om = pd.read_csv(os.path.join(tdata, "seminr-mobi-hoc-ts-outer-model.csv"), index_col=0)
    r_paths = pd.read_csv(os.path.join(tdata, "seminr-mobi-hoc-ts-paths.csv"), index_col=0).transpose()
    out["R/outer_index"] = np.array([str(v) for v in r_om.index])
    out["R/weight"] = r_om["weight"].values.astype(np.float64)
    out["R/loading"] = r_om["loading"].values.astype(np.float64)
    out["R/path_lvs"] = np.array([str(v) for v in r_paths.index])
    out["R/path_coefficients"] = r_paths.loc[list(r_paths.index), list(r_paths.index)].values.astype(np.float64)
    assert not [k for k, v in out.items() if np.asarray(v).dtype == object]
    np.savez_compressed(os.path.join(HERE, "hoc.npz"), **out)
    print("wrote hoc.npz")


def make_missing():
    """Bootstrap replicates of data WITH missing values: the reference re-imputes each resample with the column means
    of its own observed rows (bootstrap.py:57 -> estimator.py:33 -> config.py:299-305 -> util.py:61-68)."""
    sat = pd.read_csv(os.path.join(REF, "tests", "data", "satisfaction.csv"), index_col=0)
    path = satisfaction_structure()
    lvs = list(path)
    blocks = {lv: [m for m in sat.columns if m.startswith(SAT_PREFIX[lv])] for lv in lvs}
    mvs = [m for lv in lvs for m in blocks[lv]]
    X = sat.loc[:, mvs].astype(np.float64).copy()
    rng = np.random.default_rng(77)
    holes = [(int(r), int(cix)) for r, cix in zip(rng.integers(0, X.shape[0], 40), rng.integers(0, X.shape[1], 40))]
    holes += [(int(r), 3) for r in rng.integers(0, X.shape[0], 25)]  # one column with many holes
    for r, cix in holes:
        X.iat[r, cix] = np.nan
    idx = rng.integers(0, X.shape[0], (16, X.shape[0]), dtype=np.int32)
    out = {"X": X.to_numpy(), "idx": idx, "lvs": np.array(lvs), "mvs": np.array(mvs),
           "block_sizes": np.array([len(blocks[lv]) for lv in lvs], dtype=np.int32),
           "path": path.loc[lvs, lvs].to_numpy(dtype=np.int8)}
    for sname, mname, scaled in (("centroid", "A", True), ("path", "B", True), ("factorial", "A", False)):
        mode = Mode.A if mname == "A" else Mode.B
        tag = "boot/%s/%s/%s" % (sname, mname, "scaled" if scaled else "unscaled")
        r = run_reference_replicates(X, path, blocks, {lv: mode for lv in lvs}, SCHEMES[sname], scaled, idx)
        print(tag, "ok", int(r["ok"].sum()), "iters", np.bincount(r["iterations"]))
        out.update(flatten(tag, r))
    np.savez_compressed(os.path.join(HERE, "missing.npz"), **out)


def make_categorical():
    """Ordinal / nominal scales and non-metric missing data on russa (reference tests/test_regression_nonmetric.py:
    97-137): outputs of the reference + the R golden inner summaries of its own tests."""
    from plspm.scale import Scale
    tdata = os.path.join(REF, "tests", "data")
    russa = pd.read_csv(os.path.join(tdata, "russa.csv"), index_col=0)
    st = c.Structure()
    st.add_path(["AGRI", "IND"], ["POLINS"])
    rpath = st.path()
    lvs = list(rpath)
    out = {"columns": np.array(list(russa.columns)), "X": russa.to_numpy(dtype=np.float64), "lvs": np.array(lvs),
           "path": rpath.loc[lvs, lvs].to_numpy(dtype=np.int8)}
    O, Nm = Scale.ORD, Scale.NOM
    cases = {
        "categorical": (Mode.A, {"IND": [("gnpr", O), ("labo", O)], "POLINS": [("ecks", None), ("death", None), ("demo", Nm), ("inst", None)],
                                 "AGRI": [("gini", None), ("farm", None), ("rent", None)]}, "russa.categorical.inner_summary.csv", None),
        "categorical_mode_b": (Mode.B, {"AGRI": [("gini", None), ("farm", None), ("rent", None)], "IND": [("gnpr", O), ("labo", O)],
                                        "POLINS": [("ecks", None), ("death", None), ("demo", Nm), ("inst", None)]},
                               "russa.categorical.mode_b.inner_summary.csv", None),
        "missing": (Mode.A, {"AGRI": [("gini", None), ("farm", None), ("rent", None)], "IND": [("gnpr", None), ("labo", None)],
                             "POLINS": [("ecks", None), ("death", None), ("demo", None), ("inst", None)]},
                    "russa.missing.inner_summary.csv", [(0, 0), (3, 3), (5, 5)]),
    }
    for name, (mode, spec, csv, holes) in cases.items():
        data = russa.copy()
        for r, col in holes or []:
            data.iloc[r, col] = np.nan
        for sname in (("centroid", "factorial", "path") if name == "categorical" else ("centroid",)):
            config = c.Config(rpath, default_scale=Scale.NUM)
            for lv, mvs in spec.items():
                config.add_lv(lv, mode, *[c.MV(m, sc) for m, sc in mvs])
            calc = Plspm(data, config, SCHEMES[sname], 100, 0.0000001)
            mvs_all = [m for lv in lvs for m, _ in spec[lv]]
            om = calc.outer_model()
            tag = "%s/%s/" % (name, sname)
            out[tag + "mvs"] = np.array(mvs_all)
            out[tag + "weights"] = om.loc[mvs_all, "weight"].to_numpy(dtype=np.float64)
            out[tag + "loadings"] = om.loc[mvs_all, "loading"].to_numpy(dtype=np.float64)
            out[tag + "scores"] = calc.scores().loc[:, lvs].to_numpy(dtype=np.float64)
            out[tag + "path_coefficients"] = calc.path_coefficients().loc[lvs, lvs].to_numpy(dtype=np.float64)
            out[tag + "crossloadings"] = calc.crossloadings().loc[mvs_all, lvs].to_numpy(dtype=np.float64)
            isum = calc.inner_summary().loc[lvs]
            for col in ("r_squared", "block_communality", "mean_redundancy", "ave"):
                out[tag + "inner_summary/" + col] = isum[col].to_numpy(dtype=np.float64)
            out[tag + "gof"] = np.float64(calc.goodness_of_fit())
            print(tag, "weights", np.round(out[tag + "weights"], 4))
        exp = pd.read_csv(os.path.join(tdata, csv), index_col=0).loc[lvs]
        for col in ("r_squared", "block_communality", "mean_redundancy", "ave"):
            out["R/%s/%s" % (name, col)] = exp[col].to_numpy(dtype=np.float64)
        out["R/%s/type" % name] = np.array([str(v) for v in exp["type"]])
        if holes:
            out[name + "/holes"] = np.array(holes, dtype=np.int64)
    # a few bootstrap replicates of the categorical model with injected indices (bootstrap.py:54-66)
    idx = np.random.default_rng(5).integers(0, russa.shape[0], (6, russa.shape[0]), dtype=np.int32)
    out["categorical/boot/idx"] = idx
    mode, spec, _, _ = cases["categorical"]
    config = c.Config(rpath, default_scale=Scale.NUM)
    for lv, mvs in spec.items():
        config.add_lv(lv, mode, *[c.MV(m, sc) for m, sc in mvs])
    filtered = config.filter(russa)
    n = filtered.shape[0]
    calculator = ref_weights.WeightsCalculatorFactory(config, 100, 1e-7, np.sqrt(n / (n - 1)), Scheme.CENTROID)
    estimator = Estimator(config)
    mvs_all = [m for lv in lvs for m, _ in spec[lv]]
    W = np.full((len(idx), len(mvs_all)), np.nan)
    ok = np.zeros(len(idx), dtype=np.int8)
    for b in range(len(idx)):
        try:
            fd, sc, w = estimator.estimate(calculator, filtered.iloc[idx[b], :])
            W[b] = w.loc[mvs_all, "weight"].to_numpy()
            ok[b] = 1
        except Exception as e:
            print("categorical replicate", b, "failed in the reference:", repr(e))
    out["categorical/boot/weights"], out["categorical/boot/ok"] = W, ok
    assert not [k for k, v in out.items() if np.asarray(v).dtype == object]
    np.savez_compressed(os.path.join(HERE, "categorical.npz"), **out)


def make_collinear():
    """A Mode-B block with an exactly duplicated column: the reference's lstsq (mode.py:50-52, gelsd) returns the
    minimum-norm weights, the engine's Cholesky reports the block singular and the replicate is dropped.  The fixture
    pins both sides of that documented difference."""
    from plspm_b200.synth import make_synthetic
    X, pm = make_synthetic(300, 3, 3, 5)
    X = X.copy()
    X[:, 4] = X[:, 3]  # block 1: columns 3 and 4 identical
    lvs = ["L0", "L1", "L2"]
    mvs = ["x%d" % i for i in range(9)]
    blocks = {lv: mvs[3 * i:3 * i + 3] for i, lv in enumerate(lvs)}
    path = pd.DataFrame(pm, index=lvs, columns=lvs)
    df = pd.DataFrame(X, columns=mvs)
    out = {"X": X, "path": np.asarray(pm, dtype=np.int8), "block_sizes": np.array([3, 3, 3], dtype=np.int32)}
    r = run_reference(df, path, blocks, {lv: Mode.B for lv in lvs}, Scheme.CENTROID, True)
    out["ref/weights"] = r["weights"]
    out["ref/r_squared"] = r["r_squared"]
    out["ref/iterations"] = r["iterations"]
    print("collinear Mode B: reference weights", np.round(r["weights"], 4), "iterations", r["iterations"])
    np.savez_compressed(os.path.join(HERE, "collinear.npz"), **out)


if __name__ == "__main__":
    if "--only-collinear" in sys.argv:
        make_collinear()
    elif "--only-categorical" in sys.argv:
        make_categorical()
    elif "--only-hoc" in sys.argv:
        make_hoc()
    elif "--only-missing" in sys.argv:
        make_missing()
    else:
        main()
        make_hoc()
        make_missing()
        make_collinear()
        make_categorical()
