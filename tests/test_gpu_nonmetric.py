"""GPU parity of the numeric non-metric path (Scale.NUM / Scale.RAW, complete data):
the CUDA path through the C ABI vs oracle/plspm_oracle_nonmetric.py and vs outputs of the reference itself
(tests/golden/nonmetric.npz).  Tolerance 1e-6 relative; iteration counts identical."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import plspm_oracle as orc
from oracle import plspm_oracle_nonmetric as onm
from plspm_b200.synth import make_synthetic
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
REL = 1e-6
SCHEMES = ("centroid", "factorial", "path")


@pytest.fixture(scope="module")
def eng():
    from plspm_b200 import engine
    engine.load()
    assert engine.device_count() > 0, "no CUDA device"
    engine.set_device(0)
    return engine


@pytest.fixture(scope="module")
def nm():
    return np.load(os.path.join(GOLDEN, "nonmetric.npz"), allow_pickle=False)


def check(got, ref, cross=True):
    assert got["status"] == 0
    assert got["iterations"] == ref["iterations"]
    np.testing.assert_allclose(got["weights"], ref["weights"], rtol=REL)
    np.testing.assert_allclose(got["scores"], ref["scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(got["loadings"], ref["loadings"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(got["path_coefficients"], ref["path_coefficients"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(got["total_effects"], ref["total_effects"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(got["r_squared"], ref["r_squared"], rtol=REL, atol=1e-9)
    if cross:
        np.testing.assert_allclose(got["crossloadings"], ref["crossloadings"], rtol=REL, atol=1e-8)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("mode", (0, 1))
def test_russa_vs_oracle_and_reference(eng, nm, scheme, mode):
    X, bs, path = nm["russa/X"], nm["russa/block_sizes"], nm["russa/path"]
    model = eng.Model(bs, [mode] * 3, path, True, eng.TILES_FULL, numeric=True)
    data = eng.Data(model, X)
    got = eng.fit(model, data, scheme, tol=1e-7)
    check(got, onm.fit_num(X, bs, [mode] * 3, path, scheme, tol=1e-7))
    tag = "russa/%s/%s/" % (scheme, "AB"[mode])
    np.testing.assert_allclose(got["weights"], nm[tag + "weights"], rtol=REL)
    np.testing.assert_allclose(got["scores"], nm[tag + "scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(got["loadings"], nm[tag + "loadings"], rtol=REL)
    np.testing.assert_allclose(got["path_coefficients"], nm[tag + "path_coefficients"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(got["crossloadings"], nm[tag + "crossloadings"], rtol=REL, atol=1e-9)
    if mode == 0:  # R plspm values held by the reference's own tests
        np.testing.assert_allclose(got["weights"], nm["R/russa/%s/weight" % scheme], rtol=REL)
        np.testing.assert_allclose(got["loadings"], nm["R/russa/%s/loading" % scheme], rtol=REL)


def test_mobi_mixed_modes(eng, nm):
    X, bs, path, modes = nm["mobi/X"], nm["mobi/block_sizes"], nm["mobi/path"], nm["mobi/modes"]
    model = eng.Model(bs, modes, path, True, eng.TILES_FULL, numeric=True)
    data = eng.Data(model, X)
    got = eng.fit(model, data, "path", tol=1e-8)
    check(got, onm.fit_num(X, bs, modes, path, "path", tol=1e-8))
    np.testing.assert_allclose(got["weights"], nm["mobi/weights"], rtol=REL)
    np.testing.assert_allclose(got["path_coefficients"], nm["mobi/path_coefficients"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(got["weights"], nm["R/mobi/weight"], rtol=1e-5)  # reference test_regression_seminr.py:42


@pytest.mark.parametrize("N,L,K,mode,scheme", [(3000, 6, 4, 0, "centroid"), (20000, 8, 5, 1, "path"),
                                               (5000, 12, 3, 0, "factorial"), (2500, 5, 11, 0, "path")])
def test_synthetic_full_and_sparse_tiles(eng, N, L, K, mode, scheme):
    X, path = make_synthetic(N, L, K, 11 + L)
    ref = onm.fit_num(X, [K] * L, [mode] * L, path, scheme)
    full = eng.Model([K] * L, [mode] * L, path, True, eng.TILES_FULL, numeric=True)
    data = eng.Data(full, X)
    check(eng.fit(full, data, scheme), ref)
    # a data handle serves every model with the same column layout: sparse-tile twin on the same upload
    sparse = eng.Model([K] * L, [mode] * L, path, True, eng.TILES_SPARSE, numeric=True)
    got = eng.fit(sparse, data, scheme) if sparse.full_tiles else None
    if got is None:
        with pytest.raises(eng.EngineError):
            eng.fit(sparse, data, scheme)  # crossloadings need the full tile set
    # bootstrap replicates on the sparse tile set vs the oracle on the resampled rows
    rng = np.random.default_rng(5)
    idx = rng.integers(0, N, size=(5, N)).astype(np.int32)
    rows, status, iters = eng.bootstrap(sparse, data, scheme, 0, 5, idx=idx)
    w, r2, total, direct, load = sparse.split_row(rows)
    for b in range(5):
        rb = onm.fit_num(X[idx[b]], [K] * L, [mode] * L, path, scheme)
        assert status[b] == 0 and iters[b] == rb["iterations"]
        np.testing.assert_allclose(w[b], rb["weights"], rtol=REL)
        np.testing.assert_allclose(load[b], rb["loadings"], rtol=REL, atol=1e-9)
        np.testing.assert_allclose(r2[b], rb["r_squared"], rtol=REL, atol=1e-9)
        pairs = list(zip(sparse.effects_from, sparse.effects_to))
        np.testing.assert_allclose(direct[b], [rb["path_coefficients"][t, f] for f, t in pairs], rtol=REL, atol=1e-9)
        np.testing.assert_allclose(total[b], [rb["total_effects"][t, f] for f, t in pairs], rtol=REL, atol=1e-9)


def test_bootstrap_seeded_matches_engine_indices(eng):
    N, L, K = 4000, 6, 4
    X, path = make_synthetic(N, L, K, 3)
    model = eng.Model([K] * L, [0] * L, path, True, numeric=True)
    data = eng.Data(model, X)
    rows, status, iters = eng.bootstrap(model, data, "centroid", 7, 40, seed=99)
    assert (status == 0).all()
    w = model.split_row(rows)[0]
    for b in (0, 17, 39):
        idx = eng.resample_indices(99, 7 + b, N)
        rb = onm.fit_num(X[idx], [K] * L, [0] * L, path, "centroid")
        assert iters[b] == rb["iterations"]
        np.testing.assert_allclose(w[b], rb["weights"], rtol=REL)
    # replicate batches do not depend on how the range is cut
    rows2, _, _ = eng.bootstrap(model, data, "centroid", 20, 10, seed=99)
    np.testing.assert_array_equal(rows2, rows[13:23])


def test_not_converged_status(eng, nm):
    X, bs, path = nm["russa/X"], nm["russa/block_sizes"], nm["russa/path"]
    model = eng.Model(bs, [0] * 3, path, True, eng.TILES_FULL, numeric=True)
    data = eng.Data(model, X)
    got = eng.fit(model, data, "centroid", tol=1e-30, max_iter=3)
    assert got["status"] == eng.STATUS_NOT_CONVERGED and got["iterations"] == 4
    with pytest.raises(orc.NotConverged):
        onm.fit_num(X, bs, [0] * 3, path, "centroid", tol=1e-30, max_iter=3)


def test_dropin_api_scale_num(eng, nm):
    """The reference-facing call: Plspm(data, Config(..., default_scale=Scale.NUM)) (test_regression_plspm.py style)."""
    import plspm.config as c
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scale import Scale
    from plspm.scheme import Scheme
    lvs, mvs = [str(v) for v in nm["russa/lvs"]], [str(v) for v in nm["russa/mvs"]]
    frame = pd.DataFrame(nm["russa/X"], columns=mvs)
    path = pd.DataFrame(nm["russa/path"], index=lvs, columns=lvs)
    bs = nm["russa/block_sizes"]
    config = c.Config(path, default_scale=Scale.NUM)
    o = 0
    for lv, k in zip(lvs, bs):
        config.add_lv(lv, Mode.A, *[c.MV(m) for m in mvs[o:o + k]])
        o += k
    calc = Plspm(frame, config, Scheme.CENTROID, 100, 0.0000001, bootstrap=True, bootstrap_iterations=50,
                 bootstrap_seed=4)
    om = calc.outer_model()
    np.testing.assert_allclose(om.loc[mvs, "weight"], nm["russa/centroid/A/weights"], rtol=REL)
    np.testing.assert_allclose(om.loc[mvs, "loading"], nm["russa/centroid/A/loadings"], rtol=REL)
    np.testing.assert_allclose(calc.scores().loc[:, lvs].to_numpy(), nm["russa/centroid/A/scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(calc.path_coefficients().loc[lvs, lvs].to_numpy(), nm["russa/centroid/A/path_coefficients"],
                               rtol=REL, atol=1e-9)
    np.testing.assert_allclose(calc.crossloadings().loc[mvs, lvs].to_numpy(), nm["russa/centroid/A/crossloadings"],
                               rtol=REL, atol=1e-9)
    boot = calc.bootstrap()
    assert np.isfinite(boot.weights().loc[:, "mean"]).all() and len(boot.samples()["weights"]) > 40
    # mixed RAW + NUM scales are promoted to NUM (config.py:311-313) and give the same estimates
    config2 = c.Config(path, default_scale=Scale.RAW)
    o = 0
    for lv, k in zip(lvs, bs):
        config2.add_lv(lv, Mode.A, *[c.MV(m, Scale.NUM if i == 0 else None) for i, m in enumerate(mvs[o:o + k])])
        o += k
    calc2 = Plspm(frame, config2, Scheme.CENTROID, 100, 0.0000001)
    np.testing.assert_allclose(calc2.outer_model().loc[mvs, "weight"], om.loc[mvs, "weight"], rtol=1e-12)
    # ordinal / nominal scales take the reference-style host path (tests/test_host_nonmetric.py pins it to the reference)
    config3 = c.Config(path, default_scale=Scale.ORD)
    o = 0
    for lv, k in zip(lvs, bs):
        config3.add_lv(lv, Mode.A, *[c.MV(m) for m in mvs[o:o + k]])
        o += k
    calc3 = Plspm(frame, config3, Scheme.CENTROID)
    assert calc3.iterations() > 0 and np.isfinite(calc3.outer_model().loc[mvs, "weight"]).all()


# ---- higher-order constructs, two-stage approach -----------------------------------
MOBI_PREFIX = {"Expectation": "CUEX", "Quality": "PERQ", "Loyalty": "CUSL", "Image": "IMAG", "Complaints": "CUSCO",
               "Value": "PERV"}


@pytest.fixture(scope="module")
def hz():
    return np.load(os.path.join(GOLDEN, "hoc.npz"), allow_pickle=False)


def mobi_hoc_config(frame, hoc_mode):
    import plspm.config as c
    from plspm.mode import Mode
    from plspm.scale import Scale
    structure = c.Structure()
    structure.add_path(["Expectation", "Quality"], ["Satisfaction"])
    structure.add_path(["Satisfaction"], ["Complaints", "Loyalty"])
    config = c.Config(structure.path(), default_scale=Scale.NUM)
    config.add_higher_order("Satisfaction", hoc_mode, ["Image", "Value"])
    for lv in ("Expectation", "Quality", "Loyalty", "Image", "Complaints", "Value"):
        config.add_lv_with_columns_named(lv, Mode.B if lv == "Quality" else Mode.A, frame, MOBI_PREFIX[lv])
    return config


@pytest.mark.parametrize("tag", ("path", "centroid_b"))
def test_hoc_two_stage_dropin(eng, hz, tag):
    """The reference's own higher-order test (tests/test_regression_seminr.py:49-74) through the drop-in API."""
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scheme import Scheme
    frame = pd.DataFrame(hz["mobi/X"], columns=[str(v) for v in hz["mobi/mvs"]])
    scheme, hoc_mode, tol = (Scheme.PATH, Mode.A, 1e-8) if tag == "path" else (Scheme.CENTROID, Mode.B, 1e-7)
    calc = Plspm(frame, mobi_hoc_config(frame, hoc_mode), scheme, 100, tol)
    lvs = [str(v) for v in hz[tag + "/lvs"]]
    index = [str(v) for v in hz[tag + "/outer_index"]]
    om = calc.outer_model()
    assert set(om.index) == set(index)
    np.testing.assert_allclose(om.loc[index, "weight"], hz[tag + "/weights"], rtol=REL)
    np.testing.assert_allclose(om.loc[index, "loading"], hz[tag + "/loadings"], rtol=REL)
    np.testing.assert_allclose(om.loc[index, "communality"], hz[tag + "/communality"], rtol=REL)
    np.testing.assert_allclose(calc.path_coefficients().loc[lvs, lvs].to_numpy(), hz[tag + "/path_coefficients"],
                               rtol=REL, atol=1e-9)
    np.testing.assert_allclose(calc.scores().loc[:, lvs].to_numpy(), hz[tag + "/scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(calc.inner_summary().loc[lvs, "r_squared"], hz[tag + "/r_squared"], rtol=REL, atol=1e-9)
    if tag == "path":  # seminr values, tolerances of the reference's test
        r_index = [str(v) for v in hz["R/outer_index"]]
        common = [m for m in r_index if m in om.index]
        np.testing.assert_allclose(om.loc[common, "weight"], [hz["R/weight"][r_index.index(m)] for m in common], rtol=1e-4)
        np.testing.assert_allclose(om.loc[common, "loading"], [hz["R/loading"][r_index.index(m)] for m in common], rtol=1e-4)
        r_lvs = [str(v) for v in hz["R/path_lvs"]]
        np.testing.assert_allclose(calc.path_coefficients().loc[r_lvs, r_lvs].to_numpy(), hz["R/path_coefficients"],
                                   rtol=1e-6, atol=1e-9)
    uni = calc.unidimensionality()
    assert np.isfinite(uni.loc["Expectation", "eig_1st"]) and np.isnan(uni.loc["Satisfaction", "eig_1st"])


def test_hoc_bootstrap_replicates_vs_oracle(eng, hz):
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scheme import Scheme
    mvs = [str(v) for v in hz["mobi/mvs"]]
    X = hz["mobi/X"]
    frame = pd.DataFrame(X, columns=mvs)
    N = X.shape[0]
    idx = np.random.default_rng(8).integers(0, N, size=(10, N)).astype(np.int32)
    calc = Plspm(frame, mobi_hoc_config(frame, Mode.A), Scheme.PATH, 100, 1e-8, bootstrap=True, bootstrap_iterations=10,
                 processes=1, bootstrap_indices=idx)
    boot = calc.bootstrap()
    status, _ = boot.replicate_status()
    assert (status == 0).all()
    lvs = list(calc.path_coefficients().index)
    path = calc.path_coefficients().loc[lvs, lvs].to_numpy() != 0
    modes = {lv: 0 for lv in MOBI_PREFIX}
    modes["Quality"] = 1
    modes["Satisfaction"] = 0
    samples = boot.samples()
    for b in (0, 4, 9):
        blocks = {lv: X[idx[b]][:, [i for i, m in enumerate(mvs) if m.startswith(p)]] for lv, p in MOBI_PREFIX.items()}
        _, s2, names = onm.fit_num_hoc(blocks, lvs, path.astype(np.int8), modes, {"Satisfaction": ["Image", "Value"]},
                                       "path", 1e-8)
        got = samples["weights"].iloc[b]
        np.testing.assert_allclose([got["Image"], got["Value"]], [s2["weights"][names.index("Image")],
                                                                  s2["weights"][names.index("Value")]], rtol=REL)
        np.testing.assert_allclose(samples["r_squared"].iloc[b].loc[lvs], s2["r_squared"], rtol=REL, atol=1e-9)
        np.testing.assert_allclose(samples["paths"].iloc[b].loc["Satisfaction -> Loyalty"],
                                   s2["path_coefficients"][lvs.index("Loyalty"), lvs.index("Satisfaction")], rtol=REL)
    assert np.isfinite(boot.paths().to_numpy()).all()
