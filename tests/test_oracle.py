"""Pins the CPU oracle (oracle/plspm_oracle.py) against
 (1) the R-generated golden vectors used by the reference's own tests
     (/root/reference/tests/data/satisfaction.*.csv, test_regression_metric.py:43-94), and
 (2) outputs of the reference itself (tests/golden/make_golden.py).
CPU only."""
import os

import numpy as np
import pytest

from oracle import plspm_oracle as orc
from plspm_b200.synth import make_synthetic

SCHEMES = ("centroid", "factorial", "path")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def sat_fit(sat, scheme, mode, scaled):
    L = len(sat["block_sizes"])
    return orc.fit(sat["X"], sat["block_sizes"], [mode] * L, sat["path"], scheme, scaled)


def test_r_golden_scores_weights_paths(sat):
    r = sat_fit(sat, "centroid", orc.MODE_A, False)
    assert r["iterations"] == 4
    np.testing.assert_allclose(r["scores"], sat["R/scores"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(r["weights"], sat["R/centroid/weight"], rtol=1e-9)
    np.testing.assert_allclose(r["loadings"], sat["R/centroid/loading"], rtol=1e-9)
    np.testing.assert_allclose(r["crossloadings"], sat["R/crossloadings"], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(r["r_squared"], sat["R/inner_summary/r_squared"], rtol=1e-9, atol=1e-14)
    lvs = list(sat["lvs"])
    for f, t, d, ind, tot in zip(sat["R/effects_from"], sat["R/effects_to"], sat["R/effects_direct"],
                                 sat["R/effects_indirect"], sat["R/effects_total"]):
        i, j = lvs.index(t), lvs.index(f)
        np.testing.assert_allclose(r["path_coefficients"][i, j], d, rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(r["indirect_effects"][i, j], ind, rtol=1e-8, atol=1e-13)
        np.testing.assert_allclose(r["total_effects"][i, j], tot, rtol=1e-8, atol=1e-13)


@pytest.mark.parametrize("scheme", ("path", "factorial"))
def test_r_golden_other_schemes(sat, scheme):
    r = sat_fit(sat, scheme, orc.MODE_A, False)
    np.testing.assert_allclose(r["weights"], sat["R/%s/weight" % scheme], rtol=1e-9)
    np.testing.assert_allclose(r["loadings"], sat["R/%s/loading" % scheme], rtol=1e-9)


def test_r_golden_mode_b_rsquared(sat):
    r = sat_fit(sat, "centroid", orc.MODE_B, False)
    np.testing.assert_allclose(r["r_squared"], sat["R/modeb/inner_summary/r_squared"], rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("mode", ("A", "B"))
@pytest.mark.parametrize("scaled", (False, True))
def test_vs_reference_satisfaction(sat, scheme, mode, scaled):
    tag = "ref/%s/%s/%s/" % (scheme, mode, "scaled" if scaled else "unscaled")
    r = sat_fit(sat, scheme, orc.MODE_A if mode == "A" else orc.MODE_B, scaled)
    assert r["iterations"] == int(sat[tag + "iterations"])
    tol = 1e-9 if mode == "A" else 1e-8
    assert rel(r["weights"], sat[tag + "weights"]) < tol
    np.testing.assert_allclose(r["scores"], sat[tag + "scores"], rtol=tol, atol=1e-11)
    np.testing.assert_allclose(r["path_coefficients"], sat[tag + "path_coefficients"], rtol=tol, atol=1e-12)
    np.testing.assert_allclose(r["r_squared"], sat[tag + "r_squared"], rtol=tol, atol=1e-12)
    np.testing.assert_allclose(r["loadings"], sat[tag + "loadings"], rtol=tol, atol=1e-12)
    np.testing.assert_allclose(r["crossloadings"], sat[tag + "crossloadings"], rtol=tol, atol=1e-12)


def test_vs_reference_mixed_modes(sat):
    r = orc.fit(sat["X"], sat["block_sizes"], sat["mixed_modes"], sat["path"], "path", True)
    tag = "ref/path/mixed/scaled/"
    assert r["iterations"] == int(sat[tag + "iterations"])
    assert rel(r["weights"], sat[tag + "weights"]) < 1e-8
    np.testing.assert_allclose(r["path_coefficients"], sat[tag + "path_coefficients"], rtol=1e-8, atol=1e-12)


def test_vs_reference_synthetic(syn):
    for name in syn["cases"]:
        N, L, K, seed = (int(v) for v in syn[name + "/gen"])
        X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[name + "/reverse"]))
        mode = orc.MODE_A if str(syn[name + "/mode"]) == "A" else orc.MODE_B
        r = orc.fit(X, [K] * L, [mode] * L, path, str(syn[name + "/scheme"]), bool(syn[name + "/scaled"]))
        assert r["iterations"] == int(syn[name + "/iterations"]), name
        assert rel(r["weights"], syn[name + "/weights"]) < 1e-8, name
        np.testing.assert_allclose(r["scores"], syn[name + "/scores"], rtol=1e-8, atol=1e-10, err_msg=name)
        np.testing.assert_allclose(r["path_coefficients"], syn[name + "/path_coefficients"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(r["total_effects"][np.tril_indices(L, -1)],
                                   _total_from_effects(syn, name, L)[np.tril_indices(L, -1)], rtol=1e-8, atol=1e-11)


def _total_from_effects(fx, name, L):
    lvs = list(fx[name + "/lvs"])
    T = np.zeros((L, L))
    for f, t, v in zip(fx[name + "/effects_from"], fx[name + "/effects_to"], fx[name + "/effects_total"]):
        T[lvs.index(t), lvs.index(f)] = v
    return T


def test_sign_flip_case_present(syn):
    # syn_b has reverse-coded blocks: scores flipped, weights keep their sign (quirk Q6)
    N, L, K, seed = (int(v) for v in syn["syn_b/gen"])
    X, path = make_synthetic(N, L, K, seed, reverse_blocks=(1, 3))
    r = orc.fit(X, [K] * L, [0] * L, path, "centroid", False)
    assert (r["signs"] < 0).any()
    np.testing.assert_allclose(r["weights"], syn["syn_b/weights"], rtol=1e-8)
    np.testing.assert_allclose(r["loadings"], syn["syn_b/loadings"], rtol=1e-8)


@pytest.mark.parametrize("case", ("centroid/A/unscaled", "path/B/scaled", "factorial/A/scaled"))
def test_bootstrap_replicates_vs_reference(sat, case):
    scheme, mode, sc = case.split("/")
    L = len(sat["block_sizes"])
    nrep = int(sat["boot/n_replicates"])
    idx = np.random.default_rng(1234).integers(0, 250, (1000, 250), dtype=np.int32)[:nrep]
    out, iters, status = orc.bootstrap(sat["X"], idx, sat["block_sizes"],
                                       [orc.MODE_A if mode == "A" else orc.MODE_B] * L, sat["path"], scheme,
                                       sc == "scaled")
    tag = "boot/%s/" % case
    assert (status == 0).all() and sat[tag + "ok"].all()
    np.testing.assert_array_equal(iters, sat[tag + "iterations"])
    P = int(sat["block_sizes"].sum())
    pairs = orc.effect_pairs(sat["path"])
    tol = 1e-8 if mode == "A" else 1e-7
    np.testing.assert_allclose(out[:, :P], sat[tag + "weights"], rtol=tol)
    np.testing.assert_allclose(out[:, P:P + L], sat[tag + "r_squared"], rtol=tol, atol=1e-12)
    tot = np.array([[sat[tag + "total_effects"][b, t, f] for f, t in pairs] for b in range(nrep)])
    direct = np.array([[sat[tag + "path_coefficients"][b, t, f] for f, t in pairs] for b in range(nrep)])
    ne = len(pairs)
    np.testing.assert_allclose(out[:, P + L:P + L + ne], tot, rtol=tol, atol=1e-12)
    np.testing.assert_allclose(out[:, P + L + ne:P + L + 2 * ne], direct, rtol=tol, atol=1e-12)
    np.testing.assert_allclose(out[:, P + L + 2 * ne:], sat[tag + "loadings"], rtol=tol, atol=1e-12)


@pytest.mark.parametrize("case", ("centroid/A/scaled", "path/B/scaled", "factorial/A/unscaled"))
def test_bootstrap_with_missing_values_vs_reference(case):
    """Data with missing values: every replicate is re-imputed with the column means of its own observed rows
    (reference bootstrap.py:57 -> estimator.py:33 -> config.py:299-305 -> util.py:61-68); fixture generated by
    running the reference (tests/golden/make_golden.py --only-missing)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "missing.npz"))
    scheme, mode, sc = case.split("/")
    L = len(g["block_sizes"])
    assert np.isnan(g["X"]).sum() > 40
    out, iters, status = orc.bootstrap(g["X"], g["idx"], g["block_sizes"], [orc.MODE_A if mode == "A" else orc.MODE_B] * L,
                                       g["path"], scheme, sc == "scaled")
    tag = "boot/%s/" % case
    assert (status == 0).all() and g[tag + "ok"].all()
    np.testing.assert_array_equal(iters, g[tag + "iterations"])
    P = int(g["block_sizes"].sum())
    tol = 1e-8 if mode == "A" else 1e-7
    np.testing.assert_allclose(out[:, :P], g[tag + "weights"], rtol=tol)
    np.testing.assert_allclose(out[:, P:P + L], g[tag + "r_squared"], rtol=tol, atol=1e-12)
    pairs = orc.effect_pairs(g["path"])
    ne = len(pairs)
    direct = np.array([[g[tag + "path_coefficients"][b, t, f] for f, t in pairs] for b in range(len(g["idx"]))])
    np.testing.assert_allclose(out[:, P + L + ne:P + L + 2 * ne], direct, rtol=tol, atol=1e-12)
    np.testing.assert_allclose(out[:, P + L + 2 * ne:], g[tag + "loadings"], rtol=tol, atol=1e-12)


def test_collinear_mode_b_block_oracle_follows_the_reference_min_norm():
    """An exactly duplicated column in a Mode-B block: the reference's lstsq (mode.py:50-52, gelsd) returns the
    minimum-norm weights (equal weights on the two copies; fixture tests/golden/collinear.npz generated by the
    reference) and so does the oracle.  The CUDA solver reaches the same answer in the moment domain (Cholesky, then
    conjugate gradients on the rank-deficient normal equations): tests/test_solver_emul.py and
    tests/test_gpu_dropin.py::test_collinear_mode_b_block_gives_the_reference_min_norm_weights."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collinear.npz"))
    w = g["ref/weights"]
    assert np.isfinite(w).all() and abs(w[3] - w[4]) < 1e-12  # the reference's min-norm answer
    r = orc.fit(g["X"], g["block_sizes"], [orc.MODE_B] * 3, g["path"], "centroid", True)
    assert r["status"] == 0 and r["iterations"] == int(g["ref/iterations"])
    np.testing.assert_allclose(r["weights"], w, rtol=1e-7)
    np.testing.assert_allclose(r["r_squared"], g["ref/r_squared"], rtol=1e-7, atol=1e-12)


def test_effect_pairs_match_reference_effects_index(sat):
    lvs = list(sat["lvs"])
    pairs = orc.effect_pairs(sat["path"])
    ref = list(zip(sat["ref/centroid/A/unscaled/effects_from"], sat["ref/centroid/A/unscaled/effects_to"]))
    assert [(lvs[f], lvs[t]) for f, t in pairs] == [(str(f), str(t)) for f, t in ref]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    z = np.zeros(1, dtype=np.uint64)
    out = orc.philox4x32(z, z, z, z, 0, 0)
    assert [int(v[0]) for v in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = z + np.uint64(0xFFFFFFFF)
    out = orc.philox4x32(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v[0]) for v in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = z + np.uint64(0x243F6A88)
    out = orc.philox4x32(p, z + np.uint64(0x85A308D3), z + np.uint64(0x13198A2E), z + np.uint64(0x03707344),
                         0xA4093822, 0x299F31D0)
    assert [int(v[0]) for v in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_indices_range_and_determinism():
    a = orc.philox_indices(7, 3, 1001)
    b = orc.philox_indices(7, 3, 1001)
    c = orc.philox_indices(7, 4, 1001)
    assert a.shape == (1001,) and a.min() >= 0 and a.max() < 1001
    assert (a == b).all() and (a != c).any()
    cnt = np.bincount(a, minlength=1001)
    assert 0.30 < (cnt == 0).mean() < 0.43  # ~ e^-1 of the rows are left out of a resample


def test_missing_values_are_mean_imputed():
    X, path = make_synthetic(120, 3, 3, seed=1)
    Xm = X.copy()
    Xm[5, 2] = np.nan
    Xi = X.copy()
    Xi[5, 2] = np.delete(X[:, 2], 5).mean()
    a = orc.fit(Xm, [3, 3, 3], [0, 0, 0], path, "centroid", True)
    b = orc.fit(Xi, [3, 3, 3], [0, 0, 0], path, "centroid", True)
    np.testing.assert_allclose(a["weights"], b["weights"], rtol=1e-12)


def test_not_converged_raises(sat):
    with pytest.raises(orc.NotConverged):
        orc.fit(sat["X"], sat["block_sizes"], [1] * 6, sat["path"], "centroid", True, tol=1e-30, max_iter=3)


# ---- nonmetric path with numeric scales (groundwork for SURVEY §8(f) row f3) ---------------------
@pytest.fixture(scope="module")
def nm():
    import os
    from tests.conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "nonmetric.npz"), allow_pickle=False)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("mode", ("A", "B"))
def test_nonmetric_num_russa_vs_reference(nm, scheme, mode):
    from oracle import plspm_oracle_nonmetric as onm
    L = len(nm["russa/block_sizes"])
    r = onm.fit_num(nm["russa/X"], nm["russa/block_sizes"], [0 if mode == "A" else 1] * L, nm["russa/path"], scheme,
                    tol=1e-7)
    tag = "russa/%s/%s/" % (scheme, mode)
    np.testing.assert_allclose(r["weights"], nm[tag + "weights"], rtol=1e-8)
    np.testing.assert_allclose(r["scores"], nm[tag + "scores"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(r["loadings"], nm[tag + "loadings"], rtol=1e-8)
    np.testing.assert_allclose(r["path_coefficients"], nm[tag + "path_coefficients"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(r["crossloadings"], nm[tag + "crossloadings"], rtol=1e-8, atol=1e-12)


def test_nonmetric_num_r_golden(nm):
    from oracle import plspm_oracle_nonmetric as onm
    for scheme in SCHEMES:
        r = onm.fit_num(nm["russa/X"], nm["russa/block_sizes"], [0] * 3, nm["russa/path"], scheme, tol=1e-7)
        np.testing.assert_allclose(r["weights"], nm["R/russa/%s/weight" % scheme], rtol=1e-7)
        np.testing.assert_allclose(r["loadings"], nm["R/russa/%s/loading" % scheme], rtol=1e-7)
    r = onm.fit_num(nm["russa/X"], nm["russa/block_sizes"], [0] * 3, nm["russa/path"], "centroid", tol=1e-7)
    np.testing.assert_allclose(r["scores"], nm["R/russa/scores"], rtol=1e-7, atol=1e-10)
    m = onm.fit_num(nm["mobi/X"], nm["mobi/block_sizes"], nm["mobi/modes"], nm["mobi/path"], "path", tol=1e-8)
    np.testing.assert_allclose(m["weights"], nm["mobi/weights"], rtol=1e-8)
    np.testing.assert_allclose(m["path_coefficients"], nm["mobi/path_coefficients"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(m["weights"], nm["R/mobi/weight"], rtol=1e-5)   # reference test_regression_seminr.py:42
    np.testing.assert_allclose(m["loadings"], nm["R/mobi/loading"], rtol=1e-5)


# ---- higher-order constructs, two-stage approach -----------------------------------
MOBI_PREFIX = {"Expectation": "CUEX", "Quality": "PERQ", "Loyalty": "CUSL", "Image": "IMAG", "Complaints": "CUSCO",
               "Value": "PERV"}


def mobi_hoc_case(hz, tag):
    mvs = [str(v) for v in hz["mobi/mvs"]]
    X = hz["mobi/X"]
    blocks = {lv: X[:, [i for i, m in enumerate(mvs) if m.startswith(p)]] for lv, p in MOBI_PREFIX.items()}
    lvs = [str(v) for v in hz[tag + "/lvs"]]
    edges = [("Expectation", "Satisfaction"), ("Quality", "Satisfaction"), ("Satisfaction", "Complaints"),
             ("Satisfaction", "Loyalty")]
    path = np.zeros((5, 5), dtype=np.int8)
    for f, t in edges:
        path[lvs.index(t), lvs.index(f)] = 1
    return blocks, lvs, path


@pytest.fixture(scope="module")
def hz():
    import os
    from tests.conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "hoc.npz"), allow_pickle=False)


@pytest.mark.parametrize("tag,scheme,hoc_mode,tol", [("path", "path", 0, 1e-8), ("centroid_b", "centroid", 1, 1e-7)])
def test_hoc_two_stage_vs_reference(hz, tag, scheme, hoc_mode, tol):
    from oracle import plspm_oracle_nonmetric as onm
    blocks, lvs, path = mobi_hoc_case(hz, tag)
    modes = {lv: 0 for lv in MOBI_PREFIX}
    modes["Quality"] = 1
    modes["Satisfaction"] = hoc_mode
    _, s2, names = onm.fit_num_hoc(blocks, lvs, path, modes, {"Satisfaction": ["Image", "Value"]}, scheme, tol)
    np.testing.assert_allclose(s2["path_coefficients"], hz[tag + "/path_coefficients"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(s2["scores"], hz[tag + "/scores"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(s2["r_squared"], hz[tag + "/r_squared"], rtol=1e-8, atol=1e-12)
    ref_index = [str(v) for v in hz[tag + "/outer_index"]]
    # stage-2 manifest names: constituents by name, first-order MVs in block order (sorted like the reference index)
    mvs = [str(v) for v in hz["mobi/mvs"]]
    label = {}
    for lv, p in MOBI_PREFIX.items():
        for j, m in enumerate([m for m in mvs if m.startswith(p)]):
            label["%s.%d" % (lv, j)] = m
    got_w = {label.get(n, n): w for n, w in zip(names, s2["weights"])}
    got_l = {label.get(n, n): w for n, w in zip(names, s2["loadings"])}
    np.testing.assert_allclose([got_w[m] for m in ref_index], hz[tag + "/weights"], rtol=1e-8)
    np.testing.assert_allclose([got_l[m] for m in ref_index], hz[tag + "/loadings"], rtol=1e-8)
    if tag == "path":  # seminr values of the reference's own test (test_regression_seminr.py:65-74)
        r_index = [str(v) for v in hz["R/outer_index"]]
        common = [m for m in r_index if m in got_w]
        assert {"Image", "Value"} <= set(common)
        np.testing.assert_allclose([got_w[m] for m in common], [hz["R/weight"][r_index.index(m)] for m in common], rtol=1e-4)
        np.testing.assert_allclose([got_l[m] for m in common], [hz["R/loading"][r_index.index(m)] for m in common], rtol=1e-4)
        r_lvs = [str(v) for v in hz["R/path_lvs"]]
        perm = [r_lvs.index(lv) for lv in lvs]
        np.testing.assert_allclose(s2["path_coefficients"], hz["R/path_coefficients"][np.ix_(perm, perm)], rtol=1e-6, atol=1e-9)
