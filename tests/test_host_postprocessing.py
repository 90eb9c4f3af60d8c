"""Host post-processing of engine outputs (InnerModel table, OuterModel, InnerSummary, effects,
bootstrap summaries, unidimensionality) checked against the reference outputs / R golden values,
with the ORACLE standing in for the engine (tests only).  CPU only."""
import numpy as np
import pandas as pd
import pytest

import plspm.config as c
from oracle import plspm_oracle as orc
from plspm.bootstrap import _create_summary
from plspm.inner_model import InnerModel, effects_table
from plspm.inner_summary import InnerSummary
from plspm.mode import Mode
from plspm.outer_model import OuterModel
from plspm.unidimensionality import Unidimensionality


def sat_config(sat, mode=Mode.A, scaled=False):
    lvs = [str(v) for v in sat["lvs"]]
    mvs = [str(v) for v in sat["mvs"]]
    path = pd.DataFrame(sat["path"].astype(int), index=lvs, columns=lvs)
    cfg = c.Config(path, scaled=scaled)
    off = 0
    blocks = {}
    for lv, k in zip(lvs, sat["block_sizes"]):
        blocks[lv] = mvs[off:off + k]
        cfg.add_lv(lv, mode, *[c.MV(m) for m in blocks[lv]])
        off += k
    return cfg, path, lvs, mvs, blocks


def test_inner_model_table_vs_reference_and_r(sat):
    cfg, path, lvs, mvs, blocks = sat_config(sat)
    r = orc.fit(sat["X"], sat["block_sizes"], [0] * 6, sat["path"], "centroid", False)
    scores = pd.DataFrame(r["scores"], columns=lvs)
    for kwargs in ({}, dict(path_coefficients=r["path_coefficients"], r_squared=r["r_squared"])):
        im = InnerModel(path, scores, **kwargs)
        tab = im.inner_model()
        ref_idx = [str(v) for v in sat["ref/centroid/A/unscaled/inner_model_index"]]
        np.testing.assert_allclose(tab.loc[ref_idx, ["estimate", "std error", "t", "p>|t|"]].values,
                                   sat["ref/centroid/A/unscaled/inner_model"], rtol=1e-7, atol=1e-12)
        sat_rows = tab[tab["to"] == "SAT"].set_index("from").loc[[str(v) for v in sat["R/inner_model_SAT_from"]]]
        np.testing.assert_allclose(sat_rows[["estimate", "std error", "t", "p>|t|"]].values, sat["R/inner_model_SAT"],
                                   rtol=1e-7)
        np.testing.assert_allclose(im.path_coefficients().values, r["path_coefficients"], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(im.r_squared().values, sat["R/inner_summary/r_squared"], rtol=1e-8, atol=1e-13)
        assert im.endogenous() == ["EXPE", "QUAL", "VAL", "SAT", "LOY"]
        eff = im.effects()
        assert list(eff["from"]) == [str(v) for v in sat["R/effects_from"]]
        assert list(eff["to"]) == [str(v) for v in sat["R/effects_to"]]
        for col in ("direct", "indirect", "total"):
            np.testing.assert_allclose(eff[col].values, sat["R/effects_" + col], rtol=1e-7, atol=1e-12)


def test_effects_table_two_lvs():
    pc = pd.DataFrame([[0.0, 0.0], [0.4, 0.0]], index=["a", "b"], columns=["a", "b"])
    eff = effects_table(pc)
    assert list(eff.index) == ["a -> b"] and eff.loc["a -> b", "total"] == 0.4 and eff.loc["a -> b", "indirect"] == 0


def test_outer_model_and_inner_summary_vs_r(sat):
    cfg, path, lvs, mvs, blocks = sat_config(sat)
    r = orc.fit(sat["X"], sat["block_sizes"], [0] * 6, sat["path"], "centroid", False)
    scores = pd.DataFrame(r["scores"], columns=lvs)
    im = InnerModel(path, scores, r["path_coefficients"], r["r_squared"])
    om = OuterModel(pd.DataFrame(r["weights"], index=mvs, columns=["weight"]), r["loadings"], r["crossloadings"], lvs,
                    blocks, im.r_squared())
    tab = om.model()
    assert list(tab.columns) == ["weight", "loading", "communality", "redundancy"]
    for col in ("weight", "loading", "communality", "redundancy"):
        np.testing.assert_allclose(tab[col].values, sat["R/centroid/" + col], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(om.crossloadings().values, sat["R/crossloadings"], rtol=1e-7, atol=1e-12)
    isum = InnerSummary(cfg, im.r_squared(), im.r_squared_adj(), tab)
    s = isum.summary()
    for col in ("r_squared", "block_communality", "mean_redundancy", "ave"):
        np.testing.assert_allclose(s[col].values, sat["R/inner_summary/" + col], rtol=1e-8, atol=1e-13)
    assert list(s["type"]) == ["Exogenous"] + ["Endogenous"] * 5
    assert abs(isum.goodness_of_fit() - float(sat["R/gof"])) < 1e-9


def test_mode_b_inner_summary_vs_r(sat):
    cfg, path, lvs, mvs, blocks = sat_config(sat, Mode.B)
    r = orc.fit(sat["X"], sat["block_sizes"], [1] * 6, sat["path"], "centroid", False)
    im = InnerModel(path, pd.DataFrame(r["scores"], columns=lvs), r["path_coefficients"], r["r_squared"])
    om = OuterModel(pd.DataFrame(r["weights"], index=mvs, columns=["weight"]), r["loadings"], r["crossloadings"], lvs,
                    blocks, im.r_squared())
    s = InnerSummary(cfg, im.r_squared(), im.r_squared_adj(), om.model()).summary()
    for col in ("r_squared", "block_communality", "mean_redundancy"):
        np.testing.assert_allclose(s[col].values, sat["R/modeb/inner_summary/" + col], rtol=1e-7, atol=1e-12)
    assert s["ave"].isna().all()


def test_single_item_constructs_have_no_gof(sat):
    lvs = [str(v) for v in sat["lvs"]]
    path = pd.DataFrame(sat["path"].astype(int), index=lvs, columns=lvs)
    cfg = c.Config(path)
    for lv in lvs:
        cfg.add_lv(lv, Mode.A, c.MV(lv.lower() + "1"))
    om = pd.DataFrame({"communality": 1.0, "redundancy": 0.5}, index=[lv.lower() + "1" for lv in lvs])
    isum = InnerSummary(cfg, pd.Series(0.5, index=lvs), pd.Series(0.5, index=lvs), om)
    with pytest.raises(ValueError):
        isum.goodness_of_fit()


def test_create_summary_columns_and_values():
    rng = np.random.default_rng(0)
    data = pd.DataFrame(rng.standard_normal((200, 3)), columns=["a", "b", "c"])
    orig = pd.Series([1.0, 2.0, 3.0], index=["a", "b", "c"])
    s = _create_summary(data, orig)
    assert list(s.columns) == ["original", "mean", "std.error", "perc.025", "perc.975", "t stat."]
    np.testing.assert_allclose(s.values, orc.summary(data.values, orig.values), rtol=1e-12)


def test_unidimensionality_vs_r(sat):
    if "R/unidim/eig_1st" not in sat.files:
        pytest.skip("fixture without unidimensionality values")
    cfg, path, lvs, mvs, blocks = sat_config(sat)
    df = pd.DataFrame(sat["X"], columns=mvs)
    u = Unidimensionality(cfg, df, np.sqrt(250 / 249)).summary()
    for col in ("mvs", "cronbach_alpha", "dillon_goldstein_rho", "eig_1st", "eig_2nd"):
        np.testing.assert_allclose(u[col].values.astype(float), sat["R/unidim/" + col], rtol=1e-7)
    assert list(u["mode"]) == ["A"] * 6


def test_util_helpers_match_the_reference_semantics():
    """reference tests/test_util.py (impute, rank) + dummy / groupby_mean / treat_numpy restatements"""
    import plspm.util as util
    frame = pd.DataFrame({"a": [1, 2, np.nan, 3, np.nan], "b": [1, 2, 3, 4, 5], "c": [1, np.nan, 3, 0, 4]})
    expected = pd.DataFrame({"a": [1, 2, 2, 3, 2], "b": [1, 2, 3, 4, 5], "c": [1, 2, 3, 0, 4]})
    np.testing.assert_array_equal(expected, util.impute(frame))
    data = pd.Series([0.75, -1.5, 3, -1.5, 15])
    assert util.rank(data).astype(int).equals(pd.Series([2, 1, 3, 1, 4]))
    d = util.dummy(util.rank(data))
    assert list(d.columns) == [1, 2, 3, 4]
    np.testing.assert_array_equal(d.to_numpy(), [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [1, 0, 0, 0], [0, 0, 0, 1]])
    g = util.groupby_mean(np.array([[2.0, 1.0, 2.0, 3.0, 1.0], [10.0, 1.0, 20.0, 5.0, 3.0]]))
    np.testing.assert_allclose(g, [[1.0, 2.0, 3.0], [2.0, 15.0, 5.0]])
    y = np.array([1.0, 2.0, np.nan, 4.0])
    t = util.treat_numpy(y)
    np.testing.assert_allclose(np.nanmean(t), 0.0, atol=1e-15)
    np.testing.assert_allclose(np.nanstd(t, ddof=1), 1.0)


def test_unidimensionality_with_no_more_observations_than_manifest_variables():
    """N <= block size: the reference runs its PCA on the transposed block (unidimensionality.py:46); mirrored here and
    checked against the reference's literal computation with scikit-learn."""
    sk = pytest.importorskip("sklearn.decomposition")
    import plspm.config as c
    from plspm.mode import Mode
    rng = np.random.default_rng(3)
    n, k = 5, 7
    cols = ["a%d" % i for i in range(k)] + ["b0", "b1"]
    df = pd.DataFrame(np.column_stack([rng.normal(size=(n, k)) + rng.normal(size=(n, 1)), rng.normal(size=(n, 2))]), columns=cols)
    s = c.Structure()
    s.add_path(["A"], ["B"])
    cfg = c.Config(s.path(), scaled=False)
    cfg.add_lv("A", Mode.A, *[c.MV(m) for m in cols[:k]])
    cfg.add_lv("B", Mode.A, c.MV("b0"), c.MV("b1"))
    corr_n = np.sqrt(n / (n - 1))
    ours = Unidimensionality(cfg, df, corr_n).summary().loc["A"]
    blk = df[cols[:k]]
    inp = ((blk - blk.mean()) / blk.std() * corr_n).transpose()
    scores = sk.PCA().fit_transform(inp)
    sd = np.std(scores, axis=0)
    ca = max(0, (2 * np.tril(inp.corr(), -1).sum() / (inp.sum(axis=1).var() / corr_n ** 2)) * (k / (k - 1)))
    cr = np.corrcoef(np.column_stack((inp.values, scores[:, 0])), rowvar=False)[:, -1][:-1]
    rho = sum(cr) ** 2 / (sum(cr) ** 2 + (k - np.sum(np.power(cr, 2))))
    np.testing.assert_allclose([ours["eig_1st"], ours["eig_2nd"], ours["cronbach_alpha"], ours["dillon_goldstein_rho"]],
                               [sd[0] ** 2, sd[1] ** 2, ca, rho], rtol=1e-9, atol=1e-12)
