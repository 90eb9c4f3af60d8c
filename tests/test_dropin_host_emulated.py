"""The drop-in host package (plspm-python_b200/plspm) end to end on the CPU.

The CUDA engine is replaced -- in this test module only -- by the TEST-ONLY host emulation of the solver
(tests/emul: csrc/solver_core.h / solver_num.h compiled for the host), so everything above the C ABI runs
without a GPU: Plspm(), the estimator (incl. the two-stage approach for higher-order constructs), the
outer / inner model tables, the inner summary and the bootstrap summaries.  Golden values are the outputs of
the reference itself and the R values its own tests hold (tests/golden/*.npz).  The same checks run against the
real library in tests/test_gpu_dropin.py / test_gpu_nonmetric.py."""
import os
import types

import numpy as np
import pandas as pd
import pytest

from oracle import plspm_oracle as orc
from tests.conftest import GOLDEN
from tests.emul import emul

SCHEME_NAMES = {0: "centroid", 1: "factorial", 2: "path"}


class FakeModel:
    def __init__(self, block_sizes, modes, path, scaled, tile_policy=0, numeric=False):
        self.block_sizes = [int(v) for v in block_sizes]
        self.modes = [int(v) for v in modes]
        self.path = np.ascontiguousarray(path, dtype=np.int8)
        self.scaled, self.numeric, self.tile_policy = bool(scaled), bool(numeric), int(tile_policy)
        info = emul.model_info(self.block_sizes, self.modes, self.path, self.scaled, self.tile_policy)
        self.L, self.P = len(self.block_sizes), info["P"]
        self.n_out, self.n_effects, self.full_tiles = info["n_out"], info["n_eff"], bool(info["full"])
        self.effects_from, self.effects_to = info["eff_from"], info["eff_to"]

    def split_row(self, rows):
        P, L, E = self.P, self.L, self.n_effects
        return (rows[..., :P], rows[..., P:P + L], rows[..., P + L:P + L + E], rows[..., P + L + E:P + L + 2 * E],
                rows[..., P + L + 2 * E:])

    def close(self):
        pass


class FakeData:
    def __init__(self, model, X):
        self.model, self.X, self.N = model, np.ascontiguousarray(X, dtype=np.float64), int(np.asarray(X).shape[0])

    def close(self):
        pass


def _scheme(s):
    return SCHEME_NAMES[s] if isinstance(s, (int, np.integer)) else s


def _run(model, data, scheme, tol, max_iter, idx=None):
    if model.numeric:
        return emul.fit_num(data.X, model.block_sizes, model.modes, model.path, _scheme(scheme), idx=idx, tol=tol,
                            max_iter=max_iter, tile_policy=model.tile_policy)
    return emul.fit(data.X, model.block_sizes, model.modes, model.path, _scheme(scheme), model.scaled, idx=idx, tol=tol,
                    max_iter=max_iter, tile_policy=model.tile_policy)


def fake_fit(model, data, scheme, tol=1e-6, max_iter=100, want_scores=True):
    return _run(model, data, scheme, tol, max_iter)


def fake_bootstrap(model, data, scheme, rep_begin, rep_count, seed=0, idx=None, tol=1e-6, max_iter=100, out_device_ptr=0):
    rows = np.zeros((rep_count, model.n_out))
    status = np.zeros(rep_count, dtype=np.int32)
    iters = np.zeros(rep_count, dtype=np.int32)
    for b in range(rep_count):
        pick = idx[b] if idx is not None else orc.philox_indices(seed, rep_begin + b, data.N)
        r = _run(model, data, scheme, tol, max_iter, idx=pick)
        rows[b], status[b], iters[b] = r["out_row"], r["status"], r["iterations"]
    return rows, status, iters


@pytest.fixture()
def host_only(monkeypatch):
    """plspm_b200.engine replaced by the host emulation inside the session / bootstrap modules."""
    import plspm_b200.session as session
    from plspm_b200 import engine as real
    fake = types.SimpleNamespace(
        Model=FakeModel, Data=FakeData, fit=fake_fit, bootstrap=fake_bootstrap,
        resample_indices=lambda seed, rep, n: orc.philox_indices(seed, rep, n),
        TILES_AUTO=real.TILES_AUTO, TILES_FULL=real.TILES_FULL, TILES_SPARSE=real.TILES_SPARSE,
        STATUS_OK=real.STATUS_OK, STATUS_NOT_CONVERGED=real.STATUS_NOT_CONVERGED, STATUS_SINGULAR=real.STATUS_SINGULAR,
        EngineError=real.EngineError, pool_trim=lambda: None)
    monkeypatch.setattr(session, "engine", fake)
    import plspm_b200
    monkeypatch.setattr(plspm_b200, "engine", fake, raising=False)
    return fake


def satisfaction_config(sat, frame, mode, scaled=False):
    import plspm.config as c
    lvs = [str(v) for v in sat["lvs"]]
    path = pd.DataFrame(sat["path"], index=lvs, columns=lvs)
    config = c.Config(path, scaled=scaled)
    mvs = [str(v) for v in sat["mvs"]]
    o = 0
    for lv, k in zip(lvs, sat["block_sizes"]):
        config.add_lv(lv, mode, *[c.MV(m) for m in mvs[o:o + int(k)]])
        o += int(k)
    return config, lvs, mvs


def test_satisfaction_tables_match_r_values(host_only, sat):
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scheme import Scheme
    frame = pd.DataFrame(sat["X"], columns=[str(v) for v in sat["mvs"]])
    config, lvs, mvs = satisfaction_config(sat, frame, Mode.A)
    calc = Plspm(frame, config, Scheme.CENTROID)
    assert calc.iterations() == 4
    om = calc.outer_model()
    np.testing.assert_allclose(om.loc[mvs, "weight"], sat["R/centroid/weight"], rtol=1e-6)
    np.testing.assert_allclose(om.loc[mvs, "loading"], sat["R/centroid/loading"], rtol=1e-6)
    np.testing.assert_allclose(calc.scores().loc[:, lvs].to_numpy(), sat["R/scores"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(calc.crossloadings().loc[mvs, lvs].to_numpy(), sat["R/crossloadings"], rtol=1e-6, atol=1e-9)
    effects = calc.effects()
    for f, t, d, tot in zip(sat["R/effects_from"], sat["R/effects_to"], sat["R/effects_direct"], sat["R/effects_total"]):
        row = effects.loc["%s -> %s" % (f, t)]
        np.testing.assert_allclose([row["direct"], row["total"]], [d, tot], rtol=1e-6, atol=1e-10)
    summary = calc.inner_summary()
    assert set(summary.index) == set(lvs) and 0.0 < calc.goodness_of_fit() < 1.0
    inner = calc.inner_model()
    assert {"estimate", "std error", "t", "p>|t|"} <= set(inner.columns)
    uni = calc.unidimensionality()
    assert (uni.loc[:, "cronbach_alpha"] > 0.5).all()


def test_bootstrap_summaries_on_the_host(host_only, sat):
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scheme import Scheme
    frame = pd.DataFrame(sat["X"], columns=[str(v) for v in sat["mvs"]])
    config, lvs, mvs = satisfaction_config(sat, frame, Mode.A)
    idx = np.random.default_rng(1234).integers(0, 250, (20, 250), dtype=np.int32)
    calc = Plspm(frame, config, Scheme.CENTROID, bootstrap=True, bootstrap_iterations=20, processes=1,
                 bootstrap_indices=idx)
    boot = calc.bootstrap()
    status, iters = boot.replicate_status()
    assert (status == 0).all() and (iters >= 3).all()
    weights = boot.weights()
    assert list(weights.columns) == ["original", "mean", "std.error", "perc.025", "perc.975", "t stat."]
    np.testing.assert_allclose(weights.loc[mvs, "original"], calc.outer_model().loc[mvs, "weight"], rtol=1e-12)
    # replicate rows equal the oracle's on the same resamples
    rows, oiters, ostatus = orc.bootstrap(sat["X"], idx, sat["block_sizes"], [0] * 6, sat["path"], "centroid", False)
    P = len(mvs)
    np.testing.assert_allclose(boot.samples()["weights"].loc[:, mvs].to_numpy(), rows[:, :P], rtol=1e-8)
    assert len(boot.paths()) == 10 and len(boot.r_squared()) == 5 and len(boot.total_effects()) == 15


def test_scale_num_and_mixed_raw(host_only):
    import plspm.config as c
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scale import Scale
    from plspm.scheme import Scheme
    nm = np.load(os.path.join(GOLDEN, "nonmetric.npz"), allow_pickle=False)
    lvs, mvs = [str(v) for v in nm["russa/lvs"]], [str(v) for v in nm["russa/mvs"]]
    frame = pd.DataFrame(nm["russa/X"], columns=mvs)
    path = pd.DataFrame(nm["russa/path"], index=lvs, columns=lvs)
    for scheme, tag in ((Scheme.CENTROID, "centroid"), (Scheme.PATH, "path"), (Scheme.FACTORIAL, "factorial")):
        config = c.Config(path, default_scale=Scale.NUM)
        o = 0
        for lv, k in zip(lvs, nm["russa/block_sizes"]):
            config.add_lv(lv, Mode.A, *[c.MV(m, Scale.RAW if (o + i) % 2 else None) for i, m in enumerate(mvs[o:o + int(k)])])
            o += int(k)
        calc = Plspm(frame, config, scheme, 100, 1e-7)
        om = calc.outer_model()
        np.testing.assert_allclose(om.loc[mvs, "weight"], nm["russa/%s/A/weights" % tag], rtol=1e-6)
        np.testing.assert_allclose(om.loc[mvs, "loading"], nm["russa/%s/A/loadings" % tag], rtol=1e-6)
        np.testing.assert_allclose(om.loc[mvs, "weight"], nm["R/russa/%s/weight" % tag], rtol=1e-6)
        np.testing.assert_allclose(calc.path_coefficients().loc[lvs, lvs].to_numpy(), nm["russa/%s/A/path_coefficients" % tag],
                                   rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("tag", ("path", "centroid_b"))
def test_higher_order_construct_two_stage(host_only, tag):
    """the reference's own test (tests/test_regression_seminr.py:49-74) on the CPU"""
    import plspm.config as c
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    from plspm.scale import Scale
    from plspm.scheme import Scheme
    hz = np.load(os.path.join(GOLDEN, "hoc.npz"), allow_pickle=False)
    frame = pd.DataFrame(hz["mobi/X"], columns=[str(v) for v in hz["mobi/mvs"]])
    prefix = {"Expectation": "CUEX", "Quality": "PERQ", "Loyalty": "CUSL", "Image": "IMAG", "Complaints": "CUSCO", "Value": "PERV"}
    scheme, hoc_mode, tol = (Scheme.PATH, Mode.A, 1e-8) if tag == "path" else (Scheme.CENTROID, Mode.B, 1e-7)
    structure = c.Structure()
    structure.add_path(["Expectation", "Quality"], ["Satisfaction"])
    structure.add_path(["Satisfaction"], ["Complaints", "Loyalty"])
    config = c.Config(structure.path(), default_scale=Scale.NUM)
    config.add_higher_order("Satisfaction", hoc_mode, ["Image", "Value"])
    for lv in ("Expectation", "Quality", "Loyalty", "Image", "Complaints", "Value"):
        config.add_lv_with_columns_named(lv, Mode.B if lv == "Quality" else Mode.A, frame, prefix[lv])
    calc = Plspm(frame, config, scheme, 100, tol)
    lvs = [str(v) for v in hz[tag + "/lvs"]]
    index = [str(v) for v in hz[tag + "/outer_index"]]
    om = calc.outer_model()
    assert set(om.index) == set(index)
    np.testing.assert_allclose(om.loc[index, "weight"], hz[tag + "/weights"], rtol=1e-6)
    np.testing.assert_allclose(om.loc[index, "loading"], hz[tag + "/loadings"], rtol=1e-6)
    np.testing.assert_allclose(calc.path_coefficients().loc[lvs, lvs].to_numpy(), hz[tag + "/path_coefficients"],
                               rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(calc.scores().loc[:, lvs].to_numpy(), hz[tag + "/scores"], rtol=1e-6, atol=1e-8)
    if tag == "path":  # bootstrap of a construct: two engine fits per replicate, replicate ids -> Philox indices
        config_b = c.Config(structure.path(), default_scale=Scale.NUM)
        config_b.add_higher_order("Satisfaction", hoc_mode, ["Image", "Value"])
        for lv in ("Expectation", "Quality", "Loyalty", "Image", "Complaints", "Value"):
            config_b.add_lv_with_columns_named(lv, Mode.B if lv == "Quality" else Mode.A, frame, prefix[lv])
        boot = Plspm(frame, config_b, scheme, 100, tol, bootstrap=True, bootstrap_iterations=10, processes=1,
                     bootstrap_seed=3).bootstrap()
        status, _ = boot.replicate_status()
        assert (status == 0).all()
        assert {"Image", "Value"} <= set(boot.weights().index) and np.isfinite(boot.paths().to_numpy()).all()
        assert abs(boot.paths().loc["Satisfaction -> Loyalty", "mean"] - 0.63) < 0.1
    # a metric configuration cannot estimate a construct (neither can the reference's metric path)
    metric = c.Config(structure.path())
    metric.add_higher_order("Satisfaction", Mode.A, ["Image", "Value"])
    for lv in ("Expectation", "Quality", "Loyalty", "Image", "Complaints", "Value"):
        metric.add_lv_with_columns_named(lv, Mode.A, frame, prefix[lv])
    with pytest.raises(NotImplementedError):
        Plspm(frame, metric, scheme)


def test_not_converged_and_argument_rules(host_only, sat):
    from plspm.mode import Mode
    from plspm.plspm import Plspm
    frame = pd.DataFrame(sat["X"], columns=[str(v) for v in sat["mvs"]])
    config, _, _ = satisfaction_config(sat, frame, Mode.B, scaled=True)
    with pytest.raises(Exception, match="Could not converge after 101 iterations"):
        Plspm(frame, config, tolerance=1e-300)
    config, _, _ = satisfaction_config(sat, frame, Mode.A)
    with pytest.raises(AssertionError):
        Plspm(frame, config, tolerance=0)
    config, _, _ = satisfaction_config(sat, frame, Mode.A)
    with pytest.raises(Exception, match="at least 10 observations"):
        Plspm(frame.iloc[:9], config, bootstrap=True)
