"""Bootstrap (reference plspm/bootstrap.py): row-resampling bootstrap of the whole estimation.

Reference: p forked processes x (iterations // p) replicates, each a full host-side fit, results
pickled through a Queue (bootstrap.py:83-113).  Here the replicates are the batch dimension of the
CUDA kernels on the HBM-resident data; with torch.distributed initialised (one process per GPU)
the global replicate ids are split across ranks and one all-gather returns every row to every
rank.  `num_processes` is kept for signature compatibility and ignored.

Differences that are deliberate: replicates are independent draws keyed
by (seed, global replicate id) -- the reference's forked workers share one RNG state and repeat
each other's resamples; failed replicates are dropped like the reference's bare `except: pass`.
"""
import numpy as np
import pandas as pd

from plspm_b200 import distributed as pdist


def _create_summary(data: pd.DataFrame, original) -> pd.DataFrame:
    """original / mean / std.error / perc.025 / perc.975 / t stat. per column (bootstrap.py:24-32)."""
    sd = data.std(axis=0)
    summary = pd.DataFrame({"original": pd.Series(original).reindex(data.columns).astype(float),
                            "mean": data.mean(axis=0), "std.error": sd, "perc.025": data.quantile(0.025, axis=0),
                            "perc.975": data.quantile(0.975, axis=0)}, index=data.columns)
    summary["t stat."] = summary["original"] / sd
    return summary


class Bootstrap:
    """Bootstrap validation results; returned by :meth:`plspm.Plspm.bootstrap`."""

    def __init__(self, config, data: pd.DataFrame, inner_model, outer_model, calculator, iterations: int,
                 num_processes: int = 1, seed: int = None, indices: np.ndarray = None):
        rank, world = pdist.rank_world()
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        seed = pdist.broadcast_int(seed)
        begin, count = pdist.shard_range(iterations, rank, world)
        idx = None if indices is None else np.ascontiguousarray(indices[begin:begin + count], dtype=np.int32)
        scheme, tol, its = calculator.scheme(), calculator.tolerance(), calculator.iterations()
        two_stage = bool(calculator.config().hoc())
        buf = None
        if two_stage:
            model, lvs, mvs, rows, status, iters = self._two_stage_rows(calculator, data, begin, count, seed, idx)
        else:
            session = calculator.session(data)
            model, lvs, mvs = session.model, session.lvs, session.mvs
            buf = None if getattr(session, "host", False) else pdist.send_buffer(iterations, model.n_out)
        if two_stage:
            pass  # rows were produced replicate by replicate on the host side of the two engine fits
        elif buf is not None:  # NCCL: the solver writes its rows straight into the all-gather send buffer
            _, status, iters = session.bootstrap(scheme, tol, its, begin, count, seed, idx,
                                                 out_device_ptr=buf.data_ptr())
            rows = None
        else:
            rows, status, iters = session.bootstrap(scheme, tol, its, begin, count, seed, idx)
        rows, status, iters = pdist.allgather_rows(rows, status, iters, iterations, model.n_out, device_buffer=buf)
        self._status, self._iterations, self._seed = status, iters, seed
        if not (status == 0).any():  # (after the collective: every rank sees the same statuses and raises together)
            raise Exception("Bootstrapping failed: no replicate could be estimated")
        rows = rows[status == 0]  # bootstrap.py:67-68
        w, r2, total, direct, load = model.split_row(rows)
        cols = mvs if two_stage else list(data.columns)  # stage-2 manifest variables include the constituents
        weights = pd.DataFrame(w, columns=mvs).loc[:, cols]
        loadings = pd.DataFrame(load, columns=mvs).loc[:, cols]
        r_squared = pd.DataFrame(r2, columns=lvs)
        names = [lvs[f] + " -> " + lvs[t] for f, t in zip(model.effects_from, model.effects_to)]
        eff_index = list(inner_model.effects().index)
        total_effects = pd.DataFrame(total, columns=names).reindex(columns=eff_index)
        paths = pd.DataFrame(direct, columns=names).reindex(columns=eff_index)
        om = outer_model.model()
        self._weights = _create_summary(weights, om.loc[:, "weight"])
        self._r_squared = _create_summary(r_squared, inner_model.r_squared()).loc[inner_model.endogenous(), :]
        self._total_effects = _create_summary(total_effects, inner_model.effects().loc[:, "total"])
        self._paths = _create_summary(paths, inner_model.effects().loc[:, "direct"])
        self._loading = _create_summary(loadings, om.loc[:, "loading"])
        self._samples = dict(weights=weights, r_squared=r_squared, total_effects=total_effects, paths=paths,
                             loadings=loadings)

    @staticmethod
    def _two_stage_rows(calculator, data, begin, count, seed, idx):
        """Higher-order constructs: the stage-2 manifest variables are the replicate's own stage-1 scores, so
        every replicate is a two-stage estimate of its resampled rows (bootstrap.py:54-66 as written: two
        engine fits per replicate).  Failed replicates keep status != 0 and are dropped (bootstrap.py:67-68).
        Nothing raises here: a rank whose replicates all fail must still reach the all-gather (the row layout
        comes from the configuration, not from a successful fit)."""
        from plspm.estimator import Estimator
        from plspm_b200 import engine
        config = calculator.config()
        estimator = Estimator(config)
        # layout of the stage-2 (structural) model: a construct's block = the scores of its constituents
        path = config.path()
        lvs = list(path)
        blocks = {lv: (list(config.hoc()[lv]) if lv in config.hoc() else list(config.mvs(lv))) for lv in lvs}
        mvs = [mv for lv in lvs for mv in blocks[lv]]
        model = engine.Model([len(blocks[lv]) for lv in lvs], [config.mode(lv).value.engine_id for lv in lvs],
                             path.loc[lvs, lvs].to_numpy(dtype=np.int8), True, numeric=True)
        n = data.shape[0]
        rows = np.zeros((count, model.n_out))
        status, iters = np.ones(count, dtype=np.int32), np.zeros(count, dtype=np.int32)
        ef, et = model.effects_from, model.effects_to
        for b in range(count):
            pick = idx[b] if idx is not None else engine.resample_indices(seed, begin + b, n)
            try:
                estimator.estimate(calculator, data.iloc[pick, :].reset_index(drop=True), want_final_data=False)
            except NotImplementedError:
                raise
            except Exception:
                continue
            session, res = estimator.last_result()
            assert session.mvs == mvs and session.lvs == lvs
            rows[b] = np.concatenate([res["weights"], res["r_squared"], res["total_effects"][et, ef],
                                      res["path_coefficients"][et, ef], res["loadings"]])
            status[b], iters[b] = 0, res["iterations"]
        return model, lvs, mvs, rows, status, iters

    def weights(self) -> pd.DataFrame:
        """Outer weights calculated from bootstrap validation."""
        return self._weights

    def r_squared(self) -> pd.DataFrame:
        """R squared for latent variables calculated from bootstrap validation."""
        return self._r_squared

    def total_effects(self) -> pd.DataFrame:
        """Total effects for paths calculated from bootstrap validation."""
        return self._total_effects

    def paths(self) -> pd.DataFrame:
        """Direct effects for paths calculated from bootstrap validation."""
        return self._paths[self._paths["mean"] != 0]

    def loading(self) -> pd.DataFrame:
        """Loadings of manifest variables calculated from bootstrap validation."""
        return self._loading

    # extras (not in the reference)
    def samples(self) -> dict:
        """Per-replicate values behind the summaries (DataFrames keyed weights / r_squared / ...)."""
        return self._samples

    def replicate_status(self):
        """(status, iterations) per global replicate id; status 0 = converged."""
        return self._status, self._iterations
