"""Lowering of a plspm Config + DataFrame to engine handles, kept alive so that the single fit and
every bootstrap replicate reuse the same HBM-resident observation matrix."""
from __future__ import annotations

import numpy as np
import pandas as pd

from plspm_b200 import engine


class EngineSession:
    def __init__(self, config, data: pd.DataFrame, path: pd.DataFrame = None, scaled=None, tile_policy: int = 0):
        self.numeric = not config.metric()
        if self.numeric:
            config.check_scales()  # TypeError / NotImplementedError for anything but all-NUM / RAW scales
        path = config.path() if path is None else path
        self.lvs = list(path)
        self.blocks = {lv: list(config.mvs(lv)) for lv in self.lvs}
        self.mvs = [mv for lv in self.lvs for mv in self.blocks[lv]]  # ODM row order (config.py:140-144)
        self.index = data.index
        self.path = path
        self.missing = bool(getattr(config, "missing", lambda: False)())
        frame = data.loc[:, self.mvs]
        if self.missing and self.numeric:
            raise NotImplementedError("non-metric data with missing values are outside the accelerated path")
        self.raw = None
        if self.missing:
            self.raw = np.ascontiguousarray(frame.to_numpy(dtype=np.float64))  # (with NaN: the bootstrap re-imputes per replicate)
            frame = frame.fillna(frame.mean(skipna=True))  # util.impute (util.py:61-68): the single fit
        X = np.ascontiguousarray(frame.to_numpy(dtype=np.float64))
        self.scaled = config.scaled() if scaled is None else bool(scaled)
        spec = ([len(self.blocks[lv]) for lv in self.lvs], [config.mode(lv).value.engine_id for lv in self.lvs],
                path.loc[self.lvs, self.lvs].to_numpy(dtype=np.int8), self.scaled)
        self.model = engine.Model(*spec, tile_policy, numeric=self.numeric)
        # The numeric non-metric solver has no cross-moment pass: the single fit (which reports crossloadings)
        # uses a full-tile twin of the model on the same resident data; bootstrap keeps the sparse tile set.
        self.fit_model = self.model
        if self.numeric and not self.model.full_tiles:
            self.fit_model = engine.Model(*spec, engine.TILES_FULL, numeric=True)
        self.data = engine.Data(self.model, X)
        self.N = self.data.N
        self.spec = spec
        self.aug = None  # (base model, augmented model, augmented data) of the bootstrap with missing values

    def _augmented(self):
        """Handles for the bootstrap on data with missing values: the matrix [x0 | m] (missing entries as 0, one 0/1
        indicator per column that has missing entries, appended to its block) under full-tile models.  Every moment
        of a replicate imputed with ITS observed column means (util.py:61-68 under bootstrap.py:57) is a closed form
        in the moments of this matrix (csrc/kernels_impute.cuh)."""
        if self.aug is None:
            sizes, modes, path, scaled = self.spec
            nan = np.isnan(self.raw)
            has = nan.any(axis=0)
            cols, aug_sizes, o = [], [], 0
            for k in sizes:
                block = list(range(o, o + k))
                miss = [p for p in block if has[p]]
                cols.append(np.where(nan[:, block], 0.0, self.raw[:, block]))
                if miss:
                    cols.append(nan[:, miss].astype(np.float64))
                aug_sizes.append(k + len(miss))
                o += k
            Xa = np.ascontiguousarray(np.concatenate(cols, axis=1))
            base = engine.Model(sizes, modes, path, scaled, engine.TILES_FULL)
            aug_model = engine.Model(aug_sizes, modes, path, scaled, engine.TILES_FULL)
            aug_data = engine.Data(aug_model, Xa)
            aug_data.set_imputation(base, has)
            self.aug = (base, aug_model, aug_data)
        return self.aug

    def fit(self, scheme, tol: float, iterations: int, want_scores: bool = True):
        res = engine.fit(self.fit_model, self.data, scheme.value.engine_id, tol, iterations, want_scores)
        if res["status"] == engine.STATUS_NOT_CONVERGED:  # weights.py:185-186
            raise Exception("Could not converge after " + str(res["iterations"]) + " iterations")
        if res["status"] != engine.STATUS_OK:
            raise Exception("PLS-PM estimation failed: a block or inner regression is singular")
        return res

    def bootstrap(self, scheme, tol: float, iterations: int, rep_begin: int, rep_count: int, seed: int = 0, idx=None,
                  out_device_ptr: int = 0):
        if self.missing:
            base, aug_model, aug_data = self._augmented()
            return engine.bootstrap(aug_model, aug_data, scheme.value.engine_id, rep_begin, rep_count, seed, idx, tol,
                                    iterations, out_device_ptr)
        return engine.bootstrap(self.model, self.data, scheme.value.engine_id, rep_begin, rep_count, seed, idx, tol,
                                iterations, out_device_ptr)

    def close(self, trim_pool: bool = True):
        if self.aug is not None:
            self.aug[2].close()
            self.aug[1].close()
            self.aug[0].close()
            self.aug = None
        self.data.close()
        if self.fit_model is not self.model:
            self.fit_model.close()
        self.model.close()
        if trim_pool:  # the process may share the device with torch / NCCL: give the cached buffers back
            engine.pool_trim()
