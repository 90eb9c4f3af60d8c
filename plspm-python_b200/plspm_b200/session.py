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
        if self.missing:
            frame = frame.fillna(frame.mean(skipna=True))  # util.impute (util.py:61-68), single fit only
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
            raise NotImplementedError("bootstrap with missing values is not supported by the CUDA path yet")
        return engine.bootstrap(self.model, self.data, scheme.value.engine_id, rep_begin, rep_count, seed, idx, tol,
                                iterations, out_device_ptr)

    def close(self, trim_pool: bool = True):
        self.data.close()
        if self.fit_model is not self.model:
            self.fit_model.close()
        self.model.close()
        if trim_pool:  # the process may share the device with torch / NCCL: give the cached buffers back
            engine.pool_trim()
