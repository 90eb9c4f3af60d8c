"""InnerSummary (reference plspm/inner_summary.py:23-65): per-LV type, R^2, block communality, mean
redundancy, AVE and the goodness-of-fit index.  O(P) host arithmetic on the outer-model table."""
import math

import numpy as np
import pandas as pd

from plspm.mode import Mode


class InnerSummary:
    def __init__(self, config, r_squared: pd.Series, r_squared_adj: pd.Series, outer_model: pd.DataFrame):
        path = config.path()
        lvs = list(path)
        endogenous = path.sum(axis=1).astype(bool)
        rows = []
        weighted, sizes = [], []
        for lv in lvs:
            mvs = config.mvs(lv)
            comm = outer_model.loc[mvs, "communality"]
            ave = comm.sum() / (comm.sum() + (1 - comm).sum()) if config.mode(lv) == Mode.A else np.nan
            rows.append({"type": "Endogenous" if endogenous[lv] else "Exogenous", "r_squared": r_squared[lv],
                         "r_squared_adj": r_squared_adj[lv], "block_communality": comm.mean(),
                         "mean_redundancy": outer_model.loc[mvs, "redundancy"].mean(), "ave": ave})
            if len(mvs) > 1:
                sizes.append(len(mvs))
                weighted.append(comm.mean())
        self._summary = pd.DataFrame(rows, index=lvs)
        if sum(sizes) > 0:
            mean_comm = sum(c * k for c, k in zip(weighted, sizes)) / sum(sizes)
            r2_endo = (r_squared * endogenous)
            self._gof = float(np.sqrt(mean_comm * r2_endo[r2_endo != 0].mean()))
        else:
            self._gof = float("NaN")

    def summary(self) -> pd.DataFrame:
        return self._summary

    def goodness_of_fit(self) -> float:
        if math.isnan(self._gof):
            raise ValueError("Cannot calculate goodness-of-fit if all constructs are single-item.")
        return self._gof
