"""Estimator (reference plspm/estimator.py:24-74): per-fit orchestration.

The reference clones the calculator, treats the data on the host and runs the weight iteration
(twice, quirk Q1: estimator.py:36,43,52).  Here the raw filtered data go to the device once; the
treatment (config.py:299-305) is folded into the covariance the solver works on, and the iteration
runs once.

Higher-order constructs (two-stage approach, estimator.py:41-52): the stage-1 path expansion
(`hoc_path_first_stage`) is provided and tested against the reference; the estimation itself is
only defined for nonmetric data in the reference (see `estimate`) and is therefore not offered.
"""
from typing import Tuple

import pandas as pd

import plspm.config as c


class Estimator:
    def __init__(self, config):
        self._config = config
        self._last = None
        self._first_stage_path = self.hoc_path_first_stage(config) if config.hoc() else None

    def estimate(self, calculator, data: pd.DataFrame, want_final_data: bool = True) -> Tuple[pd.DataFrame, pd.DataFrame, pd.DataFrame]:
        config = calculator.config()
        if not config.hoc():
            session = calculator.session(data)
            res, scores, weights = calculator.run(session)
            self._config = config
            self._last = (session, res)
            final_data = config.treat(data).loc[:, session.mvs] if want_final_data else None
            return final_data, scores, weights
        # Higher-order constructs: in the reference only the NONMETRIC path can run the two-stage
        # approach -- its metric path fails in stage 2 ("matrices are not aligned", weights.py:30: the
        # stage-1 score columns are not in the outer design matrix) -- so there is no reference behaviour
        # to reproduce for metric data, and the nonmetric path is outside the accelerated scope (f3).
        raise NotImplementedError("higher order constructs need the nonmetric path, which is outside the "
                                  "accelerated path of plspm_b200")

    def config(self):
        return self._config

    def last_result(self):
        """(EngineSession, raw engine outputs) of the most recent estimate()."""
        return self._last

    def hoc_path_first_stage(self, config) -> pd.DataFrame:
        """Path matrix of stage 1 (reference estimator.py:60-74): every predecessor of a HOC points to
        all of its constituent LVs, every constituent points to the HOC's successors, the HOC is dropped."""
        path = config.path()
        for hoc, members in config.hoc().items():
            structure = c.Structure(path)
            into = path.loc[hoc]
            out_of = path.loc[:, hoc]
            for lv in list(into[into == 1].index):
                structure.add_path([lv], members)
            for lv in list(out_of[out_of == 1].index):
                structure.add_path(members, [lv])
            path = structure.path().drop(hoc).drop(hoc, axis=1)
        return path
