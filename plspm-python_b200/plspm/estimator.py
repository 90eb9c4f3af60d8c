"""Estimator (reference plspm/estimator.py:24-74): per-fit orchestration.

The reference clones the calculator, treats the data on the host and runs the weight iteration
(twice, quirk Q1: estimator.py:36,43,52).  Here the raw filtered data go to the device once; the
treatment (config.py:299-305) is folded into the covariance the solver works on, and the iteration
runs once.  Higher-order constructs (two-stage, estimator.py:41-52) are not on the accelerated
path yet.
"""
from typing import Tuple

import pandas as pd


class Estimator:
    def __init__(self, config):
        if config.hoc():
            raise NotImplementedError("higher order constructs are outside the accelerated path of plspm_b200")
        self._config = config
        self._last = None

    def estimate(self, calculator, data: pd.DataFrame, want_final_data: bool = True) -> Tuple[pd.DataFrame, pd.DataFrame, pd.DataFrame]:
        config = calculator.config()
        session = calculator.session(data)
        res, scores, weights = calculator.run(session)
        self._config = config
        self._last = (session, res)
        final_data = config.treat(data).loc[:, session.mvs] if want_final_data else None
        return final_data, scores, weights

    def config(self):
        return self._config

    def last_result(self):
        """(EngineSession, raw engine outputs) of the most recent estimate()."""
        return self._last
