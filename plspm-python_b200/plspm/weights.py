"""WeightsCalculatorFactory (reference plspm/weights.py:157-187).

The reference builds a _MetricWeights object and loops `iterate()` in Python (weights.py:172-187).
Here `calculate` hands the whole loop -- initial weights, score/inner/outer steps, the convergence
test, the final normalisation and sign vote (weights.py:28-70) -- to the CUDA engine and wraps
the result in the DataFrames the reference returns.
"""
import pandas as pd

from plspm.scheme import Scheme
from plspm_b200.session import EngineSession


class WeightsCalculatorFactory:
    def __init__(self, config, iterations: int, tolerance: float, correction: float, scheme: Scheme):
        self._iterations, self._tolerance = iterations, tolerance
        self._config, self._correction, self._scheme = config, correction, scheme
        self._sessions = {}

    def clone(self):
        other = WeightsCalculatorFactory(self._config.clone(), self._iterations, self._tolerance, self._correction,
                                         self._scheme)
        other._sessions = self._sessions  # device-resident data is shared, never copied
        return other

    def config(self):
        return self._config

    def scheme(self):
        return self._scheme

    def tolerance(self):
        return self._tolerance

    def iterations(self):
        return self._iterations

    def session(self, data: pd.DataFrame, path: pd.DataFrame = None, scaled=None) -> EngineSession:
        """Engine handles for (data, path); cached by object identity so bootstrap reuses the upload."""
        path = self._config.path() if path is None else path
        key = (id(data), tuple(path.columns), self._config.scaled() if scaled is None else scaled)
        hit = self._sessions.get(key)
        if hit is None or hit[0] is not data:
            config = self._config
            host = (not config.metric()) and (not config.numeric() or bool(data.loc[:, [mv for lv in list(path) for mv in config.mvs(lv)]].isnull().values.any()))
            if host and not config.hoc():
                # ordinal / nominal scales, non-metric data with missing values: the reference-style host path
                from plspm.nonmetric_host import HostNonmetricSession
                hit = (data, HostNonmetricSession(config, data, path, self._correction))
            else:
                hit = (data, EngineSession(config, data, path, scaled))
            self._sessions[key] = hit
        return hit[1]

    def run(self, session: EngineSession, want_scores: bool = True):
        """(raw engine result dict, scores DataFrame, weights DataFrame)"""
        res = session.fit(self._scheme, self._tolerance, self._iterations, want_scores)
        scores = pd.DataFrame(res["scores"], index=session.index, columns=session.lvs) if want_scores else None
        weights = pd.DataFrame(res["weights"], index=session.mvs, columns=["weight"])
        return res, scores, weights

    def calculate(self, data: pd.DataFrame, path: pd.DataFrame):
        """Reference signature: `data` is ALREADY treated (estimator.py:33,39), so the engine is asked
        not to rescale it (centring is idempotent).  Returns (data, scores, weights)."""
        session = self.session(data, path, scaled=False)
        _, scores, weights = self.run(session)
        return data, scores, weights
