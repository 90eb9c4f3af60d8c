"""Estimator (reference plspm/estimator.py:24-74): per-fit orchestration.

The reference clones the calculator, treats the data on the host and runs the weight iteration
(twice, quirk Q1: estimator.py:36,43,52).  Here the raw filtered data go to the device once; the
treatment (config.py:299-305) is folded into the covariance the solver works on, and the iteration
runs once.

Higher-order constructs (two-stage approach, estimator.py:41-52): two engine fits -- the expanded
first-order model (`hoc_path_first_stage`), then the structural model with the constituents' stage-1
scores as the construct's manifest variables.  Like the reference, only nonmetric (NUM / RAW) data can
run it.
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
        # Higher-order constructs, two-stage approach (estimator.py:41-52).  In the reference only the
        # NONMETRIC path can run it -- the metric path fails in stage 2 ("matrices are not aligned",
        # weights.py:30: the stage-1 score columns are not in the outer design matrix).
        if config.metric():
            raise NotImplementedError("higher order constructs need nonmetric data (set default_scale=Scale.NUM): "
                                      "the reference's metric path cannot estimate them either")
        from plspm.scale import Scale
        from plspm_b200.session import EngineSession
        self._release()
        config = config.clone()  # the construct is added to a private copy (estimator.py:30-31)
        stage1 = EngineSession(config, data, self._first_stage_path)
        try:
            res1 = stage1.fit(calculator.scheme(), calculator.tolerance(), calculator.iterations(), True)
            first = dict(zip(stage1.lvs, res1["scores"].T))
        finally:
            stage1.close()
        data2 = data.copy()
        for hoc, members in config.hoc().items():
            for lv in members:  # stage-1 scores of the constituents become the construct's manifest variables
                data2[lv] = first[lv]
            config.add_lv(hoc, config.mode(hoc), *[c.MV(lv, Scale.NUM) for lv in members])
        session = EngineSession(config, data2, config.path())
        self._owned = session
        res, scores, weights = calculator.run(session)
        self._config = config
        self._last = (session, res)
        final_data = config.treat(data2).loc[:, session.mvs] if want_final_data else None
        return final_data, scores, weights

    def _release(self):
        """Closes the stage-2 engine session of a previous two-stage estimate (they are not cached)."""
        owned = getattr(self, "_owned", None)
        if owned is not None:
            owned.close()
            self._owned = None

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
