"""OuterModel (reference plspm/outer_model.py:21-42): weight, loading, communality, redundancy per
manifest variable, and the crossloadings table.  The P x L correlations between manifest variables
and scores are a by-product of the solver's final pass, so nothing here
touches N-length data."""
import numpy as np
import pandas as pd


class OuterModel:
    def __init__(self, weights: pd.DataFrame, loadings: np.ndarray, crossloadings: np.ndarray, lvs: list,
                 blocks: dict, r_squared: pd.Series):
        mvs = list(weights.index)
        self._crossloadings = pd.DataFrame(crossloadings, index=mvs, columns=lvs)
        loading = pd.Series(loadings, index=mvs, name="loading")
        communality = (loading ** 2).rename("communality")
        lv_of = {mv: lv for lv in lvs for mv in blocks[lv]}
        redundancy = (communality * np.array([r_squared[lv_of[mv]] for mv in mvs])).rename("redundancy")
        self._model = pd.concat([weights, loading, communality, redundancy], axis=1)

    def model(self) -> pd.DataFrame:
        return self._model

    def crossloadings(self) -> pd.DataFrame:
        return self._crossloadings
