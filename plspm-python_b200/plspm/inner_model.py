"""InnerModel (reference plspm/inner_model.py): path coefficients, R^2, effects, and the
regression table (std error, t, p).

Path coefficients, R^2 and total effects come from the CUDA solver (they are also what every
bootstrap replicate produces).  The inference columns are host post-processing on L x L
quantities only: with centred unit-variance scores, X'X of the
regression with intercept is diag(N, N R_pp), so std errors need just R and N.
"""
import numpy as np
import pandas as pd
from scipy import stats


def effects_table(path_coefficients: pd.DataFrame) -> pd.DataFrame:
    """from / to / direct / indirect / total rows in the reference's order (inner_model.py:33-61)."""
    lvs = list(path_coefficients)
    B = path_coefficients.values.astype(np.float64)
    n = len(lvs)
    indirect = np.zeros_like(B)
    if n != 2:
        power = B.copy()
        for _ in range(1, n):
            power = power @ B
            indirect = indirect + power
    total = B + indirect
    rows, index = [], []
    for f in range(n):
        for t in range(n):
            if f != t and total[t, f] != 0:
                rows.append({"from": lvs[f], "to": lvs[t], "direct": B[t, f], "indirect": indirect[t, f],
                             "total": total[t, f]})
                index.append(lvs[f] + " -> " + lvs[t])
    cols = ["from", "to", "direct", "indirect", "total"]
    return pd.DataFrame(rows, index=index, columns=cols) if rows else pd.DataFrame(columns=cols)


class InnerModel:
    def __init__(self, path: pd.DataFrame, scores: pd.DataFrame, path_coefficients: np.ndarray = None,
                 r_squared: np.ndarray = None):
        lvs = list(path.index)
        n_obs = scores.shape[0]
        S = scores.loc[:, lvs].to_numpy(dtype=np.float64)
        Sc = S - S.mean(axis=0)
        R = (Sc.T @ Sc) / n_obs
        sd = np.sqrt(np.diag(R))
        R = R / np.outer(sd, sd)
        pm = path.loc[lvs, lvs].to_numpy()
        self._endogenous = [lv for i, lv in enumerate(lvs) if pm[i].sum() > 0]
        B = np.zeros((len(lvs), len(lvs)))
        r2 = np.zeros(len(lvs))
        r2_adj = np.zeros(len(lvs))
        frames = []
        for i, lv in enumerate(lvs):
            pred = np.where(pm[i] == 1)[0]
            if len(pred) == 0:
                continue
            Rpp_inv = np.linalg.pinv(R[np.ix_(pred, pred)])
            if path_coefficients is not None:
                beta = path_coefficients[i, pred] * sd[pred] / sd[i]   # engine result (standardised scale)
                rsq = float(r_squared[i])
            else:
                beta = Rpp_inv @ R[pred, i]
                rsq = float(beta @ R[pred, i])
            k = len(pred)
            df = n_obs - k - 1
            sigma2 = n_obs * sd[i] ** 2 * (1.0 - rsq) / df
            bse = np.sqrt(sigma2 * np.diag(Rpp_inv) / (n_obs * sd[pred] ** 2))
            est = beta * sd[i] / sd[pred]
            t = est / bse
            B[i, pred] = est
            r2[i] = rsq
            r2_adj[i] = 1 - (1 - rsq) * (n_obs - 1) / (n_obs - k - 1)
            frames.append(pd.DataFrame({"from": [lvs[j] for j in pred], "to": lv, "estimate": est, "std error": bse,
                                        "t": t, "p>|t|": 2.0 * stats.t.sf(np.abs(t), df),
                                        "index": [lvs[j] + " -> " + lv for j in pred]}))
        self._path_coefficients = pd.DataFrame(B, index=lvs, columns=lvs)
        self._r_squared = pd.Series(r2, index=lvs, name="r_squared")
        self._r_squared_adj = pd.Series(r2_adj, index=lvs, name="r_squared_adj")
        self._summaries = pd.concat(frames).reset_index(drop=True) if frames else None
        self._effects = effects_table(self._path_coefficients)

    def path_coefficients(self) -> pd.DataFrame:
        return self._path_coefficients

    def r_squared(self) -> pd.Series:
        return self._r_squared

    def r_squared_adj(self) -> pd.Series:
        return self._r_squared_adj

    def inner_model(self) -> pd.DataFrame:
        return self._summaries.set_index(["index"])

    def effects(self) -> pd.DataFrame:
        return self._effects

    def endogenous(self) -> list:
        return self._endogenous
