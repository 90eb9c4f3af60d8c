"""statsmodels.api stand-in: OLS via pinv + add_constant (see package docstring)."""
import numpy as np
import pandas as pd
from scipy import stats as _stats


class _Fit:
    def __init__(self, params, bse, tvalues, pvalues, rsquared):
        self.params, self.bse, self.tvalues, self.pvalues, self.rsquared = params, bse, tvalues, pvalues, rsquared


class OLS:
    def __init__(self, endog, exog):
        self._names = list(exog.columns) if isinstance(exog, pd.DataFrame) else None
        self._y = np.asarray(endog, dtype=np.float64).reshape(-1)
        x = np.asarray(exog, dtype=np.float64)
        self._x = x.reshape(-1, 1) if x.ndim == 1 else x

    def fit(self):
        x, y = self._x, self._y
        pinv = np.linalg.pinv(x)
        beta = pinv @ y
        resid = y - x @ beta
        ssr = float(resid @ resid)
        df_resid = x.shape[0] - np.linalg.matrix_rank(x)
        cov = (ssr / df_resid) * (pinv @ pinv.T)
        bse = np.sqrt(np.diag(cov))
        with np.errstate(divide="ignore", invalid="ignore"):
            t = beta / bse
        p = 2.0 * _stats.t.sf(np.abs(t), df_resid)
        has_const = bool(np.any((np.ptp(x, axis=0) == 0) & (x[0, :] != 0)))
        tss = float(((y - y.mean()) ** 2).sum()) if has_const else float((y ** 2).sum())
        rsq = 1.0 - ssr / tss
        if self._names is not None:
            wrap = lambda v: pd.Series(v, index=self._names)
            return _Fit(wrap(beta), wrap(bse), wrap(t), wrap(p), rsq)
        return _Fit(beta, bse, t, p, rsq)


def add_constant(data, prepend=True, has_constant="skip"):
    if isinstance(data, (pd.DataFrame, pd.Series)):
        frame = data.to_frame() if isinstance(data, pd.Series) else data
        vals = frame.values.astype(np.float64)
        if has_constant == "skip" and np.any((np.ptp(vals, axis=0) == 0) & (vals[0, :] != 0)):
            return frame
        out = frame.copy()
        out.insert(0 if prepend else out.shape[1], "const", 1.0)
        return out
    x = np.asarray(data, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    if has_constant == "skip" and np.any((np.ptp(x, axis=0) == 0) & (x[0, :] != 0)):
        return x
    ones = np.ones((x.shape[0], 1))
    return np.column_stack((ones, x) if prepend else (x, ones))
