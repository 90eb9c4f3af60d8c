"""Unidimensionality diagnostics (reference plspm/unidimensionality.py:30-63): Cronbach's alpha,
Dillon-Goldstein's rho and the first two eigenvalues per block.

Independent of the weight-estimation path and computed lazily on the host
from K x K block correlation matrices: the reference's PCA of the standardised block
(scikit-learn) has the eigenvalues of the block correlation matrix as its component variances and
eigenvector * sqrt(eigenvalue) as the correlations with the first component.
"""
import numpy as np
import pandas as pd

from plspm.mode import Mode


class Unidimensionality:
    def __init__(self, config, data: pd.DataFrame, correction: float):
        self._config, self._data, self._correction = config, data, correction

    def summary(self) -> pd.DataFrame:
        cfg = self._config
        lvs = list(cfg.path())
        out = pd.DataFrame({"mode": pd.Series(dtype="str"), "mvs": pd.Series(dtype="float"),
                            "cronbach_alpha": pd.Series(dtype="float"),
                            "dillon_goldstein_rho": pd.Series(dtype="float"), "eig_1st": pd.Series(dtype="float"),
                            "eig_2nd": pd.Series(dtype="float")}, index=lvs)
        for lv in lvs:
            mvs = cfg.mvs(lv)
            k = len(mvs)
            out.loc[lv, "mode"] = cfg.mode(lv).name
            out.loc[lv, "mvs"] = k
            if not set(mvs) <= set(self._data.columns):
                continue  # a higher-order construct: its "manifest variables" are stage-1 scores, not data columns
            block = self._data.loc[:, mvs].to_numpy(dtype=np.float64)
            if np.isnan(block).any():
                continue
            if block.shape[0] <= k:
                # no more observations than manifest variables: the reference runs its PCA on the TRANSPOSED block
                # (unidimensionality.py:46), i.e. over the observations; mirrored literally (SVD of the column-centred
                # transposed block = scikit-learn's PCA, including its sign convention)
                self._transposed(out, lv, block, k)
                continue
            corr = np.atleast_2d(np.corrcoef(block, rowvar=False))
            evals, evecs = np.linalg.eigh(corr)
            evals, evecs = evals[::-1], evecs[:, ::-1]
            out.loc[lv, "eig_1st"] = evals[0]
            out.loc[lv, "eig_2nd"] = evals[1] if k > 1 else np.nan
            if cfg.mode(lv) == Mode.A:
                if k > 1:
                    off = 2.0 * np.tril(corr, -1).sum()
                    out.loc[lv, "cronbach_alpha"] = max(0.0, (off / (k + off)) * (k / (k - 1)))
                load = evecs[:, 0] * np.sqrt(evals[0])
                num = load.sum() ** 2
                out.loc[lv, "dillon_goldstein_rho"] = num / (num + (k - (load ** 2).sum()))
        return out

    def _transposed(self, out: pd.DataFrame, lv: str, block: np.ndarray, k: int):
        z = (block - block.mean(axis=0)) / block.std(axis=0, ddof=1) * self._correction
        inp = z.T                                            # [k MVs x N observations]
        centred = inp - inp.mean(axis=0)
        U, S, Vt = np.linalg.svd(centred, full_matrices=False)
        flip = np.sign(U[np.argmax(np.abs(U), axis=0), np.arange(U.shape[1])])  # sklearn.utils.extmath.svd_flip (u-based)
        flip[flip == 0] = 1.0
        scores = U * S * flip
        sd = scores.std(axis=0)
        out.loc[lv, "eig_1st"] = sd[0] ** 2
        out.loc[lv, "eig_2nd"] = sd[1] ** 2 if k > 1 else np.nan
        if self._config.mode(lv) == Mode.A:
            if k > 1:
                num = 2.0 * np.tril(pd.DataFrame(inp).corr().to_numpy(), -1).sum()
                den = inp.sum(axis=1).var(ddof=1) / self._correction ** 2
                out.loc[lv, "cronbach_alpha"] = max(0.0, (num / den) * (k / (k - 1)))
            corr = np.corrcoef(np.column_stack((inp, scores[:, 0])), rowvar=False)[:-1, -1]
            num = corr.sum() ** 2
            out.loc[lv, "dillon_goldstein_rho"] = num / (num + (k - (corr ** 2).sum()))
