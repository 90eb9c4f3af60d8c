"""Host path of the non-metric estimator for what the device path does not cover ("ORD / NOM
stay on the reference-style host path"): ordinal / nominal scales (optimal scaling, reference scale.py:42-89) and
non-metric data with missing values (weights.py:87-94, mode.py:33-37).

This is NOT a fallback of the accelerated path: metric data and all-NUM / RAW non-metric data always run on the
device and fail loudly without it.  Ordinal / nominal quantification regroups the categories of every manifest
variable in every outer iteration (a data-dependent, per-column pooling of adjacent categories), which has no
second-moment form; the reference runs it in NumPy and so does this module, with the same dataflow:

    treat (config.py:306-319) -> initial scores (weights.py:75-98) -> iterate: inner weights (scheme.py:27-54),
    quantify every MV against its LV's inner estimate (scale.py), outer weights + scores (mode.py:31-42, 54-61),
    stop on sum (|y_old| - |y_new|)^2 (weights.py:120) -> final weights (weights.py:122-133).

`HostNonmetricSession` has the interface of `plspm_b200.session.EngineSession` (fit / bootstrap / close), so the
drop-in classes above it do not care which one they hold.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

import plspm.util as util
from plspm.mode import Mode
from plspm.scale import Scale


def _ols(y: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Least-squares coefficients the way statsmodels' OLS computes them (pinv of the design)."""
    return np.linalg.pinv(X) @ y


def inner_weights(letter: str, path: np.ndarray, y: np.ndarray) -> np.ndarray:
    """E[j, i] = weight of score j in the inner estimate of LV i (scheme.py:27-28, 36-37, 45-54)."""
    link = path + path.T
    if letter == "C":
        return np.sign(np.corrcoef(y, rowvar=False) * link)
    if letter == "F":
        return np.cov(y, rowvar=False) * link
    E = path.astype(np.float64).copy()
    for i in range(path.shape[0]):
        pred = path[i, :] == 1
        if pred.any():
            E[pred, i] = _ols(y[:, i], y[:, pred])
        succ = path[:, i] == 1
        if succ.any():
            E[succ, i] = np.corrcoef(np.column_stack((y[:, succ], y[:, i])), rowvar=False)[:-1, -1]
    return E


class _Block:
    """One latent variable's manifest variables during the iteration."""

    def __init__(self, values: np.ndarray, scales: list, dummies: list, mode):
        self.initial = values.copy()      # treated (and, for ORD / NOM, ranked) columns: what every quantification starts from
        self.current = values.copy()      # the quantified columns of the iteration in flight
        self.scales, self.dummies, self.mode = scales, dummies, mode
        self.observed = None if not np.isnan(values).any() else (~np.isnan(values)).astype(np.float64)
        self.weights = None


class HostNonmetricWeights:
    def __init__(self, treated: pd.DataFrame, config, correction: float, path: pd.DataFrame, dummies: dict):
        self.lvs = list(path)
        self.path = path.loc[self.lvs, self.lvs].to_numpy(dtype=np.float64)
        self.correction = correction
        self.index = treated.index
        self.blocks, self.mvs = [], []
        n = treated.shape[0]
        self.scores = np.zeros((n, len(self.lvs)))
        for i, lv in enumerate(self.lvs):
            mvs = list(config.mvs(lv))
            self.mvs.extend(mvs)
            blk = _Block(treated.loc[:, mvs].to_numpy(dtype=np.float64), [config.scale(mv) for mv in mvs],
                         [dummies.get(mv) for mv in mvs], config.mode(lv))
            k = len(mvs)
            w0 = np.full(k, 1.0 / np.sqrt(k))
            if blk.observed is None:
                self.scores[:, i] = blk.current @ w0
            else:  # weights.py:87-94: per-row rescaling by the observed part of the weight vector
                denom = ((w0 * blk.observed) ** 2).sum(axis=1)
                if (denom == 0).any():
                    raise ValueError("All mvs for lv " + lv + " in row " + str(int(np.argmax(denom == 0))) + " are NaN.")
                self.scores[:, i] = np.nansum(blk.current * w0, axis=1) / denom
            self.blocks.append(blk)

    # ---- quantification of one manifest variable (scale.py) -------------------------------------------------
    def _mode_b_target(self, blk: _Block, j: int, z: np.ndarray, betas: dict, key: int) -> np.ndarray:
        """scale.py:73 / weights.py:135-146: in a Mode-B block the column is quantified against the part of the inner
        estimate the OTHER columns of the block do not explain."""
        if blk.mode != Mode.B or blk.current.shape[1] == 1:
            return z
        if key not in betas:
            design = np.column_stack((np.ones(len(z)), blk.current))
            betas[key] = _ols(z, design)[1:]
        b = betas[key]
        others = np.delete(blk.current, j, axis=1) @ np.delete(b, j)
        return (z - others) / b[j]

    @staticmethod
    def _category_means(dummies: np.ndarray, z: np.ndarray) -> np.ndarray:
        return (dummies * z[:, None]).sum(axis=0) / dummies.sum(axis=0)

    def _monotone(self, start: np.ndarray, dummies: np.ndarray, z: np.ndarray, sign: int):
        """scale.py:53-66: pool adjacent categories until their means are monotone (increasing for sign = +1)."""
        scaling = start
        while True:
            width = dummies.shape[1]
            for c in range(width - 1):
                if np.sign(scaling[c] - scaling[c + 1]) == sign:
                    dummies[:, c + 1] += dummies[:, c]
                    dummies = np.delete(dummies, c, axis=1)
                    scaling = self._category_means(dummies, z)
                    break
            if dummies.shape[1] == 1 or dummies.shape[1] == width:
                break
        x = dummies @ scaling
        return x, np.var(x)

    def _quantify(self, blk: _Block, j: int, z: np.ndarray, betas: dict, key: int) -> np.ndarray:
        scale, col = blk.scales[j], blk.initial[:, j]
        if scale == Scale.RAW:
            return col
        if scale == Scale.NUM:  # scale.py:28-31
            finite = np.isfinite(col).sum()
            return util.treat_numpy(col) * np.sqrt(finite / (finite - 1))
        z = self._mode_b_target(blk, j, z, betas, key)
        means = util.groupby_mean(np.array([col, z]))[1]
        if scale == Scale.NOM:  # scale.py:83-88
            return util.treat_numpy(blk.dummies[j] @ means) * self.correction
        up, var_up = self._monotone(means, blk.dummies[j].astype(np.float64).copy(), z, 1)
        down, var_down = self._monotone(means, blk.dummies[j].astype(np.float64).copy(), z, -1)
        x = -down if var_up < var_down else up  # scale.py:77
        return util.treat_numpy(x) * self.correction

    # ---- one outer iteration (weights.py:108-120) --------------------------------------------------------------
    def iterate(self, scheme) -> float:
        betas = {}
        before = self.scores.copy()
        Z = self.scores @ inner_weights(scheme.value.letter, self.path, self.scores)
        for i, blk in enumerate(self.blocks):
            z = Z[:, i]
            for j in range(blk.current.shape[1]):
                blk.current[:, j] = self._quantify(blk, j, z, betas, i)
            if blk.mode == Mode.A:  # mode.py:31-42
                if blk.observed is not None:
                    w = np.nansum(blk.current * z[:, None], axis=0) / ((blk.observed * z[:, None]) ** 2).sum(axis=0)
                    y = np.nansum(blk.current * w, axis=1) / ((blk.observed * w) ** 2).sum(axis=1)
                else:
                    w = blk.current.T @ z / (z ** 2).sum()
                    y = blk.current @ w
            else:  # mode.py:54-61
                if blk.observed is not None:
                    raise Exception("Missing nonmetric data is not supported in mode B. LV with missing data: " + self.lvs[i])
                w = np.linalg.lstsq(blk.current, z, rcond=None)[0]
                y = blk.current @ w
            blk.weights = w
            self.scores[:, i] = util.treat_numpy(y) * self.correction
        return float(((np.abs(before) - np.abs(self.scores)) ** 2).sum())

    def calculate(self):
        """weights.py:122-133: quantified data, scores, weights rescaled so that the weighted block sum has unit
        (population) variance."""
        P, L = len(self.mvs), len(self.lvs)
        W = np.zeros((P, L))
        data_new = np.zeros((len(self.index), P))
        o = 0
        for i, blk in enumerate(self.blocks):
            k = blk.current.shape[1]
            W[o:o + k, i] = blk.weights
            data_new[:, o:o + k] = blk.current
            o += k
        composite = pd.DataFrame(data_new).dot(pd.DataFrame(W))  # (pandas dot: a NaN in a row makes the composite NaN, skipped by std)
        factors = 1.0 / (composite.std(axis=0, skipna=True).to_numpy() / self.correction)
        weights = (W * factors[None, :]).sum(axis=1)
        return (pd.DataFrame(data_new, index=self.index, columns=self.mvs),
                pd.DataFrame(self.scores.copy(), index=self.index, columns=self.lvs),
                pd.DataFrame(weights, index=self.mvs, columns=["weight"]))


def treat_nonmetric(config, data: pd.DataFrame):
    """config.py:306-319: standardise every column (population sd), replace ORD / NOM columns by the ranks of their
    distinct values and build their indicator matrices.  Returns (treated frame, {mv: dummies})."""
    kinds = {mv: config.scale(mv) for mv in data.columns}
    if None in kinds.values():
        raise TypeError("If you supply a scale for any MV, you must either supply a scale for all of them or specify a default scale.")
    if set(kinds.values()) == {Scale.RAW, Scale.NUM}:
        kinds = dict.fromkeys(kinds, Scale.NUM)
    n = data.shape[0]
    treated = (util.treat(data) / np.sqrt((n - 1) / n)).astype(np.float64)
    dummies = {}
    for mv, kind in kinds.items():
        if kind in (Scale.ORD, Scale.NOM):
            if data[mv].isnull().any():
                raise NotImplementedError("ordinal / nominal manifest variables with missing values are not supported: " + mv)
            treated[mv] = util.rank(treated[mv])
            dummies[mv] = util.dummy(treated[mv]).to_numpy(dtype=np.float64)
    return treated, dummies, kinds


class _HostScaleView:
    """What HostNonmetricWeights asks a Config for, with the RAW + NUM -> NUM promotion of config.py:311-313 applied
    (without mutating the caller's Config)."""

    def __init__(self, config, kinds):
        self._config, self._kinds = config, kinds

    def mvs(self, lv):
        return self._config.mvs(lv)

    def mode(self, lv):
        return self._config.mode(lv)

    def scale(self, mv):
        return self._kinds[mv]


class _RowLayout:
    """Layout of a bootstrap row, [weights P | r_squared L | total effects E | direct effects E | loadings P], with the
    effect pairs in the order of InnerModel.effects() restricted to structurally reachable pairs -- what
    plspm_b200.engine.Model reports for a device model (csrc/plspm_model.cpp), computed here without a device."""

    def __init__(self, P: int, path: np.ndarray):
        L = path.shape[0]
        reach = path.astype(bool)
        for _ in range(L):
            reach = reach | ((reach.astype(np.int64) @ reach.astype(np.int64)) > 0)
        pairs = [(f, t) for f in range(L) for t in range(L) if f != t and reach[t, f]]
        self.P, self.L, self.n_effects = P, L, len(pairs)
        self.effects_from = np.array([f for f, _ in pairs], dtype=np.int32)
        self.effects_to = np.array([t for _, t in pairs], dtype=np.int32)
        self.n_out = 2 * P + L + 2 * len(pairs)

    def split_row(self, rows: np.ndarray):
        P, L, E = self.P, self.L, self.n_effects
        return (rows[..., :P], rows[..., P:P + L], rows[..., P + L:P + L + E], rows[..., P + L + E:P + L + 2 * E],
                rows[..., P + L + 2 * E:])


class HostNonmetricSession:
    """fit / bootstrap / close with the interface of EngineSession, computed on the host (module docstring)."""
    host = True

    def __init__(self, config, data: pd.DataFrame, path: pd.DataFrame = None, correction: float = None):
        path = config.path() if path is None else path
        self.config, self.path = config, path
        self.lvs = list(path)
        self.blocks = {lv: list(config.mvs(lv)) for lv in self.lvs}
        self.mvs = [mv for lv in self.lvs for mv in self.blocks[lv]]
        self.index = data.index
        self.frame = data.loc[:, self.mvs]
        self.N = self.frame.shape[0]
        self.numeric, self.missing = True, bool(self.frame.isnull().values.any())
        self.correction = correction
        self.model = _RowLayout(len(self.mvs), path.loc[self.lvs, self.lvs].to_numpy(dtype=np.int64))

    def _estimate(self, frame: pd.DataFrame, scheme, tol: float, iterations: int):
        n = frame.shape[0]
        correction = np.sqrt(n / (n - 1))
        treated, dummies, kinds = treat_nonmetric(self.config, frame)
        calc = HostNonmetricWeights(treated, _HostScaleView(self.config, kinds), correction, self.path, dummies)
        it = 0
        while True:  # weights.py:179-186
            it += 1
            conv = calc.iterate(scheme)
            if conv < tol or it > iterations:
                break
        status = 1 if it > iterations else 0
        data_new, scores, weights = calc.calculate()
        return data_new, scores, weights, it, status

    @staticmethod
    def _results(data_new, scores, weights, path, lvs, blocks, it, status):
        from plspm.inner_model import InnerModel
        cross = scores.apply(lambda s: data_new.corrwith(s))  # outer_model.py:27
        P = data_new.shape[1]
        load = np.zeros(P)
        o = 0
        for li, lv in enumerate(lvs):
            k = len(blocks[lv])
            load[o:o + k] = cross.to_numpy()[o:o + k, li]
            o += k
        inner = InnerModel(path, scores)
        B = inner.path_coefficients().loc[lvs, lvs].to_numpy(dtype=np.float64)
        total = np.zeros_like(B)
        eff = inner.effects()
        for f, t, v in zip(eff["from"], eff["to"], eff["total"]):
            total[lvs.index(t), lvs.index(f)] = v
        return dict(weights=weights["weight"].to_numpy(), loadings=load, crossloadings=cross.to_numpy(),
                    r_squared=inner.r_squared().loc[lvs].to_numpy(), path_coefficients=B, total_effects=total,
                    scores=scores.to_numpy(), iterations=it, status=status, final_data=data_new)

    def fit(self, scheme, tol: float, iterations: int, want_scores: bool = True):
        data_new, scores, weights, it, status = self._estimate(self.frame, scheme, tol, iterations)
        if status != 0:  # weights.py:185-186
            raise Exception("Could not converge after " + str(it) + " iterations")
        return self._results(data_new, scores, weights, self.path, self.lvs, self.blocks, it, status)

    def bootstrap(self, scheme, tol: float, iterations: int, rep_begin: int, rep_count: int, seed: int = 0, idx=None,
                  out_device_ptr: int = 0):
        """bootstrap.py:54-68 as written: one full host fit per resample; failures are dropped by the caller."""
        assert not out_device_ptr
        rows = np.zeros((rep_count, self.model.n_out))
        status, iters = np.ones(rep_count, dtype=np.int32), np.zeros(rep_count, dtype=np.int32)
        ef, et = self.model.effects_from, self.model.effects_to
        for b in range(rep_count):
            pick = idx[b] if idx is not None else _philox_indices(seed, rep_begin + b, self.N)
            try:
                frame = self.frame.iloc[pick, :].reset_index(drop=True)
                data_new, scores, weights, it, st = self._estimate(frame, scheme, tol, iterations)
                if st != 0:
                    status[b], iters[b] = st, it
                    continue
                r = self._results(data_new, scores, weights, self.path, self.lvs, self.blocks, it, st)
            except NotImplementedError:
                raise
            except Exception:
                continue
            rows[b] = np.concatenate([r["weights"], r["r_squared"], r["total_effects"][et, ef],
                                      r["path_coefficients"][et, ef], r["loadings"]])
            status[b], iters[b] = 0, it
        return rows, status, iters

    def close(self, trim_pool: bool = True):
        pass


def _philox_indices(seed: int, replicate: int, n: int) -> np.ndarray:
    """The library's resample stream (Philox4x32-10 keyed by (seed, global replicate id)), so host-path bootstraps
    draw the same resamples as device-path ones; computed by the library on the device."""
    from plspm_b200 import engine
    return engine.resample_indices(seed, replicate, n)
