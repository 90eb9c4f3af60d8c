"""CPU oracle for the PLS-PM weight-estimation hot path  --  TEST INFRASTRUCTURE ONLY.

A dense float64 NumPy restatement of the reference algorithm
(GoogleCloudPlatform/plspm-python @ 37f4aaf, v0.5.7).  Each function cites the
reference file:line it follows.  It keeps the reference's *dataflow* (Y = X.W,
standardise, inner weights, Z = Y.E, per-block outer weights, convergence on
the weights) and is therefore independent of the covariance-domain formulation
the CUDA engine uses.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.

Parity pin: tests/test_oracle.py checks this oracle against
  (1) the R-generated golden CSV values the reference's own tests use
      (tests/golden/satisfaction.npz "R/..." keys, from
      /root/reference/tests/data/satisfaction.*.csv), and
  (2) outputs of the reference itself run in the build container
      (tests/golden/*.npz "ref/...", "boot/..." keys; tests/golden/make_golden.py).

Third-party arithmetic restated here (not under /root/reference):
  * statsmodels (unpinned; requirements.txt:4) OLS == least squares via pinv
    (scheme.py:50, inner_model.py:76-77)  -> numpy.linalg.lstsq / pinv;
  * scipy.linalg.lstsq (mode.py:51)        -> numpy.linalg.lstsq (same LAPACK gelsd).

Deliberate differences from the reference (results identical):
  * Q1: the reference runs the whole iteration twice per fit
    (estimator.py:39,52); the oracle runs it once.
  * weights.py:61 builds the full (P+L)^2 correlation matrix and keeps the
    P x L corner; the oracle computes that corner directly.
"""
from __future__ import annotations

import numpy as np

SCHEME_CENTROID, SCHEME_FACTORIAL, SCHEME_PATH = 0, 1, 2
MODE_A, MODE_B = 0, 1
STATUS_OK, STATUS_NOT_CONVERGED, STATUS_SINGULAR = 0, 1, 2

_SCHEME_IDS = {"centroid": 0, "factorial": 1, "path": 2, 0: 0, 1: 1, 2: 2}


class NotConverged(Exception):
    pass


# ----------------------------------------------------------------------------------------------
# data treatment
# ----------------------------------------------------------------------------------------------
def impute(X: np.ndarray) -> np.ndarray:
    """Column-mean imputation of NaNs (reference util.py:61-68)."""
    if not np.isnan(X).any():
        return X
    X = X.copy()
    means = np.nanmean(X, axis=0)
    r, cidx = np.where(np.isnan(X))
    X[r, cidx] = means[cidx]
    return X


def treat_metric(X: np.ndarray, scaled: bool) -> np.ndarray:
    """Metric branch of Config.treat (reference config.py:299-305 -> util.py:33-39).

    Centre every column; if `scaled`, divide ALL columns by one pooled scalar
    sd(all N*P values, ddof=1) * sqrt((N-1)/N)   (quirk Q2).
    """
    X = impute(np.asarray(X, dtype=np.float64))
    n = X.shape[0]
    Xc = X - X.mean(axis=0)
    if scaled:
        pooled = X.reshape(-1).std(ddof=1) * np.sqrt((n - 1) / n)
        Xc = Xc / pooled
    return Xc


def outer_design(block_sizes) -> np.ndarray:
    """P x L 0/1 outer design matrix (reference config.py:140-144, util.py:71-77)."""
    block_sizes = np.asarray(block_sizes, dtype=np.int64)
    P, L = int(block_sizes.sum()), len(block_sizes)
    odm = np.zeros((P, L))
    off = 0
    for l, k in enumerate(block_sizes):
        odm[off:off + k, l] = 1.0
        off += k
    return odm


# ----------------------------------------------------------------------------------------------
# inner-weight schemes (reference scheme.py)
# ----------------------------------------------------------------------------------------------
def _ols(y: np.ndarray, X: np.ndarray) -> np.ndarray:
    """statsmodels OLS(...).fit().params == pinv(X) @ y."""
    return np.linalg.lstsq(X, y, rcond=None)[0]


def inner_weights(scheme: int, path: np.ndarray, Y: np.ndarray) -> np.ndarray:
    C = (path + path.T).astype(np.float64)
    if scheme == SCHEME_CENTROID:  # scheme.py:27-28
        return np.sign(np.corrcoef(Y, rowvar=False) * C)
    if scheme == SCHEME_FACTORIAL:  # scheme.py:36-37 (cov, ddof=1: quirk Q3)
        return np.cov(Y, rowvar=False) * C
    if scheme == SCHEME_PATH:  # scheme.py:45-54
        E = path.astype(np.float64).copy()
        R = np.corrcoef(Y, rowvar=False)
        for i in range(E.shape[0]):
            follow = path[i, :] == 1
            if follow.any():
                E[follow, i] = _ols(Y[:, i], Y[:, follow])
            predec = path[:, i] == 1
            if predec.any():
                E[predec, i] = R[predec, i]
        return E
    raise ValueError("unknown scheme")


# ----------------------------------------------------------------------------------------------
# the iteration (reference weights.py:26-70, 172-187; mode.py:28-29, 50-52)
# ----------------------------------------------------------------------------------------------
def estimate_weights(Xc: np.ndarray, block_sizes, modes, path, scheme, tol=1e-6, max_iter=100):
    """WeightsCalculatorFactory.calculate on treated data.

    Returns dict(weights [P], scores [N, L], iterations, crossloadings_sign_votes [L], W_final [P, L]).
    Raises NotConverged like weights.py:185-186.
    """
    scheme = _SCHEME_IDS[scheme]
    block_sizes = np.asarray(block_sizes, dtype=np.int64)
    modes = np.asarray(modes, dtype=np.int64)
    path = np.asarray(path, dtype=np.int64)
    N, P = Xc.shape
    L = len(block_sizes)
    offs = np.concatenate(([0], np.cumsum(block_sizes)))
    odm = outer_design(block_sizes)
    correction = np.sqrt(N / (N - 1))

    # weights.py:28-34  initial weights
    wf = correction / (Xc @ odm).std(axis=0, ddof=1)
    W = odm * wf
    w_old = W.sum(axis=1)

    iteration = 0
    while True:  # weights.py:179-184
        iteration += 1
        # weights.py:43-44
        Y = Xc @ W
        Y = (Y - Y.mean(axis=0)) / Y.std(axis=0, ddof=1) / correction
        E = inner_weights(scheme, path, Y)  # weights.py:45
        Z = Y @ E  # weights.py:46
        for l in range(L):  # weights.py:47-50
            blk = slice(offs[l], offs[l + 1])
            if modes[l] == MODE_A:  # mode.py:28-29
                W[blk, l] = Xc[:, blk].T @ Z[:, l] / N
            else:  # mode.py:50-52
                W[blk, l] = np.linalg.lstsq(Xc[:, blk], Z[:, l], rcond=None)[0]
        w_new = W.sum(axis=1)
        conv = float(((np.abs(w_old) - np.abs(w_new)) ** 2).sum())  # weights.py:51-52
        w_old = w_new
        if (conv < tol) or (iteration > max_iter):
            break
    if iteration > max_iter:  # weights.py:185-186 (quirk Q4)
        raise NotConverged("Could not converge after " + str(iteration) + " iterations")

    # weights.py:56-70
    wf = 1.0 / ((Xc @ W).std(axis=0, ddof=1) / correction)
    Wf = W * wf
    S = Xc @ Wf
    Xs = Xc - Xc.mean(axis=0)
    Ss = S - S.mean(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        cor = (Xs.T @ Ss) / np.sqrt(np.outer((Xs ** 2).sum(axis=0), (Ss ** 2).sum(axis=0)))
    odm_final = (Wf != 0).astype(np.float64)
    votes = np.copysign(1.0, cor * odm_final).sum(axis=0)  # quirk Q6: all P rows vote
    w_sign = np.copysign(1.0, votes)
    S = S * w_sign
    return dict(weights=Wf.sum(axis=1), scores=S, iterations=iteration, signs=w_sign, crossloadings=cor * w_sign)


# ----------------------------------------------------------------------------------------------
# inner model on the scores (reference inner_model.py:33-61, 66-83)
# ----------------------------------------------------------------------------------------------
def inner_model(path: np.ndarray, S: np.ndarray):
    path = np.asarray(path, dtype=np.int64)
    N, L = S.shape
    B = np.zeros((L, L))
    r2 = np.zeros(L)
    for i in range(L):
        pred = np.where(path[i, :] == 1)[0]
        if len(pred) == 0:
            continue
        A = np.column_stack((np.ones(N), S[:, pred]))
        beta = _ols(S[:, i], A)
        resid = S[:, i] - A @ beta
        B[i, pred] = beta[1:]
        r2[i] = 1.0 - float(resid @ resid) / float(((S[:, i] - S[:, i].mean()) ** 2).sum())
    # _effects, inner_model.py:33-49
    if L == 2:
        indirect = np.zeros((L, L))
        total = B.copy()
    else:
        indirect = np.zeros((L, L))
        Pk = B.copy()
        for _ in range(1, L):
            Pk = Pk @ B
            indirect = indirect + Pk
        total = B + indirect
    return dict(path_coefficients=B, r_squared=r2, indirect_effects=indirect, total_effects=total)


def fit(X, block_sizes, modes, path, scheme="centroid", scaled=True, tol=1e-6, max_iter=100):
    """Full single fit: Estimator.estimate (estimator.py:29-55, non-HOC) + InnerModel + loadings.

    X: raw [N, P] float64, columns grouped by LV in path order.  Returns everything in
    path-LV / ODM order.
    """
    Xc = treat_metric(X, scaled)
    out = estimate_weights(Xc, block_sizes, modes, path, scheme, tol, max_iter)
    out.update(inner_model(path, out["scores"]))
    odm = outer_design(block_sizes)
    out["loadings"] = (out["crossloadings"] * odm).sum(axis=1)  # outer_model.py:26-27 / bootstrap.py:65-66
    out["status"] = STATUS_OK
    return out


# ----------------------------------------------------------------------------------------------
# bootstrap (reference bootstrap.py:45-75, 24-32)
# ----------------------------------------------------------------------------------------------
def effect_pairs(path: np.ndarray):
    """(from, to) pairs in the order _effects emits rows (inner_model.py:50-60), restricted to
    structurally reachable pairs (total effect can be non-zero)."""
    path = np.asarray(path, dtype=np.int64)
    L = path.shape[0]
    reach = path.astype(bool).copy()
    for _ in range(L):
        reach = reach | ((reach.astype(np.int64) @ reach.astype(np.int64)) > 0)
    return [(f, t) for f in range(L) for t in range(L) if f != t and reach[t, f]]


def replicate_row(X, idx, block_sizes, modes, path, scheme, scaled, tol=1e-6, max_iter=100):
    """One bootstrap replicate (bootstrap.py:56-66).  Returns (row, iterations, status);
    row = [weights P | r_squared L | total effects n_eff | direct effects n_eff | loadings P]."""
    P, L = int(np.sum(block_sizes)), len(block_sizes)
    pairs = effect_pairs(path)
    n_out = 2 * P + L + 2 * len(pairs)
    try:
        r = fit(X[idx, :], block_sizes, modes, path, scheme, scaled, tol, max_iter)
    except NotConverged:
        return np.full(n_out, np.nan), max_iter + 1, STATUS_NOT_CONVERGED
    except np.linalg.LinAlgError:
        return np.full(n_out, np.nan), 0, STATUS_SINGULAR
    tot = np.array([r["total_effects"][t, f] for f, t in pairs])
    direct = np.array([r["path_coefficients"][t, f] for f, t in pairs])
    row = np.concatenate((r["weights"], r["r_squared"], tot, direct, r["loadings"]))
    return row, r["iterations"], STATUS_OK


def bootstrap(X, indices, block_sizes, modes, path, scheme="centroid", scaled=True, tol=1e-6, max_iter=100):
    """Replicates for an injected index matrix [B, N]. Returns (out [B, n_out], iters [B], status [B])."""
    rows, iters, status = [], [], []
    for b in range(indices.shape[0]):
        row, it, st = replicate_row(X, indices[b], block_sizes, modes, path, scheme, scaled, tol, max_iter)
        rows.append(row)
        iters.append(it)
        status.append(st)
    return np.array(rows), np.array(iters, dtype=np.int32), np.array(status, dtype=np.int32)


def summary(samples: np.ndarray, original: np.ndarray):
    """_create_summary (bootstrap.py:24-32): columns original, mean, std.error, perc.025, perc.975, t stat."""
    sd = samples.std(axis=0, ddof=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = original / sd
    return np.column_stack((original, samples.mean(axis=0), sd, np.quantile(samples, 0.025, axis=0),
                            np.quantile(samples, 0.975, axis=0), t))


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 resample indices: the engine's counter-based generator, restated
# (D. E. Shaw Research Random123 "philox4x32-10"; known-answer vectors in tests/test_oracle.py).
# counter = (row_group, 0, replicate_lo, replicate_hi), key = (seed_lo, seed_hi); the four 32-bit
# outputs give rows 4*row_group .. 4*row_group+3;  index = (u32 * N) >> 32.
# ----------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _MASK for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & _MASK, np.uint64(k1) & _MASK
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox_indices(seed: int, replicate: int, N: int) -> np.ndarray:
    """int32 [N] resample indices of global replicate id `replicate`."""
    groups = np.arange((N + 3) // 4, dtype=np.uint64)
    z = np.zeros_like(groups)
    r = philox4x32(groups, z, z + np.uint64(replicate & 0xFFFFFFFF), z + np.uint64((replicate >> 32) & 0xFFFFFFFF),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.stack(r, axis=1).reshape(-1)[:N]
    return ((u * np.uint64(N)) >> np.uint64(32)).astype(np.int32)
