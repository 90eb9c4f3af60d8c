"""Synthetic PLS-PM data generator used by the parity tests and bench.py.

This is the generator SURVEY.md §8(d) specifies for configs C3-C5: a
lower-triangular "chain-of-two" DAG (path[j, j-1] = 1 for j >= 1 and
path[j, j-2] = 1 for j >= 2), latent scores built from their predecessors and
re-standardised, K reflective manifest variables per latent variable with
loadings U(0.6, 0.9).  Columns are grouped by latent variable.
"""
import numpy as np


def chain_path(L: int) -> np.ndarray:
    """L x L int8 path matrix; path[i, j] = 1 means LV j -> LV i."""
    path = np.zeros((L, L), dtype=np.int8)
    for j in range(1, L):
        path[j, j - 1] = 1
        if j >= 2:
            path[j, j - 2] = 1
    return path


def make_synthetic(N: int, L: int, K: int, seed: int = 0, reverse_blocks=()):
    """Returns (X float64 [N, L*K], path int8 [L, L]).

    reverse_blocks: LV ids whose manifest variables are reverse coded (x -> -x);
    exercises the sign-vote step (reference weights.py:62-68).
    """
    rng = np.random.default_rng(seed)
    path = chain_path(L)
    eta = np.empty((N, L), dtype=np.float64)
    eta[:, 0] = rng.standard_normal(N)
    for j in range(1, L):
        preds = [j - 1] + ([j - 2] if j >= 2 else [])
        coef = 0.2 + 0.5 / np.sqrt(len(preds))
        v = coef * eta[:, preds].sum(axis=1) + 0.6 * rng.standard_normal(N)
        eta[:, j] = (v - v.mean()) / v.std()
    lam = rng.uniform(0.6, 0.9, size=(L, K))
    X = np.empty((N, L * K), dtype=np.float64)
    for j in range(L):
        noise = rng.standard_normal((N, K))
        X[:, j * K:(j + 1) * K] = eta[:, j:j + 1] * lam[j] + noise * np.sqrt(1.0 - lam[j] ** 2)
        if j in reverse_blocks:
            X[:, j * K:(j + 1) * K] *= -1.0
    return X, path
