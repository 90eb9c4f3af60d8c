"""Randomised sweep of the solver code shared with the CUDA build (csrc/solver_core.h, solver_num.h compiled for
the host by tests/emul) against the oracle: random DAGs, ragged block sizes (padding columns, multi-slot
blocks), mixed Mode A / B, every scheme, reverse-coded blocks, both tile policies, resampled rows.  CPU only."""
import numpy as np
import pytest

from oracle import plspm_oracle as orc
from oracle import plspm_oracle_nonmetric as onm
from tests.emul import emul

SCHEMES = ("centroid", "factorial", "path")


def random_model(rng, L, kmax):
    """Random lower-triangular DAG (every LV but the first has a predecessor), blocks of 1..kmax MVs, and data
    generated along the graph (some blocks reverse coded so that the sign vote matters)."""
    path = np.zeros((L, L), dtype=np.int8)
    for i in range(1, L):
        preds = rng.choice(i, size=rng.integers(1, min(i, 3) + 1), replace=False)
        path[i, preds] = 1
    sizes = [int(rng.integers(1, kmax + 1)) for _ in range(L)]
    if sum(sizes) < L + 2:
        sizes[0] += 2
    N = int(rng.integers(120, 400))
    eta = np.zeros((N, L))
    for i in range(L):
        e = rng.normal(size=N)
        pred = np.flatnonzero(path[i])
        eta[:, i] = (eta[:, pred] @ rng.uniform(0.3, 0.7, size=len(pred)) if len(pred) else 0.0) + 0.7 * e
        eta[:, i] = (eta[:, i] - eta[:, i].mean()) / eta[:, i].std()
    cols = []
    for i, k in enumerate(sizes):
        lam = rng.uniform(0.6, 0.9, size=k) * (-1.0 if rng.random() < 0.25 else 1.0)
        cols.append(eta[:, [i]] * lam + rng.normal(size=(N, k)) * np.sqrt(1 - lam ** 2))
    X = np.column_stack(cols) * rng.uniform(0.5, 20.0, size=sum(sizes)) + rng.normal(size=sum(sizes)) * 3
    modes = [int(rng.random() < 0.4) if sizes[i] > 1 else 0 for i in range(L)]
    return X, path, sizes, modes


@pytest.mark.parametrize("seed", range(72))
def test_metric_solver_on_random_models(seed):
    rng = np.random.default_rng(1000 + seed)
    X, path, sizes, modes = random_model(rng, L=int(rng.integers(2, 9)), kmax=int(rng.choice([3, 8, 11, 19])))
    scheme, scaled, policy = SCHEMES[seed % 3], bool(seed & 1), (0, 1, 2)[(seed // 3) % 3]
    idx = rng.integers(0, X.shape[0], size=X.shape[0]).astype(np.int32) if seed % 4 == 3 else None
    Xo = X if idx is None else X[idx]
    try:
        o = orc.fit(Xo, sizes, modes, path, scheme, scaled)
    except (orc.NotConverged, np.linalg.LinAlgError):
        pytest.skip("degenerate random model")
    r = emul.fit(X, sizes, modes, path, scheme, scaled, idx=idx, tile_policy=policy)
    assert r["status"] == 0 and r["iterations"] == o["iterations"]
    rel = 1e-7 if any(modes) else 1e-9
    np.testing.assert_allclose(r["weights"], o["weights"], rtol=rel, atol=1e-12)
    np.testing.assert_allclose(r["path_coefficients"], o["path_coefficients"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["r_squared"], o["r_squared"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["loadings"], o["loadings"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["total_effects"], o["total_effects"], rtol=rel, atol=1e-10)
    if idx is None:
        np.testing.assert_allclose(r["scores"], o["scores"], rtol=rel, atol=1e-9)


@pytest.mark.parametrize("seed", range(36))
def test_numeric_nonmetric_solver_on_random_models(seed):
    rng = np.random.default_rng(5000 + seed)
    X, path, sizes, modes = random_model(rng, L=int(rng.integers(2, 7)), kmax=int(rng.choice([3, 8, 12])))
    scheme, policy = SCHEMES[seed % 3], (0, 1, 2)[(seed // 3) % 3]
    idx = rng.integers(0, X.shape[0], size=X.shape[0]).astype(np.int32) if seed % 3 == 2 else None
    Xo = X if idx is None else X[idx]
    try:
        o = onm.fit_num(Xo, sizes, modes, path, scheme)
    except (orc.NotConverged, np.linalg.LinAlgError):
        pytest.skip("degenerate random model")
    r = emul.fit_num(X, sizes, modes, path, scheme, idx=idx, tile_policy=policy)
    assert r["status"] == 0 and r["iterations"] == o["iterations"]
    rel = 1e-7 if any(modes) else 1e-9
    np.testing.assert_allclose(r["weights"], o["weights"], rtol=rel, atol=1e-12)
    np.testing.assert_allclose(r["path_coefficients"], o["path_coefficients"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["r_squared"], o["r_squared"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["loadings"], o["loadings"], rtol=rel, atol=1e-10)
