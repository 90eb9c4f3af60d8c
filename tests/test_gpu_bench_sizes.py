"""Parity at the sizes bench.py measures, on the routes it measures (BASELINE.json configs 3, 4, 5 and the
north-star headline config): one full default-size bootstrap batch PLUS a tail batch, replicates picked at the
start, the end and the batch boundary, against the CPU oracle at 1e-6 with identical iteration counts; the
stage profile must show that the tensor-core routes (integer Gram, fused sign vote) produced the result.
Reference path: weights.py:172-187 under bootstrap.py:54-66."""
import numpy as np
import pytest

from oracle import plspm_oracle as orc
from plspm_b200.synth import make_synthetic

pytestmark = pytest.mark.gpu
REL = 1e-6


@pytest.fixture(scope="module")
def eng():
    from plspm_b200 import engine
    engine.load()
    assert engine.device_count() > 0, "no CUDA device"
    engine.set_device(0)
    return engine


@pytest.fixture(scope="module")
def c3(eng):
    X, path = make_synthetic(100_000, 32, 8, seed=0)  # bench.py's c3 / c4 data
    return X, path


def _bootstrap_and_check(eng, X, path, blocks, modes, scheme, reps, picks, seed=0, begin=0):
    L = len(blocks)
    model = eng.Model(blocks, modes, path, True)
    data = eng.Data(model, X)
    assert not model.full_tiles
    before = eng.redo_count()
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, scheme, begin, reps, seed=seed)
    prof = eng.profile_get()
    assert (status == 0).all()
    # the routes bench.py times: tensor-core second moments and the tensor-core sign vote, nothing redone
    assert prof["gram_i8"][1] > 0 and prof["gram"][1] == 0, prof
    assert prof["cross"][1] > 0, prof
    assert eng.redo_count() == before
    N = X.shape[0]
    for b in picks:
        idx = orc.philox_indices(seed, begin + b, N)
        ref, it, st = orc.replicate_row(X, idx, blocks, modes, path, scheme, True)
        assert st == 0 and iters[b] == it, (b, iters[b], it)
        np.testing.assert_allclose(rows[b], ref, rtol=REL, atol=1e-9, err_msg="replicate %d" % b)
    return rows, iters


def test_c3_headline_bootstrap_full_batch_plus_tail(eng, c3):
    """North-star headline config: N=100k, 32 LVs x 8 MVs, Mode A, centroid.  1184 + 100 replicates."""
    X, path = c3
    picks = [0, 1, 2, 300, 591, 777, 1000, 1182, 1183, 1184, 1185, 1200, 1250, 1281, 1282, 1283]
    _bootstrap_and_check(eng, X, path, [8] * 32, [0] * 32, "centroid", 1284, picks, seed=0, begin=5000)


def test_c3_single_fit_factorial(eng, c3):
    """BASELINE config 3: single fit, N=100k, Mode A, factorial scheme (the hot loop run once)."""
    X, path = c3
    model = eng.Model([8] * 32, [0] * 32, path, True)
    data = eng.Data(model, X)
    got = eng.fit(model, data, "factorial")
    ref = orc.fit(X, [8] * 32, [0] * 32, path, "factorial", True)
    assert got["status"] == 0 and got["iterations"] == ref["iterations"]
    for key, atol in (("weights", 0), ("scores", 1e-8), ("path_coefficients", 1e-9), ("total_effects", 1e-9),
                      ("r_squared", 1e-9), ("loadings", 1e-9), ("crossloadings", 1e-8)):
        np.testing.assert_allclose(got[key], ref[key], rtol=REL, atol=atol, err_msg=key)


def test_c4_mode_b_path_bootstrap_full_batch_plus_tail(eng, c3):
    """BASELINE config 4: N=100k, Mode B, path scheme."""
    X, path = c3
    picks = [0, 1, 590, 1183, 1184, 1185, 1250, 1283]
    _bootstrap_and_check(eng, X, path, [8] * 32, [1] * 32, "path", 1284, picks, seed=0, begin=77)


def test_c5_shaped_bootstrap(eng, monkeypatch):
    """BASELINE config 5's model (64 LVs x 16 MVs, P = 1024) at N = 60k: several 256-column chunks of the sign
    vote, blocks of two slots, a full batch plus a tail batch."""
    X, path = make_synthetic(60_000, 64, 16, seed=0)
    monkeypatch.setenv("PLSPM_MAX_BATCH", "128")
    picks = [0, 64, 127, 128, 139]
    _bootstrap_and_check(eng, X, path, [16] * 64, [0] * 64, "centroid", 140, picks, seed=3, begin=10)
