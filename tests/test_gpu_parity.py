"""GPU parity tests: the CUDA path (through the C ABI, ctypes) vs the CPU oracle and the golden fixtures.
Tolerance: 1e-6 relative on outer weights, LV scores and path coefficients (BASELINE.json north_star);
iteration counts must be identical."""
import os

import numpy as np
import pytest

from oracle import plspm_oracle as orc
from plspm_b200.synth import make_synthetic

pytestmark = pytest.mark.gpu
REL = 1e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    from plspm_b200 import engine
    engine.load()
    assert engine.device_count() > 0, "no CUDA device"
    engine.set_device(0)
    return engine


def check_fit(got, ref, rel=REL):
    assert got["status"] == 0
    assert got["iterations"] == ref["iterations"]
    np.testing.assert_allclose(got["weights"], ref["weights"], rtol=rel)
    np.testing.assert_allclose(got["scores"], ref["scores"], rtol=rel, atol=1e-8)
    np.testing.assert_allclose(got["path_coefficients"], ref["path_coefficients"], rtol=rel, atol=1e-9)
    np.testing.assert_allclose(got["total_effects"], ref["total_effects"], rtol=rel, atol=1e-9)
    np.testing.assert_allclose(got["r_squared"], ref["r_squared"], rtol=rel, atol=1e-9)
    np.testing.assert_allclose(got["loadings"], ref["loadings"], rtol=rel, atol=1e-9)
    np.testing.assert_allclose(got["crossloadings"], ref["crossloadings"], rtol=rel, atol=1e-8)


@pytest.mark.parametrize("scheme", ("centroid", "factorial", "path"))
@pytest.mark.parametrize("mode", (0, 1))
@pytest.mark.parametrize("scaled", (False, True))
def test_satisfaction_vs_oracle_and_reference(eng, sat, scheme, mode, scaled):
    model = eng.Model(sat["block_sizes"], [mode] * 6, sat["path"], scaled)
    data = eng.Data(model, sat["X"])
    got = eng.fit(model, data, scheme)
    ref = orc.fit(sat["X"], sat["block_sizes"], [mode] * 6, sat["path"], scheme, scaled)
    check_fit(got, ref)
    tag = "ref/%s/%s/%s/" % (scheme, "AB"[mode], "scaled" if scaled else "unscaled")
    assert got["iterations"] == int(sat[tag + "iterations"])
    np.testing.assert_allclose(got["weights"], sat[tag + "weights"], rtol=REL)
    np.testing.assert_allclose(got["scores"], sat[tag + "scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(got["path_coefficients"], sat[tag + "path_coefficients"], rtol=REL, atol=1e-9)


def test_satisfaction_r_golden(eng, sat):
    model = eng.Model(sat["block_sizes"], [0] * 6, sat["path"], False)
    data = eng.Data(model, sat["X"])
    got = eng.fit(model, data, "centroid")
    assert got["iterations"] == 4
    np.testing.assert_allclose(got["scores"], sat["R/scores"], rtol=REL, atol=1e-8)
    np.testing.assert_allclose(got["weights"], sat["R/centroid/weight"], rtol=REL)
    np.testing.assert_allclose(got["loadings"], sat["R/centroid/loading"], rtol=REL)
    np.testing.assert_allclose(got["crossloadings"], sat["R/crossloadings"], rtol=REL, atol=1e-9)
    lvs = list(sat["lvs"])
    for f, t, d, tot in zip(sat["R/effects_from"], sat["R/effects_to"], sat["R/effects_direct"], sat["R/effects_total"]):
        i, j = lvs.index(t), lvs.index(f)
        np.testing.assert_allclose(got["path_coefficients"][i, j], d, rtol=REL, atol=1e-10)
        np.testing.assert_allclose(got["total_effects"][i, j], tot, rtol=REL, atol=1e-10)
    for scheme in ("path", "factorial"):
        g = eng.fit(model, data, scheme)
        np.testing.assert_allclose(g["weights"], sat["R/%s/weight" % scheme], rtol=REL)
        np.testing.assert_allclose(g["loadings"], sat["R/%s/loading" % scheme], rtol=REL)


@pytest.mark.parametrize("case", ("syn_a", "syn_b", "syn_c", "syn_d", "syn_e", "syn_f", "syn_g", "syn_h"))
def test_synthetic_cases(eng, syn, case):
    N, L, K, seed = (int(v) for v in syn[case + "/gen"])
    X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[case + "/reverse"]))
    mode = 0 if str(syn[case + "/mode"]) == "A" else 1
    scheme, scaled = str(syn[case + "/scheme"]), bool(syn[case + "/scaled"])
    model = eng.Model([K] * L, [mode] * L, path, scaled)
    data = eng.Data(model, X)
    got = eng.fit(model, data, scheme)
    check_fit(got, orc.fit(X, [K] * L, [mode] * L, path, scheme, scaled))
    assert got["iterations"] == int(syn[case + "/iterations"])
    np.testing.assert_allclose(got["weights"], syn[case + "/weights"], rtol=REL)
    np.testing.assert_allclose(got["scores"], syn[case + "/scores"], rtol=REL, atol=1e-8)


@pytest.mark.parametrize("case", ("centroid/A/unscaled", "path/B/scaled", "factorial/A/scaled"))
def test_bootstrap_injected_indices_vs_reference(eng, sat, case):
    scheme, mode, sc = case.split("/")
    m = 0 if mode == "A" else 1
    nrep = int(sat["boot/n_replicates"])
    idx = np.random.default_rng(1234).integers(0, 250, (1000, 250), dtype=np.int32)
    model = eng.Model(sat["block_sizes"], [m] * 6, sat["path"], sc == "scaled")
    data = eng.Data(model, sat["X"])
    rows, status, iters = eng.bootstrap(model, data, scheme, 0, 1000, idx=idx)
    assert (status == 0).all()
    tag = "boot/%s/" % case
    np.testing.assert_array_equal(iters[:nrep], sat[tag + "iterations"])
    w, r2, tot, direct, load = model.split_row(rows)
    rel = REL if mode == "A" else 1e-5  # Mode B resamples of 250 rows are ill-conditioned
    np.testing.assert_allclose(w[:nrep], sat[tag + "weights"], rtol=rel)
    np.testing.assert_allclose(r2[:nrep], sat[tag + "r_squared"], rtol=rel, atol=1e-9)
    np.testing.assert_allclose(load[:nrep], sat[tag + "loadings"], rtol=rel, atol=1e-9)
    pc = np.array([[sat[tag + "path_coefficients"][b, t, f] for f, t in zip(model.effects_from, model.effects_to)]
                   for b in range(nrep)])
    te = np.array([[sat[tag + "total_effects"][b, t, f] for f, t in zip(model.effects_from, model.effects_to)]
                   for b in range(nrep)])
    np.testing.assert_allclose(direct[:nrep], pc, rtol=rel, atol=1e-9)
    np.testing.assert_allclose(tot[:nrep], te, rtol=rel, atol=1e-9)
    # all 1000 replicates (config C2) against the oracle
    orows, oiters, ostatus = orc.bootstrap(sat["X"], idx, sat["block_sizes"], [m] * 6, sat["path"], scheme,
                                           sc == "scaled")
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=rel, atol=1e-9)


def test_philox_stream_matches_oracle(eng):
    for seed, rep, N in ((0, 0, 250), (7, 12345678901, 1001), (2**40 + 5, 3, 4099)):
        np.testing.assert_array_equal(eng.resample_indices(seed, rep, N), orc.philox_indices(seed, rep, N))


def test_bootstrap_generated_indices_match_injected(eng, sat):
    model = eng.Model(sat["block_sizes"], [0] * 6, sat["path"], True)
    data = eng.Data(model, sat["X"])
    rows, status, iters = eng.bootstrap(model, data, "centroid", 40, 24, seed=99)
    idx = np.stack([orc.philox_indices(99, 40 + b, 250) for b in range(24)])
    rows2, status2, iters2 = eng.bootstrap(model, data, "centroid", 0, 24, idx=idx)
    np.testing.assert_array_equal(rows, rows2)
    np.testing.assert_array_equal(iters, iters2)
    orows, oiters, _ = orc.bootstrap(sat["X"], idx, sat["block_sizes"], [0] * 6, sat["path"], "centroid", True)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


def test_ragged_blocks_mixed_modes(eng):
    rng = np.random.default_rng(3)
    sizes = [1, 9, 3, 17, 2]
    L = len(sizes)
    path = np.zeros((L, L), dtype=np.int8)
    path[1, 0] = path[2, 0] = path[2, 1] = path[3, 2] = path[4, 1] = path[4, 3] = 1
    eta = rng.standard_normal((3001, L))
    for i in range(1, L):
        eta[:, i] += eta[:, :i] @ (0.5 * path[i, :i])
    X = np.concatenate([eta[:, [l]] * rng.uniform(0.5, 1.0, (1, k)) + 0.7 * rng.standard_normal((3001, k))
                        for l, k in enumerate(sizes)], axis=1) * 3.0 + 10.0
    modes = [0, 1, 0, 0, 1]
    model = eng.Model(sizes, modes, path, True)
    data = eng.Data(model, X)
    for scheme in ("centroid", "factorial", "path"):
        check_fit(eng.fit(model, data, scheme), orc.fit(X, sizes, modes, path, scheme, True))
    idx = rng.integers(0, 3001, (5, 3001), dtype=np.int32)
    rows, status, iters = eng.bootstrap(model, data, "path", 0, 5, idx=idx)
    orows, oiters, _ = orc.bootstrap(X, idx, sizes, modes, path, "path", True)
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


def test_not_converged_and_edge_inputs(eng, sat):
    model = eng.Model(sat["block_sizes"], [1] * 6, sat["path"], True)
    data = eng.Data(model, sat["X"])
    got = eng.fit(model, data, "centroid", tol=1e-30, max_iter=3)
    assert got["status"] == eng.STATUS_NOT_CONVERGED and got["iterations"] == 4
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 0)
    assert rows.shape == (0, model.n_out)
    with pytest.raises(eng.EngineError):
        eng.bootstrap(model, data, "centroid", 0, 1, idx=np.full((1, 250), 250, dtype=np.int32))
    # a resample that repeats one row N times has zero variance: flagged, never a crash
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 1, idx=np.zeros((1, 250), dtype=np.int32))
    assert status[0] != 0 or not np.isfinite(rows).all()


def test_medium_size_properties(eng):
    """N=20k, 16 LVs x 8 MVs: oracle parity on the fit, then size-independent properties of the bootstrap:
    identity resample == original fit; row permutation invariance."""
    N, L, K = 20000, 16, 8
    X, path = make_synthetic(N, L, K, seed=1)
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    got = eng.fit(model, data, "factorial")
    check_fit(got, orc.fit(X, [K] * L, [0] * L, path, "factorial", True))
    ident = np.arange(N, dtype=np.int32)[None, :]
    perm = np.random.default_rng(5).permutation(N).astype(np.int32)[None, :]
    rows, status, iters = eng.bootstrap(model, data, "factorial", 0, 2, idx=np.concatenate((ident, perm)))
    w, r2, tot, direct, load = model.split_row(rows)
    assert (status == 0).all() and (iters == got["iterations"]).all()
    np.testing.assert_allclose(w[0], got["weights"], rtol=1e-10)
    np.testing.assert_allclose(w[1], got["weights"], rtol=1e-10)
    np.testing.assert_allclose(load[0], got["loadings"], rtol=1e-10)
    np.testing.assert_allclose(r2[0], got["r_squared"], rtol=1e-9, atol=1e-12)
    # generated replicates against the oracle on the same Philox indices
    rows, status, iters = eng.bootstrap(model, data, "factorial", 5, 3, seed=11)
    idx = np.stack([orc.philox_indices(11, 5 + b, N) for b in range(3)])
    orows, oiters, _ = orc.bootstrap(X, idx, [K] * L, [0] * L, path, "factorial", True)
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


@pytest.mark.parametrize("case", ("syn_a", "syn_b", "syn_d", "syn_e", "syn_f", "syn_g"))
def test_sparse_tile_policy_synthetic(eng, syn, case):
    """TILES_SPARSE: only the Gram tiles the iteration needs + the P x L cross-moment pass (sign vote)."""
    N, L, K, seed = (int(v) for v in syn[case + "/gen"])
    X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[case + "/reverse"]))
    mode = 0 if str(syn[case + "/mode"]) == "A" else 1
    scheme, scaled = str(syn[case + "/scheme"]), bool(syn[case + "/scaled"])
    model = eng.Model([K] * L, [mode] * L, path, scaled, tile_policy=2)
    assert not model.full_tiles
    data = eng.Data(model, X)
    got = eng.fit(model, data, scheme)
    check_fit(got, orc.fit(X, [K] * L, [mode] * L, path, scheme, scaled))
    np.testing.assert_allclose(got["weights"], syn[case + "/weights"], rtol=REL)
    idx = np.random.default_rng(3).integers(0, N, (9, N), dtype=np.int32)
    rows, status, iters = eng.bootstrap(model, data, scheme, 0, 9, idx=idx)
    orows, oiters, _ = orc.bootstrap(X, idx, [K] * L, [mode] * L, path, scheme, scaled)
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


def test_sparse_equals_full_on_medium_model(eng):
    """N=30k, 24 LVs x 8 MVs with two reverse-coded blocks: the sparse and the full tile sets must give the
    same replicates (same Philox stream), and AUTO must pick the sparse set for this chain model."""
    N, L, K = 30000, 24, 8
    X, path = make_synthetic(N, L, K, seed=2, reverse_blocks=(5, 17))
    outs = {}
    for policy in (0, 1, 2):
        model = eng.Model([K] * L, [0] * L, path, True, tile_policy=policy)
        data = eng.Data(model, X)
        outs[policy] = (model.full_tiles, eng.fit(model, data, "centroid"),
                        eng.bootstrap(model, data, "centroid", 100, 40, seed=3))
    assert outs[1][0] and not outs[2][0] and not outs[0][0]
    check_fit(outs[2][1], orc.fit(X, [K] * L, [0] * L, path, "centroid", True))
    for policy in (0, 2):
        np.testing.assert_array_equal(outs[policy][2][2], outs[1][2][2])
        np.testing.assert_allclose(outs[policy][2][0], outs[1][2][0], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(outs[policy][1]["scores"], outs[1][1]["scores"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(outs[policy][1]["crossloadings"], outs[1][1]["crossloadings"], rtol=1e-9, atol=1e-12)


def test_sparse_ragged_mixed_modes(eng):
    rng = np.random.default_rng(8)
    sizes = [3, 12, 1, 9, 2, 5, 7, 20, 4]
    L = len(sizes)
    path = np.zeros((L, L), dtype=np.int8)
    for i in range(1, L):
        path[i, i - 1] = 1
    path[6, 2] = 1
    eta = rng.standard_normal((4000, L))
    for i in range(1, L):
        eta[:, i] += 0.7 * eta[:, i - 1]
    X = np.concatenate([eta[:, [l]] * rng.uniform(0.5, 1.0, (1, k)) * (-1 if l == 3 else 1)
                        + 0.6 * rng.standard_normal((4000, k)) for l, k in enumerate(sizes)], axis=1)
    modes = [0, 1, 0, 0, 0, 1, 0, 0, 1]
    model = eng.Model(sizes, modes, path, True, tile_policy=2)
    data = eng.Data(model, X)
    for scheme in ("centroid", "path", "factorial"):
        check_fit(eng.fit(model, data, scheme), orc.fit(X, sizes, modes, path, scheme, True))
    idx = rng.integers(0, 4000, (7, 4000), dtype=np.int32)
    rows, status, iters = eng.bootstrap(model, data, "path", 0, 7, idx=idx)
    orows, oiters, _ = orc.bootstrap(X, idx, sizes, modes, path, "path", True)
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


def test_fast_vote_decides_or_falls_back_exactly(eng, monkeypatch):
    """Sparse tile sets vote with an fp16 tensor-core pass when its error bound allows it; undecided
    replicates are redone with exact fp64 cross moments.  Both routes must reproduce the oracle."""
    # (1) independent blocks, N = 300k: far cross-correlations are noise ~ 0.0018, inside the 0.002 bound
    rng = np.random.default_rng(0)
    L, K, N = 6, 4, 300_000
    path = np.zeros((L, L), dtype=np.int8)
    for i in range(1, L):
        path[i, i - 1] = 1
    X = np.concatenate([rng.standard_normal((N, 1)) * 0.8 + 0.6 * rng.standard_normal((N, K)) for _ in range(L)], axis=1)
    model = eng.Model([K] * L, [0] * L, path, True, tile_policy=2)
    data = eng.Data(model, X)
    before = eng.redo_count()
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 6, seed=4)
    assert eng.redo_count() > before, "expected undecided votes on uncorrelated blocks"
    idx = np.stack([orc.philox_indices(4, b, N) for b in range(6)])
    orows, oiters, _ = orc.bootstrap(X, idx, [K] * L, [0] * L, path, "centroid", True)
    assert (status == 0).all()
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)
    # (1b) the same data in pipelined batches of two: batch 0's redo switches the fast vote off for the handle while
    # batch 1 (enqueued with it on) is already in flight -- its undecided replicates must still be redone
    data2 = eng.Data(model, X)
    monkeypatch.setenv("PLSPM_MAX_BATCH", "2")
    rows_b, status_b, iters_b = eng.bootstrap(model, data2, "centroid", 0, 6, seed=4)
    monkeypatch.delenv("PLSPM_MAX_BATCH")
    assert (status_b == 0).all()
    np.testing.assert_array_equal(iters_b, oiters)
    np.testing.assert_allclose(rows_b, orows, rtol=REL, atol=1e-9)
    # (2) a correlated chain with reverse-coded blocks: decided by the fast vote (no redo), signs still right
    N, L, K = 20000, 12, 8
    X, path = make_synthetic(N, L, K, seed=6, reverse_blocks=(2, 7))
    model = eng.Model([K] * L, [0] * L, path, True, tile_policy=2)
    data = eng.Data(model, X)
    before = eng.redo_count()
    rows, status, iters = eng.bootstrap(model, data, "factorial", 0, 8, seed=9)
    assert eng.redo_count() == before
    idx = np.stack([orc.philox_indices(9, b, N) for b in range(8)])
    orows, oiters, _ = orc.bootstrap(X, idx, [K] * L, [0] * L, path, "factorial", True)
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


@pytest.mark.parametrize("L,K,N,scheme,mode", ((40, 16, 5000, "centroid", 0), (70, 3, 4500, "factorial", 0),
                                               (20, 24, 6000, "path", 1), (33, 8, 9000, "centroid", 0)))
@pytest.mark.parametrize("policy", (1, 2))
def test_wide_models(eng, L, K, N, scheme, mode, policy):
    """Shapes beyond c3: P up to 640 (several column slabs), L > 32 / L not a multiple of 8, blocks spanning
    several slots (K = 16, 24), for both tile policies; single fit + bootstrap (fast vote where eligible)."""
    X, path = make_synthetic(N, L, K, seed=L + K, reverse_blocks=(1, L - 2))
    model = eng.Model([K] * L, [mode] * L, path, True, tile_policy=policy)
    data = eng.Data(model, X)
    got = eng.fit(model, data, scheme)
    check_fit(got, orc.fit(X, [K] * L, [mode] * L, path, scheme, True))
    rows, status, iters = eng.bootstrap(model, data, scheme, 3, 5, seed=21)
    idx = np.stack([orc.philox_indices(21, 3 + b, N) for b in range(5)])
    orows, oiters, _ = orc.bootstrap(X, idx, [K] * L, [mode] * L, path, scheme, True)
    assert (status == 0).all()
    np.testing.assert_array_equal(iters, oiters)
    np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)


def test_batching_and_replicate_ranges_are_transparent(eng, monkeypatch):
    """Replicates are keyed by their GLOBAL id: splitting a run into batches (PLSPM_MAX_BATCH) or into
    several calls over sub-ranges must give the same rows as one call."""
    N, L, K = 6000, 10, 8
    X, path = make_synthetic(N, L, K, seed=12, reverse_blocks=(3,))
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    rows, status, iters = eng.bootstrap(model, data, "centroid", 50, 23, seed=5)
    monkeypatch.setenv("PLSPM_MAX_BATCH", "4")
    rows_b, status_b, iters_b = eng.bootstrap(model, data, "centroid", 50, 23, seed=5)
    monkeypatch.delenv("PLSPM_MAX_BATCH")
    np.testing.assert_array_equal(rows, rows_b)
    np.testing.assert_array_equal(iters, iters_b)
    parts = [eng.bootstrap(model, data, "centroid", 50 + o, n, seed=5)[0] for o, n in ((0, 10), (10, 1), (11, 12))]
    np.testing.assert_array_equal(rows, np.concatenate(parts))


def test_tiny_and_degenerate_shapes(eng):
    """Two latent variables (effects: total == direct), single-item blocks, and N barely above P."""
    rng = np.random.default_rng(1)
    eta = rng.standard_normal((40, 1))
    X = np.concatenate([eta + 0.5 * rng.standard_normal((40, 3)), 0.7 * eta + 0.5 * rng.standard_normal((40, 1))], axis=1)
    path = np.array([[0, 0], [1, 0]], dtype=np.int8)
    for scheme in ("centroid", "factorial", "path"):
        for policy in (1, 2):
            model = eng.Model([3, 1], [0, 0], path, True, tile_policy=policy)
            data = eng.Data(model, X)
            got = eng.fit(model, data, scheme)
            check_fit(got, orc.fit(X, [3, 1], [0, 0], path, scheme, True))
            idx = rng.integers(0, 40, (11, 40), dtype=np.int32)
            rows, status, iters = eng.bootstrap(model, data, scheme, 0, 11, idx=idx)
            orows, oiters, ostatus = orc.bootstrap(X, idx, [3, 1], [0, 0], path, scheme, True)
            np.testing.assert_array_equal(iters, oiters)
            np.testing.assert_allclose(rows, orows, rtol=REL, atol=1e-9)
    X2 = rng.standard_normal((12, 2)) + np.arange(12)[:, None] * 0.3   # all blocks single-item
    model = eng.Model([1, 1], [0, 0], path, False)
    data = eng.Data(model, X2)
    got = eng.fit(model, data, "centroid")
    check_fit(got, orc.fit(X2, [1, 1], [0, 0], path, "centroid", False))
    np.testing.assert_allclose(np.abs(got["loadings"]), 1.0, rtol=1e-12)


def test_non_finite_input_is_refused(eng):
    X, path = make_synthetic(500, 3, 3, 1)
    model = eng.Model([3] * 3, [0] * 3, path, False)
    for bad in (np.nan, np.inf):
        Xb = X.copy()
        Xb[17, 4] = bad
        with pytest.raises(NotImplementedError):
            eng.Data(model, Xb)
    eng.Data(model, X).close()


def test_tensor_core_column_sums_match_fp64_and_overflow_falls_back(eng):
    """Column sums run as an exact int8 digit-plane GEMM for N >= 4096; multiplicities above 127 (only
    reachable with injected indices) must fall back to the fp64 kernel.  Both routes vs the oracle."""
    N, L, K = 6000, 4, 5
    X, path = make_synthetic(N, L, K, 21)
    X = X * np.array([1e-3, 1.0, 37.0, 1e4] * 5)[None, :] + 3.0  # columns of very different scale, non-zero mean
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    rng = np.random.default_rng(2)
    idx = rng.integers(0, N, size=(4, N)).astype(np.int32)
    idx[2, :300] = 11          # row 11 drawn 300+ times in replicate 2: overflows int8
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 4, idx=idx)
    w, r2, total, direct, load = model.split_row(rows)
    for b in range(4):
        ref, it, st = orc.replicate_row(X, idx[b], [K] * L, [0] * L, path, "centroid", True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=REL, atol=1e-9)
    # after the fallback the handle keeps working (fp64 route), and a fresh handle uses the int8 route again
    rows2, _, _ = eng.bootstrap(model, data, "centroid", 0, 2, idx=idx[:2])
    np.testing.assert_allclose(rows2, rows[:2], rtol=1e-9, atol=1e-12)
    data2 = eng.Data(model, X)
    rows3, _, _ = eng.bootstrap(model, data2, "centroid", 0, 2, idx=idx[:2])
    np.testing.assert_allclose(rows3, rows[:2], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("mode,scheme", [(0, "centroid"), (1, "path")])
def test_tensor_core_gram_route_is_fp64_accurate(eng, mode, scheme):
    """N >= 4096: the Gram tiles of a bootstrap batch come from the int8 digit-plane GEMM (exact integer sums
    of x~_p x~_q rounded to 2^-40 of the column bound).  Must be as good as fp64 accumulation: 1e-10 vs the
    oracle (whose own fp64 rounding is ~1e-12), ragged blocks (padding columns) included."""
    N, L = 9000, 6
    sizes = [5, 8, 3, 11, 8, 2]
    rng = np.random.default_rng(17)
    Xs, path = make_synthetic(N, L, max(sizes), 23)
    X = np.column_stack([Xs[:, l * max(sizes):l * max(sizes) + k] for l, k in enumerate(sizes)])
    X = X * rng.uniform(0.01, 300.0, size=X.shape[1])[None, :] + rng.normal(0, 50, size=X.shape[1])[None, :]
    model = eng.Model(sizes, [mode] * L, path, True)
    data = eng.Data(model, X)
    idx = rng.integers(0, N, size=(6, N)).astype(np.int32)
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, scheme, 0, 6, idx=idx)
    prof = eng.profile_get()
    assert prof["gram_i8"][1] > 0 and prof["gram"][1] == 0, prof
    for b in range(6):
        ref, it, st = orc.replicate_row(X, idx[b], sizes, [mode] * L, path, scheme, True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=1e-10, atol=1e-12)


def test_gram_routes_agree(tmp_path):
    """Second moments of a batch: the tcgen05 integer Gram with on-the-fly digits (default), round 1's resident /
    streamed digit planes + library GEMM (PLSPM_GRAM=cublas) and the fp64 kernels (PLSPM_GRAM=fp64) must agree
    (the switches are read once per process, hence the subprocesses)."""
    import subprocess
    import sys
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from plspm_b200 import engine\n"
        "from plspm_b200.synth import make_synthetic\n"
        "X, path = make_synthetic(21000, 5, 6, 9)\n"
        "model = engine.Model([6] * 5, [0] * 5, path, True)\n"
        "data = engine.Data(model, X)\n"
        "engine.profile_reset()\n"
        "rows, status, iters = engine.bootstrap(model, data, 'centroid', 0, 12, seed=5)\n"
        "prof = engine.profile_get()\n"
        "np.save(sys.argv[1], rows)\n"
        "print('gram', prof['gram'][1], 'gram_i8', prof['gram_i8'][1], int((status == 0).sum()))\n"
        % (os.path.join(ROOT, "plspm-python_b200"), ROOT))
    outs = {}
    for tag, env in (("mma", {}), ("split", {"PLSPM_COUNTS": "split"}), ("resident", {"PLSPM_GRAM": "cublas"}),
                     ("stream", {"PLSPM_GRAM": "cublas", "PLSPM_I8_GRAM_GB": "0.000001", "PLSPM_I8_CHUNK_GB": "0.000001"}),
                     ("fp64", {"PLSPM_GRAM": "fp64"})):
        out = tmp_path / (tag + ".npy")
        r = subprocess.run([sys.executable, str(script), str(out)], env={**os.environ, **env}, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        words = r.stdout.split()
        outs[tag] = (np.load(out), int(words[1]), int(words[3]), int(words[4]))
    assert outs["mma"][1] == 0 and outs["mma"][2] == 1                    # one gram_mma_kernel launch (finalize is its own stage)
    assert outs["resident"][1] == 0 and outs["resident"][2] == 2          # one GEMM + one combine
    assert outs["stream"][1] == 0 and outs["stream"][2] == 3 * 6            # 6 chunks of 4096 rows: generate, GEMM, combine
    assert outs["fp64"][1] >= 1 and outs["fp64"][2] == 0
    assert outs["mma"][3] == outs["resident"][3] == outs["stream"][3] == outs["fp64"][3] == 12
    np.testing.assert_allclose(outs["stream"][0], outs["resident"][0], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(outs["fp64"][0], outs["resident"][0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(outs["mma"][0], outs["fp64"][0], rtol=1e-10, atol=1e-12)
    # multiplicity images written straight by resample_images_kernel (default) vs uint32 table + two image builders
    assert np.array_equal(outs["mma"][0], outs["split"][0])


def test_fused_multiplicity_images_over_several_row_ranges(eng):
    """More rows than one CTA's shared memory holds as bytes (204800): the replicate's draws are scanned once per row range."""
    N, L, K = 2 * 204800 + 333, 3, 3
    X, path = make_synthetic(N, L, K, 43)
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, "centroid", 11, 3, seed=9)
    prof = eng.profile_get()
    assert prof["gram_i8"][1] == 1 and prof["gram"][1] == 0 and prof["colsum"][1] == 0, prof
    for b in (0, 2):
        idx = orc.philox_indices(9, 11 + b, N)
        ref, it, st = orc.replicate_row(X, idx, [K] * L, [0] * L, path, "centroid", True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=1e-6, atol=1e-9)


def test_fused_multiplicity_images_with_user_indices_and_overflow(eng):
    """resample_images_kernel counts a replicate's draws as packed bytes in shared memory: user-supplied indices take
    the same kernel, rows at the image padding edge (N not a multiple of 64 / 128) must land right, and a row drawn
    more than 127 times is flagged before its byte can carry -- the batch is then redone on the fp64 route."""
    N, L, K = 8192 + 77, 4, 5
    X, path = make_synthetic(N, L, K, 41)
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    rng = np.random.default_rng(5)
    idx = rng.integers(0, N, size=(5, N)).astype(np.int32)
    idx[1, :300] = N - 1            # 300 copies of the last row: above the int8 range, and past 255
    idx[3, 1000:1100] = 4242        # 100 copies: fits
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 5, idx=idx)
    prof = eng.profile_get()
    assert prof["gram_i8"][1] == 1 and prof["gram"][1] >= 1, prof   # tensor-core attempt, then the fp64 redo
    for b in range(5):
        ref, it, st = orc.replicate_row(X, idx[b], [K] * L, [0] * L, path, "centroid", True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=1e-6, atol=1e-9)
    # without the overflow the batch stays on the tensor-core route (a fresh handle: the fallback is sticky per handle)
    data = eng.Data(model, X)
    idx[1, :300] = rng.integers(0, N, size=300)
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 5, idx=idx)
    prof = eng.profile_get()
    assert prof["gram_i8"][1] == 1 and prof["gram"][1] == 0 and prof["colsum"][1] == 0, prof
    for b in range(5):
        ref, it, st = orc.replicate_row(X, idx[b], [K] * L, [0] * L, path, "centroid", True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("outlier,integer_route", ((1.0e3, True), (1.0e6, False)))
def test_integer_gram_and_outlier_columns(eng, outlier, integer_route):
    """The integer Gram rounds every pair product to 2^-47 of the product of the two COLUMN BOUNDS.  A moderate
    outlier (1000 sd) costs nothing visible; a column whose bound is set by a few gross outliers (1e6 sd) would
    lose the variance of replicates that miss them (measured 4e-6 on the weights), so the upload detects it
    (heavy-tail guard) and the fp64 kernels take over.  Either way the rows match the oracle at 1e-6."""
    N, L, K = 8192, 5, 4
    X, path = make_synthetic(N, L, K, 31)
    X[1234, 6] = outlier    # one outlier in a column of unit scale
    X[:, 13] *= 1.0e-5      # and a column of tiny scale
    model = eng.Model([K] * L, [0] * L, path, True)
    data = eng.Data(model, X)
    eng.profile_reset()
    rows, status, iters = eng.bootstrap(model, data, "centroid", 0, 6, seed=11)
    prof = eng.profile_get()
    assert (prof["gram_i8"][1] > 0) == integer_route and (prof["gram"][1] > 0) == (not integer_route), prof
    for b in range(6):
        ref, it, st = orc.replicate_row(X, orc.philox_indices(11, b, N), [K] * L, [0] * L, path, "centroid", True)
        assert status[b] == st == 0 and iters[b] == it
        np.testing.assert_allclose(rows[b], ref, rtol=REL, atol=1e-9)


def test_bootstrap_host_matches_resident_path_and_prefetches_the_images(eng):
    """plspm_bootstrap_host (upload + bootstrap + release in one call): the multiplicity images of its first batch are
    generated on a side stream while X crosses PCIe.  Rows must be identical to the resident-data path, for one batch,
    for several batches (only the first is prefetched) and for injected indices (no prefetch)."""
    N, L, K = 9000, 6, 5
    X, path = make_synthetic(N, L, K, 29)
    model = eng.Model([K] * L, [0] * L, path, True, eng.TILES_SPARSE)
    data = eng.Data(model, X)
    for reps in (37, 300):
        ref_rows, ref_status, ref_iters = eng.bootstrap(model, data, "centroid", 5, reps, seed=77)
        eng.profile_reset()
        rows, status, iters = eng.bootstrap_host(model, np.ascontiguousarray(X), "centroid", 5, reps, seed=77)
        prof = eng.profile_get()
        assert prof["gram_i8"][1] >= 1 and prof["gram"][1] == 0 and prof["colsum"][1] == 0, prof
        assert np.array_equal(rows, ref_rows) and np.array_equal(status, ref_status) and np.array_equal(iters, ref_iters)
    # one prefetch kernel on the side stream is not a stage launch of the handle's stream: the counts stage shows
    # only the batches after the first (none here)
    eng.profile_reset()
    eng.bootstrap_host(model, np.ascontiguousarray(X), "centroid", 0, 64, seed=1)
    assert eng.profile_get()["counts"][1] == 0
    idx = np.random.default_rng(2).integers(0, N, size=(9, N)).astype(np.int32)
    a = eng.bootstrap(model, data, "centroid", 0, 9, idx=idx)[0]
    b = eng.bootstrap_host(model, np.ascontiguousarray(X), "centroid", 0, 9, idx=idx)[0]
    assert np.array_equal(a, b)
    ref, it, st = orc.replicate_row(X, idx[4], [K] * L, [0] * L, path, "centroid", True)
    np.testing.assert_allclose(b[4], ref, rtol=1e-6, atol=1e-9)
