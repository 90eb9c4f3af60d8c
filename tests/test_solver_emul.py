"""Solver logic (csrc/solver_core.h compiled for the host, tests/emul) vs the oracle.  CPU only.
The GPU parity tests (tests/test_gpu_parity.py) run the same comparisons through the CUDA library."""
import numpy as np
import pytest

from oracle import plspm_oracle as orc
from plspm_b200.synth import make_synthetic
from tests.emul import emul

REL = 1e-6  # north_star tolerance; the covariance-domain solver is in practice ~1e-12


def check(r, o, L, rel=1e-9, crossloadings=True):
    assert r["status"] == 0
    assert r["iterations"] == o["iterations"]
    np.testing.assert_allclose(r["weights"], o["weights"], rtol=rel)
    np.testing.assert_allclose(r["scores"], o["scores"], rtol=rel, atol=1e-10)
    np.testing.assert_allclose(r["path_coefficients"], o["path_coefficients"], rtol=rel, atol=1e-11)
    np.testing.assert_allclose(r["total_effects"], o["total_effects"], rtol=rel, atol=1e-11)
    np.testing.assert_allclose(r["r_squared"], o["r_squared"], rtol=rel, atol=1e-11)
    np.testing.assert_allclose(r["loadings"], o["loadings"], rtol=rel, atol=1e-11)
    if crossloadings:
        np.testing.assert_allclose(r["crossloadings"], o["crossloadings"], rtol=rel, atol=1e-10)


@pytest.mark.parametrize("scheme", ("centroid", "factorial", "path"))
@pytest.mark.parametrize("mode", (0, 1))
@pytest.mark.parametrize("scaled", (False, True))
def test_satisfaction(sat, scheme, mode, scaled):
    L = 6
    o = orc.fit(sat["X"], sat["block_sizes"], [mode] * L, sat["path"], scheme, scaled)
    r = emul.fit(sat["X"], sat["block_sizes"], [mode] * L, sat["path"], scheme, scaled)
    check(r, o, L, rel=1e-9 if mode == 0 else 1e-7)


def test_satisfaction_golden_r(sat):
    r = emul.fit(sat["X"], sat["block_sizes"], [0] * 6, sat["path"], "centroid", False)
    assert r["iterations"] == 4
    np.testing.assert_allclose(r["weights"], sat["R/centroid/weight"], rtol=REL)
    np.testing.assert_allclose(r["scores"], sat["R/scores"], rtol=REL, atol=1e-9)
    np.testing.assert_allclose(r["loadings"], sat["R/centroid/loading"], rtol=REL)


@pytest.mark.parametrize("case", ("syn_a", "syn_b", "syn_c", "syn_d", "syn_e", "syn_f", "syn_g", "syn_h"))
def test_synthetic_cases(syn, case):
    N, L, K, seed = (int(v) for v in syn[case + "/gen"])
    X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[case + "/reverse"]))
    mode = 0 if str(syn[case + "/mode"]) == "A" else 1
    scheme, scaled = str(syn[case + "/scheme"]), bool(syn[case + "/scaled"])
    o = orc.fit(X, [K] * L, [mode] * L, path, scheme, scaled)
    r = emul.fit(X, [K] * L, [mode] * L, path, scheme, scaled)
    check(r, o, L, rel=1e-8)
    np.testing.assert_allclose(r["weights"], syn[case + "/weights"], rtol=REL)


def test_bootstrap_rows(sat):
    idx = np.random.default_rng(1234).integers(0, 250, (1000, 250), dtype=np.int32)[:12]
    for scheme, mode, scaled in (("centroid", 0, False), ("path", 1, True)):
        out, iters, status = orc.bootstrap(sat["X"], idx, sat["block_sizes"], [mode] * 6, sat["path"], scheme, scaled)
        for b in range(idx.shape[0]):
            r = emul.fit(sat["X"], sat["block_sizes"], [mode] * 6, sat["path"], scheme, scaled, idx=idx[b])
            assert r["iterations"] == iters[b]
            np.testing.assert_allclose(r["out_row"], out[b], rtol=1e-7, atol=1e-10)


def test_ragged_blocks_and_mixed_modes():
    rng = np.random.default_rng(3)
    sizes = [1, 9, 3, 17, 2]
    L = len(sizes)
    path = np.zeros((L, L), dtype=np.int8)
    path[1, 0] = path[2, 0] = path[2, 1] = path[3, 2] = path[4, 1] = path[4, 3] = 1
    eta = rng.standard_normal((300, L))
    for i in range(1, L):
        eta[:, i] += eta[:, :i] @ (0.5 * path[i, :i])
    X = np.concatenate([eta[:, [l]] * rng.uniform(0.5, 1.0, (1, k)) + 0.7 * rng.standard_normal((300, k))
                        for l, k in enumerate(sizes)], axis=1) * 3.0 + 10.0
    modes = [0, 1, 0, 0, 1]
    for scheme in ("centroid", "factorial", "path"):
        o = orc.fit(X, sizes, modes, path, scheme, True)
        r = emul.fit(X, sizes, modes, path, scheme, True)
        check(r, o, L, rel=1e-8)


def test_not_converged_status(sat):
    r = emul.fit(sat["X"], sat["block_sizes"], [1] * 6, sat["path"], "centroid", True, tol=1e-30, max_iter=3)
    assert r["status"] == 1 and r["iterations"] == 4  # max_iter + 1 iterate calls (quirk Q4)


def test_model_tables(sat):
    info = emul.model_info(sat["block_sizes"], [0] * 6, sat["path"], 0)
    assert info["P"] == 27 and info["Ppad"] == 48 and info["n_tiles"] == 21 and info["n_eff"] == 15
    pairs = orc.effect_pairs(sat["path"])
    assert [(int(f), int(t)) for f, t in zip(info["eff_from"], info["eff_to"])] == pairs


@pytest.mark.parametrize("case", ("syn_b", "syn_d", "syn_e", "syn_f"))
def test_sparse_tile_set_matches_oracle(syn, case):
    """TILES_SPARSE (policy 2): only the Gram tiles the iteration needs + the P x L cross-moment pass."""
    N, L, K, seed = (int(v) for v in syn[case + "/gen"])
    X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[case + "/reverse"]))
    mode = 0 if str(syn[case + "/mode"]) == "A" else 1
    scheme, scaled = str(syn[case + "/scheme"]), bool(syn[case + "/scaled"])
    r = emul.fit(X, [K] * L, [mode] * L, path, scheme, scaled, tile_policy=2)
    assert r["info"]["full"] == 0
    check(r, orc.fit(X, [K] * L, [mode] * L, path, scheme, scaled), L, rel=1e-8)
    idx = np.random.default_rng(1).integers(0, N, N, dtype=np.int32)
    rb = emul.fit(X, [K] * L, [mode] * L, path, scheme, scaled, idx=idx, tile_policy=2)
    row, it, st = orc.replicate_row(X, idx, [K] * L, [mode] * L, path, scheme, scaled)
    assert rb["iterations"] == it
    np.testing.assert_allclose(rb["out_row"], row, rtol=1e-7, atol=1e-10)


def test_sparse_ragged_blocks():
    rng = np.random.default_rng(8)
    sizes = [3, 12, 1, 9, 2, 5, 7]
    L = len(sizes)
    path = np.zeros((L, L), dtype=np.int8)
    for i in range(1, L):
        path[i, i - 1] = 1
    path[6, 2] = 1
    eta = rng.standard_normal((500, L))
    for i in range(1, L):
        eta[:, i] += 0.7 * eta[:, i - 1]
    X = np.concatenate([eta[:, [l]] * rng.uniform(0.5, 1.0, (1, k)) * (-1 if l == 3 else 1)
                        + 0.6 * rng.standard_normal((500, k)) for l, k in enumerate(sizes)], axis=1)
    modes = [0, 1, 0, 0, 0, 1, 0]
    for scheme in ("centroid", "path"):
        r = emul.fit(X, sizes, modes, path, scheme, True, tile_policy=2)
        assert r["info"]["full"] == 0
        check(r, orc.fit(X, sizes, modes, path, scheme, True), L, rel=1e-8)


def test_low_precision_vote_logic(syn):
    """Solver phase 3: far LV pairs vote with the sign of a perturbed low-precision cross moment when the
    error bound allows it; otherwise the replicate is flagged and redone exactly.  Either way the result
    must equal the oracle."""
    emul.set_vote_mode(1)
    try:
        decided = 0
        for case in ("syn_b", "syn_e", "syn_f"):
            N, L, K, seed = (int(v) for v in syn[case + "/gen"])
            X, path = make_synthetic(N, L, K, seed, reverse_blocks=tuple(int(v) for v in syn[case + "/reverse"]))
            mode = 0 if str(syn[case + "/mode"]) == "A" else 1
            scheme, scaled = str(syn[case + "/scheme"]), bool(syn[case + "/scaled"])
            r = emul.fit(X, [K] * L, [mode] * L, path, scheme, scaled, tile_policy=2)
            decided += 0 if emul.last_ambiguous() else 1
            # (the fast vote serves bootstrap rows, which carry no crossloadings)
            check(r, orc.fit(X, [K] * L, [mode] * L, path, scheme, scaled), L, rel=1e-8,
                  crossloadings=emul.last_ambiguous())
        assert decided >= 2  # strongly correlated chains are decided by the fast vote
        # independent blocks: the far cross moments are noise around zero -> must fall back, still exact
        rng = np.random.default_rng(0)
        L, K, N = 6, 4, 300_000  # noise correlations ~ 1/sqrt(N) = 0.0018 sit inside the 0.002 error bound
        path = np.zeros((L, L), dtype=np.int8)
        for i in range(1, L):
            path[i, i - 1] = 1
        X = np.concatenate([rng.standard_normal((N, 1)) * 0.8 + 0.6 * rng.standard_normal((N, K)) for _ in range(L)],
                           axis=1)
        r = emul.fit(X, [K] * L, [0] * L, path, "centroid", True, tile_policy=2)
        assert emul.last_ambiguous()
        check(r, orc.fit(X, [K] * L, [0] * L, path, "centroid", True), L, rel=1e-8)
    finally:
        emul.set_vote_mode(0)


def test_rank_deficient_blocks_and_regressions_give_the_reference_min_norm_solution():
    """Rank-deficient least squares: the reference's SVD-based solvers (scipy lstsq in mode.py:50-52, statsmodels' pinv in
    scheme.py:50 / inner_model.py:76-77) return minimum-norm coefficients.  The solver falls back from Cholesky to
    conjugate gradients on the (consistent, positive semi-definite) normal equations, which converge to the same
    solution.  (a) Mode-B block with a duplicated column against the REFERENCE's output (tests/golden/collinear.npz);
    (b) two identical blocks feeding a third LV (singular inner regression) against the oracle, all schemes."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collinear.npz"))
    for policy in (1, 2):
        r = emul.fit(g["X"], g["block_sizes"], [1, 1, 1], g["path"], "centroid", True, tile_policy=policy)
        assert r["status"] == 0 and r["iterations"] == int(g["ref/iterations"])
        np.testing.assert_allclose(r["weights"], g["ref/weights"], rtol=1e-9)
        np.testing.assert_allclose(r["r_squared"], g["ref/r_squared"], rtol=1e-9, atol=1e-12)
        assert abs(r["weights"][3] - r["weights"][4]) < 1e-12  # equal weights on the two copies: the min-norm answer
    rng = np.random.default_rng(1)
    N = 400
    f = rng.standard_normal((N, 1))
    b0 = f + 0.5 * rng.standard_normal((N, 3))
    X = np.concatenate([b0, b0.copy(), f + 0.4 * rng.standard_normal((N, 3))], axis=1)
    path = np.zeros((3, 3), dtype=np.int8)
    path[2, 0] = path[2, 1] = 1
    for scheme in ("centroid", "path", "factorial"):
        ref = orc.fit(X, [3, 3, 3], [0, 0, 0], path, scheme, True)
        for policy in (1, 2):
            r = emul.fit(X, [3, 3, 3], [0, 0, 0], path, scheme, True, tile_policy=policy)
            assert r["status"] == 0 and r["iterations"] == ref["iterations"]
            check(r, ref, 3, rel=1e-8)
            assert abs(r["path_coefficients"][2, 0] - r["path_coefficients"][2, 1]) < 1e-9


def test_resumed_phases_equal_a_second_iteration():
    """Sparse tile sets: phase 1 stores the converged iteration state (w, V, dinv, R, pooled scale, counts) and phases
    2 / 3 resume from it (what the CUDA library does since round 2).  Must give exactly what iterating again gives."""
    rng = np.random.default_rng(5)
    for scheme, mode, scaled in (("centroid", 0, True), ("path", 1, True), ("factorial", 0, False)):
        X, path = make_synthetic(700, 7, 4, 13, reverse_blocks=(3,))
        idx = rng.integers(0, 700, 700).astype(np.int32)
        outs = []
        for vote_mode in (0, 1):
            emul.set_vote_mode(vote_mode)
            pair = []
            for resume in (True, False):
                emul.set_resume(resume)
                pair.append(emul.fit(X, [4] * 7, [mode] * 7, path, scheme, scaled, idx=idx, tile_policy=2))
            emul.set_resume(True)
            emul.set_vote_mode(0)
            a, b = pair
            assert a["info"]["full"] == 0 and a["iterations"] == b["iterations"] and a["status"] == b["status"] == 0
            for key in ("out_row", "weights", "loadings", "r_squared", "path_coefficients", "total_effects"):
                np.testing.assert_array_equal(a[key], b[key], err_msg="%s %s vote_mode=%d" % (scheme, key, vote_mode))
            outs.append(a)
        ref, it, st = orc.replicate_row(X, idx, [4] * 7, [mode] * 7, path, scheme, scaled)
        assert st == 0 and it == outs[0]["iterations"]
        np.testing.assert_allclose(outs[0]["out_row"], ref, rtol=1e-8, atol=1e-11)


# ---- non-metric path with numeric scales (solver_num.h) ------------------------------------------
@pytest.fixture(scope="module")
def nm():
    import os
    from tests.conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "nonmetric.npz"), allow_pickle=False)


@pytest.mark.parametrize("scheme", ("centroid", "factorial", "path"))
@pytest.mark.parametrize("mode", (0, 1))
def test_nonmetric_num_russa(nm, scheme, mode):
    from oracle import plspm_oracle_nonmetric as onm
    o = onm.fit_num(nm["russa/X"], nm["russa/block_sizes"], [mode] * 3, nm["russa/path"], scheme, tol=1e-7)
    r = emul.fit_num(nm["russa/X"], nm["russa/block_sizes"], [mode] * 3, nm["russa/path"], scheme, tol=1e-7,
                     tile_policy=1)
    check(r, o, 3, rel=1e-8)
    tag = "russa/%s/%s/" % (scheme, "AB"[mode])
    np.testing.assert_allclose(r["weights"], nm[tag + "weights"], rtol=1e-7)
    np.testing.assert_allclose(r["scores"], nm[tag + "scores"], rtol=1e-7, atol=1e-9)


def test_nonmetric_num_mobi_and_synthetic(nm):
    from oracle import plspm_oracle_nonmetric as onm
    o = onm.fit_num(nm["mobi/X"], nm["mobi/block_sizes"], nm["mobi/modes"], nm["mobi/path"], "path", tol=1e-8)
    r = emul.fit_num(nm["mobi/X"], nm["mobi/block_sizes"], nm["mobi/modes"], nm["mobi/path"], "path", tol=1e-8,
                     tile_policy=1)
    check(r, o, 5, rel=1e-8)
    np.testing.assert_allclose(r["weights"], nm["mobi/weights"], rtol=1e-7)
    X, path = make_synthetic(700, 7, 5, seed=4, reverse_blocks=(2,))
    for scheme, mode in (("centroid", 0), ("factorial", 1), ("path", 0)):
        o = onm.fit_num(X, [5] * 7, [mode] * 7, path, scheme)
        for policy in (1, 2):
            r = emul.fit_num(X, [5] * 7, [mode] * 7, path, scheme, tile_policy=policy)
            check(r, o, 7, rel=1e-8, crossloadings=(policy == 1))
    idx = np.random.default_rng(2).integers(0, 700, 700, dtype=np.int32)
    o = onm.fit_num(X[idx], [5] * 7, [0] * 7, path, "centroid")
    r = emul.fit_num(X, [5] * 7, [0] * 7, path, "centroid", idx=idx, tile_policy=2)
    assert r["iterations"] == o["iterations"]
    np.testing.assert_allclose(r["weights"], o["weights"], rtol=1e-8)
    np.testing.assert_allclose(r["path_coefficients"], o["path_coefficients"], rtol=1e-8, atol=1e-11)


def test_nonmetric_not_converged(nm):
    r = emul.fit_num(nm["russa/X"], nm["russa/block_sizes"], [1] * 3, nm["russa/path"], "centroid", tol=1e-30,
                     max_iter=3, tile_policy=1)
    assert r["status"] == 1 and r["iterations"] == 4
