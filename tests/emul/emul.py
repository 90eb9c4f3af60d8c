"""ctypes wrapper of the TEST-ONLY host emulation (tests/emul/solver_emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libplspm_emul.so")
        srcs = [os.path.join(HERE, "solver_emul.cpp"),
                os.path.join(ROOT, "plspm-python_b200", "csrc", "plspm_model.cpp")]
        deps = srcs + [os.path.join(ROOT, "plspm-python_b200", "csrc", f)
                       for f in ("plspm_model.h", "solver_core.h", "solver_num.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so] + srcs)
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def model_info(block_sizes, modes, path, scaled, tile_policy=0):
    L = len(block_sizes)
    bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
    md = np.ascontiguousarray(modes, dtype=np.int8)
    pm = np.ascontiguousarray(path, dtype=np.int8)
    info = np.zeros(8, dtype=np.int32)
    ef = np.zeros(L * L, dtype=np.int32)
    et = np.zeros(L * L, dtype=np.int32)
    rc = lib().emul_model_info(L, _p(bs, ctypes.c_int32), _p(md, ctypes.c_int8), _p(pm, ctypes.c_int8), int(scaled),
                               tile_policy, _p(info, ctypes.c_int32), _p(ef, ctypes.c_int32), _p(et, ctypes.c_int32))
    assert rc == 0
    n = int(info[5])
    return dict(P=int(info[0]), Ppad=int(info[1]), n_tiles=int(info[2]), n_tg=int(info[3]), n_pairs=int(info[4]),
                n_eff=n, n_out=int(info[6]), full=int(info[7]), eff_from=ef[:n].copy(), eff_to=et[:n].copy())


def fit(X, block_sizes, modes, path, scheme, scaled, idx=None, tol=1e-6, max_iter=100, tile_policy=0):
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    L = len(block_sizes)
    bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
    md = np.ascontiguousarray(modes, dtype=np.int8)
    pm = np.ascontiguousarray(path, dtype=np.int8)
    info = model_info(bs, md, pm, scaled, tile_policy)
    out = dict(out_row=np.zeros(info["n_out"]), weights=np.zeros(P), loadings=np.zeros(P), r_squared=np.zeros(L),
               path_coefficients=np.zeros((L, L)), total_effects=np.zeros((L, L)), crossloadings=np.zeros((P, L)),
               scores=np.zeros((N, L)))
    iters = np.zeros(1, dtype=np.int32)
    status = np.zeros(1, dtype=np.int32)
    idx_a = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    d = ctypes.c_double
    f = lib().emul_fit
    f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int8),
                  ctypes.POINTER(ctypes.c_int8), ctypes.c_int, ctypes.c_int, ctypes.POINTER(d), ctypes.c_int64,
                  ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_double, ctypes.c_int] + \
                 [ctypes.POINTER(d)] * 8 + [ctypes.POINTER(ctypes.c_int32)] * 2
    rc = f(L, _p(bs, ctypes.c_int32), _p(md, ctypes.c_int8), _p(pm, ctypes.c_int8), int(scaled), tile_policy,
           _p(X, d), N, _p(idx_a, ctypes.c_int32), {"centroid": 0, "factorial": 1, "path": 2}[scheme], tol, max_iter,
           _p(out["out_row"], d), _p(out["weights"], d), _p(out["loadings"], d), _p(out["r_squared"], d),
           _p(out["path_coefficients"], d), _p(out["total_effects"], d), _p(out["crossloadings"], d),
           _p(out["scores"], d), _p(iters, ctypes.c_int32), _p(status, ctypes.c_int32))
    assert rc == 0
    # the criterion's decomposition (second-moment part + sign-change rows) must equal the direct evaluation
    assert lib().emul_num_decomposition_errors() == 0
    out["iterations"] = int(iters[0])
    out["status"] = int(status[0])
    out["info"] = info
    return out


def set_vote_mode(mode: int):
    """0: exact cross moments; 1: emulate the low-precision sign vote (solver phase 3) with a worst-case-ish error."""
    lib().emul_set_vote_mode(int(mode))


def set_resume(on: bool):
    """True (default): solver phases 2 / 3 resume from the converged state phase 1 saved, like the CUDA library;
    False: they iterate again from the initial weights (round 1's behaviour)."""
    lib().emul_set_resume(int(bool(on)))


def last_ambiguous() -> bool:
    return bool(lib().emul_last_ambiguous())


def fit_num(X, block_sizes, modes, path, scheme, idx=None, tol=1e-6, max_iter=100, tile_policy=0):
    """Non-metric fit with numeric scales (Scale.NUM / RAW, complete data) through num_step()."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    L = len(block_sizes)
    bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
    md = np.ascontiguousarray(modes, dtype=np.int8)
    pm = np.ascontiguousarray(path, dtype=np.int8)
    info = model_info(bs, md, pm, False, tile_policy)
    out = dict(out_row=np.zeros(info["n_out"]), weights=np.zeros(P), loadings=np.zeros(P), r_squared=np.zeros(L),
               path_coefficients=np.zeros((L, L)), total_effects=np.zeros((L, L)), crossloadings=np.zeros((P, L)),
               scores=np.zeros((N, L)))
    iters = np.zeros(1, dtype=np.int32)
    status = np.zeros(1, dtype=np.int32)
    idx_a = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    d = ctypes.c_double
    f = lib().emul_fit_num
    f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int8),
                  ctypes.POINTER(ctypes.c_int8), ctypes.c_int, ctypes.POINTER(d), ctypes.c_int64,
                  ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_double, ctypes.c_int] + \
                 [ctypes.POINTER(d)] * 8 + [ctypes.POINTER(ctypes.c_int32)] * 2
    rc = f(L, _p(bs, ctypes.c_int32), _p(md, ctypes.c_int8), _p(pm, ctypes.c_int8), tile_policy, _p(X, d), N,
           _p(idx_a, ctypes.c_int32), {"centroid": 0, "factorial": 1, "path": 2}[scheme], tol, max_iter,
           _p(out["out_row"], d), _p(out["weights"], d), _p(out["loadings"], d), _p(out["r_squared"], d),
           _p(out["path_coefficients"], d), _p(out["total_effects"], d), _p(out["crossloadings"], d),
           _p(out["scores"], d), _p(iters, ctypes.c_int32), _p(status, ctypes.c_int32))
    assert rc == 0
    # the criterion's decomposition (second-moment part + sign-change rows) must equal the direct evaluation
    assert lib().emul_num_decomposition_errors() == 0
    out["iterations"] = int(iters[0])
    out["status"] = int(status[0])
    out["info"] = info
    return out
