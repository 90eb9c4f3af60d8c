"""ctypes binding of libplspm_b200.so (the C ABI declared in include/plspm_b200.h).

This is the only place the Python host touches the CUDA library.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is present, every call
raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplspm_b200.so")

SCHEME_IDS = {"centroid": 0, "factorial": 1, "path": 2}
TILES_AUTO, TILES_FULL, TILES_SPARSE = 0, 1, 2
STATUS_OK, STATUS_NOT_CONVERGED, STATUS_SINGULAR = 0, 1, 2

EXPORTS = (
    "plspm_version", "plspm_last_error", "plspm_device_count", "plspm_set_device", "plspm_model_create",
    "plspm_model_destroy", "plspm_model_query", "plspm_model_effects", "plspm_data_create", "plspm_data_destroy",
    "plspm_fit", "plspm_bootstrap", "plspm_bootstrap_host", "plspm_resample_indices", "plspm_profile_reset",
    "plspm_profile_get", "plspm_host_alloc", "plspm_host_free", "plspm_redo_count", "plspm_model_set_numeric",
    "plspm_pool_trim", "plspm_pool_set_limit", "plspm_data_set_imputation",
)

_lib = None
_c_dp = ctypes.POINTER(ctypes.c_double)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_i8p = ctypes.POINTER(ctypes.c_int8)
_c_i64p = ctypes.POINTER(ctypes.c_int64)


class EngineError(RuntimeError):
    pass


def load():
    """Loads the CUDA library (building is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError("libplspm_b200.so is missing (%s): build it with `python __graft_entry__.py` -- "
                          "plspm_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    lib.plspm_last_error.restype = ctypes.c_char_p
    lib.plspm_device_count.argtypes = [_c_i32p]
    lib.plspm_set_device.argtypes = [i32]
    lib.plspm_model_create.argtypes = [i32, _c_i32p, _c_i8p, _c_i8p, i32, i32, ctypes.POINTER(vp)]
    lib.plspm_model_destroy.argtypes = [vp]
    lib.plspm_model_destroy.restype = None
    lib.plspm_model_query.argtypes = [vp, _c_i32p]
    lib.plspm_model_effects.argtypes = [vp, _c_i32p, _c_i32p]
    lib.plspm_model_set_numeric.argtypes = [vp, i32]
    lib.plspm_data_create.argtypes = [vp, vp, i64, i64, i32, ctypes.POINTER(vp)]
    lib.plspm_data_destroy.argtypes = [vp]
    lib.plspm_data_destroy.restype = None
    lib.plspm_fit.argtypes = [vp, vp, i32, dbl, i32] + [_c_dp] * 7 + [_c_i32p, _c_i32p]
    lib.plspm_bootstrap.argtypes = [vp, vp, i32, dbl, i32, i64, i64, u64, _c_i32p, vp, i32, _c_i32p, _c_i32p]
    lib.plspm_bootstrap_host.argtypes = [vp, vp, i64, i64, i32, dbl, i32, i64, i64, u64, _c_i32p, _c_dp, _c_i32p,
                                         _c_i32p]
    lib.plspm_resample_indices.argtypes = [u64, i64, i64, _c_i32p]
    lib.plspm_profile_get.argtypes = [_c_dp, _c_i64p]
    lib.plspm_redo_count.argtypes = [_c_i64p]
    lib.plspm_host_alloc.argtypes = [ctypes.POINTER(vp), i64]
    lib.plspm_host_free.argtypes = [vp]
    lib.plspm_pool_set_limit.argtypes = [i64]
    lib.plspm_data_set_imputation.argtypes = [vp, vp, _c_i8p]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise EngineError("plspm_b200 error %d: %s" % (rc, load().plspm_last_error().decode()))


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def device_count() -> int:
    n = ctypes.c_int32(0)
    rc = load().plspm_device_count(ctypes.byref(n))
    return int(n.value) if rc == 0 else 0


def set_device(device: int):
    _check(load().plspm_set_device(int(device)))


def scheme_id(scheme) -> int:
    if isinstance(scheme, str):
        return SCHEME_IDS[scheme.lower()]
    return int(scheme)


class Model:
    """Lowered model: LV blocks (path order), modes (0 = A, 1 = B), path matrix, `scaled` flag."""

    def __init__(self, block_sizes, modes, path, scaled: bool, tile_policy: int = 0, numeric: bool = False):
        self.block_sizes = np.ascontiguousarray(block_sizes, dtype=np.int32)
        self.modes = np.ascontiguousarray(modes, dtype=np.int8)
        self.path = np.ascontiguousarray(path, dtype=np.int8)
        self.L = int(len(self.block_sizes))
        if self.path.shape != (self.L, self.L) or len(self.modes) != self.L:
            raise ValueError("block_sizes, modes and path do not describe the same latent variables")
        self.scaled = bool(scaled)
        self._h = ctypes.c_void_p()
        _check(load().plspm_model_create(self.L, _ptr(self.block_sizes, _c_i32p), _ptr(self.modes, _c_i8p),
                                         _ptr(self.path, _c_i8p), int(self.scaled), int(tile_policy),
                                         ctypes.byref(self._h)))
        self.numeric = bool(numeric)
        if self.numeric:  # non-metric estimator for all-NUM / RAW scales (weights.py:73-133)
            _check(load().plspm_model_set_numeric(self._h, 1))
        info = np.zeros(16, dtype=np.int32)
        _check(load().plspm_model_query(self._h, _ptr(info, _c_i32p)))
        self.P, self.Ppad, self.n_tiles, self.n_tile_groups = int(info[1]), int(info[2]), int(info[3]), int(info[4])
        self.n_effects, self.n_out, self.full_tiles = int(info[6]), int(info[7]), bool(info[8])
        self.n_cross_tiles = int(info[10])
        self.n_pair_columns = int(info[12])
        ef = np.zeros(max(self.n_effects, 1), dtype=np.int32)
        et = np.zeros(max(self.n_effects, 1), dtype=np.int32)
        _check(load().plspm_model_effects(self._h, _ptr(ef, _c_i32p), _ptr(et, _c_i32p)))
        self.effects_from, self.effects_to = ef[:self.n_effects].copy(), et[:self.n_effects].copy()

    def close(self):
        if self._h:
            load().plspm_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def split_row(self, rows: np.ndarray):
        """Splits bootstrap rows [.., n_out] into (weights, r_squared, total, direct, loadings)."""
        P, L, E = self.P, self.L, self.n_effects
        return (rows[..., :P], rows[..., P:P + L], rows[..., P + L:P + L + E], rows[..., P + L + E:P + L + 2 * E],
                rows[..., P + L + 2 * E:])


class Data:
    """Observation x manifest matrix resident in HBM (columns grouped by LV in path order)."""

    def __init__(self, model: Model, X, device_ptr: int = 0, n_rows: int = 0, ld: int = 0):
        self.model = model
        self._h = ctypes.c_void_p()
        if device_ptr:
            self.N = int(n_rows)
            _check(load().plspm_data_create(model._h, ctypes.c_void_p(device_ptr), self.N, int(ld or model.P), 1,
                                            ctypes.byref(self._h)))
        else:
            X = np.asarray(X)
            if X.ndim != 2 or X.shape[1] != model.P:
                raise ValueError("X must be [N, %d]" % model.P)
            if X.dtype != np.float64 or not X.flags.c_contiguous:
                X = np.ascontiguousarray(X, dtype=np.float64)
            self.N = int(X.shape[0])
            try:
                _check(load().plspm_data_create(model._h, X.ctypes.data_as(ctypes.c_void_p), self.N, int(X.shape[1]), 0,
                                                ctypes.byref(self._h)))
            except EngineError as e:  # the library finds NaN / Inf through the column means (no host pass over X)
                if "non-finite" in str(e):
                    raise NotImplementedError("missing / non-finite values are not supported by the CUDA path yet")
                raise

    def set_imputation(self, base: Model, has_missing):
        """This handle holds the augmented matrix [x0 | missing indicators] (see include/plspm_b200.h): bootstrap
        replicates are re-imputed with the column means of their own observed rows and solved under `base`."""
        hm = np.ascontiguousarray(has_missing, dtype=np.int8)
        if hm.shape != (base.P,):
            raise ValueError("has_missing must be [%d]" % base.P)
        _check(load().plspm_data_set_imputation(self._h, base._h, _ptr(hm, _c_i8p)))
        self.rows_model = base  # keeps the base model alive; bootstrap rows have ITS layout

    def close(self):
        if self._h:
            load().plspm_data_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fit(model: Model, data: Data, scheme, tol: float = 1e-6, max_iter: int = 100, want_scores: bool = True):
    """One fit on the resident data.  Returns a dict of arrays in path-LV / ODM order."""
    P, L, N = model.P, model.L, data.N
    out = dict(weights=np.empty(P), loadings=np.empty(P), r_squared=np.empty(L), path_coefficients=np.empty((L, L)),
               total_effects=np.empty((L, L)), crossloadings=np.empty((P, L)),
               scores=np.empty((N, L)) if want_scores else None)
    iters, status = ctypes.c_int32(0), ctypes.c_int32(0)
    _check(load().plspm_fit(model._h, data._h, scheme_id(scheme), float(tol), int(max_iter), _ptr(out["weights"], _c_dp),
                            _ptr(out["loadings"], _c_dp), _ptr(out["r_squared"], _c_dp),
                            _ptr(out["path_coefficients"], _c_dp), _ptr(out["total_effects"], _c_dp),
                            _ptr(out["crossloadings"], _c_dp), _ptr(out["scores"], _c_dp), ctypes.byref(iters),
                            ctypes.byref(status)))
    out["iterations"], out["status"] = int(iters.value), int(status.value)
    return out


def bootstrap(model: Model, data: Data, scheme, rep_begin: int, rep_count: int, seed: int = 0, idx=None,
              tol: float = 1e-6, max_iter: int = 100, out_device_ptr: int = 0):
    """Replicates [rep_begin, rep_begin + rep_count).  Returns (rows [rep_count, n_out] or None when writing to a
    device buffer, status [rep_count], iters [rep_count])."""
    status = np.zeros(rep_count, dtype=np.int32)
    iters = np.zeros(rep_count, dtype=np.int32)
    idx_a = None
    if idx is not None:
        idx_a = np.ascontiguousarray(idx, dtype=np.int32)
        if idx_a.shape != (rep_count, data.N):
            raise ValueError("idx must be [rep_count, N]")
    if out_device_ptr:
        rows, optr, dev = None, ctypes.c_void_p(out_device_ptr), 1
    else:
        rows = np.empty((rep_count, getattr(data, "rows_model", model).n_out), dtype=np.float64)
        optr, dev = rows.ctypes.data_as(ctypes.c_void_p), 0
    _check(load().plspm_bootstrap(model._h, data._h, scheme_id(scheme), float(tol), int(max_iter), int(rep_begin),
                                  int(rep_count), int(seed), _ptr(idx_a, _c_i32p), optr, dev, _ptr(status, _c_i32p),
                                  _ptr(iters, _c_i32p)))
    return rows, status, iters


def bootstrap_host(model: Model, X: np.ndarray, scheme, rep_begin: int, rep_count: int, seed: int = 0, idx=None,
                   tol: float = 1e-6, max_iter: int = 100, out=None):
    """End-to-end call with a HOST matrix: upload + bootstrap + release inside the library."""
    assert X.dtype == np.float64 and X.flags.c_contiguous and X.shape[1] == model.P
    status = np.zeros(rep_count, dtype=np.int32)
    iters = np.zeros(rep_count, dtype=np.int32)
    rows = out if out is not None else np.empty((rep_count, model.n_out), dtype=np.float64)
    idx_a = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    _check(load().plspm_bootstrap_host(model._h, X.ctypes.data_as(ctypes.c_void_p), int(X.shape[0]), int(X.shape[1]),
                                       scheme_id(scheme), float(tol), int(max_iter), int(rep_begin), int(rep_count),
                                       int(seed), _ptr(idx_a, _c_i32p), _ptr(rows, _c_dp), _ptr(status, _c_i32p),
                                       _ptr(iters, _c_i32p)))
    return rows, status, iters


def resample_indices(seed: int, replicate: int, N: int) -> np.ndarray:
    out = np.empty(N, dtype=np.int32)
    _check(load().plspm_resample_indices(int(seed), int(replicate), int(N), _ptr(out, _c_i32p)))
    return out


STAGES = ("counts", "gram", "reduce", "solve", "scores", "upload", "colsum", "cross", "scoregen", "conv", "gram_i8", "finalize")


def profile_reset():
    _check(load().plspm_profile_reset())


def profile_get():
    ms = np.zeros(12, dtype=np.float64)
    n = np.zeros(12, dtype=np.int64)
    _check(load().plspm_profile_get(_ptr(ms, _c_dp), _ptr(n, _c_i64p)))
    return {s: (float(ms[i]), int(n[i])) for i, s in enumerate(STAGES)}


def redo_count() -> int:
    """Replicates redone with exact cross moments after an undecided low-precision sign vote."""
    n = ctypes.c_int64(0)
    _check(load().plspm_redo_count(ctypes.byref(n)))
    return int(n.value)


def pool_trim():
    """Returns every cached (released) device buffer of the library to the CUDA driver."""
    _check(load().plspm_pool_trim())


def pool_set_limit(nbytes: int):
    """Upper bound on the bytes of released device buffers the library keeps for the next call."""
    _check(load().plspm_pool_set_limit(int(nbytes)))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """NumPy array backed by page-locked host memory (freed when the array is collected)."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    _check(load().plspm_host_alloc(ctypes.byref(p), max(nbytes, 1)))
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p.value
    return arr


_PINNED = {}


def pinned_free(arr: np.ndarray):
    p = _PINNED.pop(arr.__array_interface__["data"][0], None)
    if p:
        load().plspm_host_free(ctypes.c_void_p(p))
