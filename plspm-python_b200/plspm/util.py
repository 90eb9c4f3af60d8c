"""Host-side helpers (reference plspm/util.py): treatment, imputation, outer design matrix, rank / dummy / group
means of the ordinal and nominal scales (not accelerated; kept for API completeness), topological sort."""
import collections

import numpy as np
import pandas as pd


def treat(data: pd.DataFrame, center: bool = True, scale: bool = True, scale_values=None) -> pd.DataFrame:
    """Centre / scale a DataFrame (reference util.py:21-40)."""
    out = data
    if center:
        out = out - out.mean()
    if scale:
        out = out / (scale_values if scale_values else out.std())
    return out


def sort_cols(data: pd.DataFrame) -> pd.DataFrame:
    return data.reindex(sorted(data.columns), axis=1)


def impute(data: pd.DataFrame) -> pd.DataFrame:
    """Mean imputation of missing metric values (reference util.py:61-68)."""
    return data.fillna(data.mean(skipna=True))


def list_to_dummy(data: dict) -> pd.DataFrame:
    """Outer design matrix: rows = MVs (dict order), columns = LVs (reference util.py:71-77)."""
    rows = [mv for mvs in data.values() for mv in mvs]
    out = pd.DataFrame(0.0, index=rows, columns=list(data.keys()))
    for lv, mvs in data.items():
        out.loc[mvs, lv] = 1.0
    return out


def treat_numpy(data: np.ndarray) -> np.ndarray:
    """Centre and scale (sd with ddof = 1) one NumPy column, ignoring NaNs (reference util.py:43-53)."""
    centred = data - np.nanmean(data)
    return centred / np.nanstd(centred, axis=0, ddof=1)


def rank(data: pd.Series) -> pd.Series:
    """Dense rank of the distinct values, 1 = smallest, ties share a rank (reference util.py:80-86).  Used by
    the ordinal / nominal scales, which this package does not accelerate; kept for API completeness."""
    order = {v: float(r) for r, v in enumerate(sorted(pd.unique(data.dropna())), start=1)}
    return data.map(order).astype(float)


def dummy(data: pd.Series) -> pd.DataFrame:
    """Indicator matrix of ranked data: column r is 1 where data == r, r = 1..#distinct (reference util.py:89-95)."""
    levels = range(1, data.unique().size + 1)
    return pd.DataFrame({r: (data == r).astype(int) for r in levels}, index=data.index, columns=list(levels))


def groupby_mean(data: np.ndarray) -> np.ndarray:
    """Row 0 of `data` holds group keys, row 1 values: returns [sorted keys; group means] (reference
    util.py:98-112), i.e. pandas' groupby(...).mean() for a 2 x n array."""
    keys, inverse = np.unique(data[0], return_inverse=True)
    sums = np.bincount(inverse, weights=data[1], minlength=keys.size)
    counts = np.bincount(inverse, minlength=keys.size)
    return np.vstack([keys.astype(np.float64), sums / counts])


class TopoSort:
    """Kahn topological sort with the reference's tie-breaking (util.py:127-160): sources are taken
    from the END of the ready queue, so the LV order of Structure.path() matches the reference.
    Unlike the reference, order() does not destroy its own state (quirk Q5)."""

    def __init__(self):
        self._indeg = collections.Counter()
        self._children = {}
        self._edges = []

    def append(self, src: str, dest: str):
        self._edges.append((src, dest))
        self._indeg[dest] += 1
        self._indeg[src] += 0
        self._children.setdefault(src, [])
        self._children.setdefault(dest, [])
        self._children[src].append(dest)

    def order(self):
        indeg = dict(self._indeg)
        ready = collections.deque(v for v in indeg if indeg[v] == 0)
        out = []
        while ready:
            v = ready.pop()
            out.append(v)
            for ch in self._children[v]:
                indeg[ch] -= 1
                if indeg[ch] == 0:
                    ready.append(ch)
        if any(d != 0 for d in indeg.values()):
            raise ValueError("Structural graph contains cycles.")
        return out

    def elements(self):
        return self._edges
