"""Host-side helpers (reference plspm/util.py).  Only the pieces the metric path needs."""
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
