"""Model specification: Structure, MV, Config (reference plspm/config.py).

Pure bookkeeping on the host.  Same constructor signatures, validation rules and error messages as
the reference so that its configuration tests read the same here; the arithmetic of `treat` is done
on the device (the metric branch here exists for API compatibility and for callers that want the
treated DataFrame).
"""
import itertools

import numpy as np
import pandas as pd

import plspm.util as util
from plspm.mode import Mode
from plspm.scale import Scale


class Structure:
    """Builds the lower-triangular path matrix from `add_path` calls (reference config.py:24-58)."""

    def __init__(self, path: pd.DataFrame = None):
        self._sorter = util.TopoSort()
        if path is not None:
            rows, cols = np.where(path.values == 1)
            for r, c in zip(rows, cols):
                self.add_path([path.columns[c]], [path.index[r]])

    def add_path(self, source: list, target: list):
        if len(source) != 1 and len(target) != 1:
            raise ValueError("Either source or target must be a list containing a single entry")
        if len(source) == 0 or len(target) == 0:
            raise ValueError("Both source and target must contain at least one entry")
        for src, dst in itertools.product(source, target):
            self._sorter.append(src, dst)

    def path(self) -> pd.DataFrame:
        order = self._sorter.order()
        matrix = pd.DataFrame(np.zeros((len(order), len(order)), int), columns=order, index=order)
        for src, dst in self._sorter.elements():
            matrix.at[dst, src] = 1
        return matrix


class MV:
    """A manifest variable: column name and (for non-metric data) its scale (reference config.py:60-81)."""

    def __init__(self, name: str, scale: Scale = None):
        self._name, self._scale = name, scale

    def name(self):
        return self._name

    def scale(self):
        return self._scale


class Config:
    """The model to estimate (reference config.py:84-319)."""

    def __init__(self, path: pd.DataFrame, scaled: bool = True, default_scale: Scale = None):
        if not isinstance(path, pd.DataFrame):
            raise TypeError("Path argument must be a Pandas DataFrame")
        if path.shape[0] != path.shape[1]:
            raise ValueError("Path argument must be a square matrix")
        values = np.asarray(path.values)
        if not np.array_equal(values, np.tril(values)):
            raise ValueError("Path argument must be a lower triangular matrix")
        if not np.isin(values, [0, 1]).all():
            raise ValueError("Path matrix element values may only be in [0, 1]")
        if list(path.columns.values) != list(path.index.values):
            raise ValueError("Path matrix must have matching row and column index names")
        self._path = path
        self._scaled = scaled
        self._default_scale = default_scale
        self._modes, self._blocks, self._hoc, self._mv_scales, self._dummies = {}, {}, {}, {}, {}
        self._metric = True
        self._missing = False

    def clone(self):
        other = Config(self._path, self._scaled, self._default_scale)
        other._modes, other._blocks = dict(self._modes), {k: list(v) for k, v in self._blocks.items()}
        other._hoc, other._mv_scales = dict(self._hoc), dict(self._mv_scales)
        other._dummies = dict(self._dummies)
        other._metric, other._missing = self._metric, self._missing
        return other

    # ---- accessors -------------------------------------------------------------------------
    def path(self):
        return self._path

    def odm(self, path: pd.DataFrame):
        return util.list_to_dummy({lv: self._blocks[lv] for lv in list(path)})

    def mv_index(self, lv, mv):
        return self._blocks[lv].index(mv)

    def mvs(self, lv):
        return self._blocks[lv]

    def hoc(self):
        return self._hoc

    def mode(self, lv: str):
        return self._modes[lv]

    def metric(self):
        return self._metric

    def scaled(self):
        return self._scaled

    def numeric(self):
        """Not in the reference: True for a non-metric configuration whose manifest variables all carry
        Scale.NUM or Scale.RAW -- the subset of the non-metric path that runs on the device."""
        kinds = set(self._mv_scales.values())
        return (not self._metric) and None not in kinds and kinds <= {Scale.NUM, Scale.RAW}

    def check_scales(self):
        """Scale bookkeeping of the non-metric branch (reference config.py:306-313)."""
        if None in self._mv_scales.values():
            raise TypeError("If you supply a scale for any MV, you must either supply a scale for all of them or specify a default scale.")
        kinds = set(self._mv_scales.values())
        if kinds == {Scale.RAW}:
            self._scaled = False
        if kinds == {Scale.RAW, Scale.NUM}:
            self._scaled = True
            self._mv_scales = dict.fromkeys(self._mv_scales, Scale.NUM)
        if not kinds <= {Scale.NUM, Scale.RAW}:
            raise NotImplementedError("ordinal / nominal scales (Scale.ORD, Scale.NOM) are outside the accelerated "
                                      "path of plspm_b200")

    def scale(self, mv: str):
        return self._mv_scales[mv]

    def dummies(self, mv: str):
        return self._dummies[mv]

    def missing(self):
        return self._missing

    # ---- model building ---------------------------------------------------------------------
    def add_lv(self, lv_name: str, mode: Mode, *mvs: MV):
        assert mode in Mode
        hoc_members = [lv for members in self._hoc.values() for lv in members]
        if lv_name not in self._path and lv_name not in hoc_members:
            raise ValueError("Latent variable " + lv_name + " is not listed in the outer model paths or higher order constructs.")
        self._modes[lv_name] = mode
        self._blocks[lv_name] = []
        for mv in mvs:
            if mv.name() in self._mv_scales:
                raise ValueError("You can only specify a column once. You can specify a higher order construct with `add_higher_order(...)`")
            if mv.name() in list(self._path):
                raise ValueError("You cannot specify MVs with the same name as LVs.")
            self._blocks[lv_name].append(mv.name())
            scale = mv.scale() if mv.scale() is not None else self._default_scale
            self._mv_scales[mv.name()] = scale
            if scale is not None:
                self._metric = False

    def remove_lv(self, lv_name: str):
        self._blocks.pop(lv_name)
        self._modes.pop(lv_name)

    def add_higher_order(self, hoc_name: str, mode: Mode, lvs: list):
        assert mode in Mode
        if hoc_name not in self._path:
            raise ValueError("Path matrix does not contain reference to higher order construct " + hoc_name)
        self._modes[hoc_name] = mode
        self._hoc[hoc_name] = lvs

    def add_lv_with_columns_named(self, lv_name: str, mode: Mode, data: pd.DataFrame, col_name_starts_with: str,
                                  default_scale: Scale = None):
        mvs = [MV(col, default_scale) for col in list(data) if col.startswith(col_name_starts_with)]
        if not mvs:
            raise ValueError("No columns were found in the data starting with " + col_name_starts_with)
        self.add_lv(lv_name, mode, *mvs)

    # ---- data ----------------------------------------------------------------------------------
    def filter(self, data: pd.DataFrame) -> pd.DataFrame:
        """Keeps the configured MV columns (insertion order) and drops rows where a whole block is
        missing (reference config.py:247-285)."""
        hoc_members = [lv for members in self._hoc.values() for lv in members]
        path_lvs = [lv for lv in list(self.path()) + hoc_members if lv not in self._hoc]
        if set(self._blocks) != set(path_lvs):
            raise ValueError(
                "The Path matrix supplied does not specify the same latent variables as you added when configuring manifest variables." +
                " Path: " + ", ".join(path_lvs) + " LVs: " + ", ".join(set(self._blocks)))
        absent = set(self._mv_scales).difference(set(data))
        if absent:
            raise ValueError("The following manifest variables you configured are not present in the data set: " + ", ".join(absent))
        data = data[list(self._mv_scales)]
        if not all(pd.api.types.is_numeric_dtype(dt) for dt in data.dtypes):
            raise ValueError("Data must only contain numeric values. Please convert any categorical data into numerical values.")
        self._missing = bool(data.isnull().values.any())
        if self._missing:
            drop = np.zeros(len(data.index), dtype=bool)
            for lv in list(self.path()):
                if lv in self._blocks:
                    drop |= data[self._blocks[lv]].isnull().values.all(axis=1)
            data = data.loc[~drop]
        return data

    def treat(self, data: pd.DataFrame) -> pd.DataFrame:
        """Centres (and, if `scaled`, divides by ONE pooled scalar) metric data (reference config.py:287-305);
        standardises every column of non-metric NUM / RAW data (config.py:306-314)."""
        if not self._metric:
            self.check_scales()
            if self._missing:
                raise NotImplementedError("non-metric data with missing values are outside the accelerated path")
            n = data.shape[0]
            return util.treat(data) / np.sqrt((n - 1) / n)  # config.py:314: unit population variance
        metric = util.impute(data) if self._missing else data
        if self._scaled:
            n = metric.shape[0]
            pooled = float(np.std(metric.values.astype(np.float64).reshape(-1), ddof=1)) * np.sqrt((n - 1) / n)
            return util.treat(metric, scale_values=pooled)
        return util.treat(metric, scale=False)
