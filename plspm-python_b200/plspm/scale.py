"""Measurement scales of non-metric data (reference plspm/scale.py:91-104).

NUM / RAW configurations run the non-metric estimator on the device (csrc/solver_num.h); ORD / NOM (optimal
scaling, scale.py:42-89) and non-metric data with missing values take the reference-style host path of
plspm/nonmetric_host.py.
"""
from enum import Enum


class Scale(Enum):
    NUM = "NUM"
    RAW = "RAW"
    ORD = "ORD"
    NOM = "NOM"
