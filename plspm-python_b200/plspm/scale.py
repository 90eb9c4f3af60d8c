"""Measurement scales of non-metric data (reference plspm/scale.py:91-104).

Kept importable for API compatibility.  The non-metric (optimal scaling) path is outside the
accelerated hot path: configuring any scale makes Plspm raise
NotImplementedError instead of silently running something else.
"""
from enum import Enum


class Scale(Enum):
    NUM = "NUM"
    RAW = "RAW"
    ORD = "ORD"
    NOM = "NOM"
