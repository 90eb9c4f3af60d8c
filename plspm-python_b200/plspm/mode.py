"""Outer-weight modes (reference plspm/mode.py:64-69).

In the reference each enum value is a strategy object whose `outer_weights_metric` runs per LV and
per iteration on the host (mode.py:28-29, 50-52).  Here the enum only names the mode: the update
itself (Mode A: block covariance with the inner proxy; Mode B: per-block least squares from a
Cholesky factor of the block covariance) runs inside the CUDA solver (csrc/solver_core.h).
"""
from enum import Enum


class _ModeTag:
    def __init__(self, letter: str, engine_id: int):
        self.letter, self.engine_id = letter, engine_id

    def __repr__(self):
        return "Mode(%s)" % self.letter

    def __eq__(self, other):
        return isinstance(other, _ModeTag) and other.letter == self.letter

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.letter)


class Mode(Enum):
    """Whether a latent variable is reflective (mode A) or formative (mode B) w.r.t. its manifest variables."""
    A = _ModeTag("A", 0)
    B = _ModeTag("B", 1)
