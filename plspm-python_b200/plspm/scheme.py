"""Inner weighting schemes (reference plspm/scheme.py:57-63).

The reference's enum values compute E = f(corr(Y)) with NumPy / statsmodels on the host
(scheme.py:27-28 centroid, 36-37 factorial, 45-54 path).  Here the enum names the scheme and the
computation happens on the L x L score correlations held in shared memory by the CUDA solver.
"""
from enum import Enum


class _SchemeTag:
    def __init__(self, letter: str, engine_id: int):
        self.letter, self.engine_id = letter, engine_id

    def __repr__(self):
        return "Scheme(%s)" % self.letter

    def __eq__(self, other):
        return isinstance(other, _SchemeTag) and other.letter == self.letter

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.letter)


class Scheme(Enum):
    """The scheme used to calculate inner weights."""
    CENTROID = _SchemeTag("C", 0)
    PATH = _SchemeTag("P", 2)
    FACTORIAL = _SchemeTag("F", 1)
