"""Plspm: the public entry point (reference plspm/plspm.py:26-169), same signature and getters.

The weight estimation, inner-model coefficients, loadings and crossloadings come from the CUDA
engine (one upload, one Gram pass, one solver CTA); the remaining getters are O(P) / O(L^2)
host post-processing of those outputs.
"""
import numpy as np
import pandas as pd

import plspm.config as c
import plspm.inner_model as im
import plspm.inner_summary as pis
import plspm.outer_model as om
import plspm.weights as w
from plspm.bootstrap import Bootstrap
from plspm.estimator import Estimator
from plspm.scheme import Scheme
from plspm.unidimensionality import Unidimensionality


class Plspm:
    """Estimates path models with latent variables using the partial least squares algorithm."""

    def __init__(self, data: pd.DataFrame, config: c.Config, scheme: Scheme = Scheme.CENTROID,
                 iterations: int = 100, tolerance: float = 0.000001, bootstrap: bool = False,
                 bootstrap_iterations: int = 100, processes: int = 2, bootstrap_seed: int = None,
                 bootstrap_indices: np.ndarray = None):
        # argument rules of the reference (plspm.py:54-61)
        if iterations < 100:
            iterations = 100
        assert tolerance > 0
        assert scheme in Scheme
        if bootstrap_iterations < 10:
            bootstrap_iterations = 100
        assert processes > 0
        assert bootstrap_iterations % processes == 0

        estimator = Estimator(config)
        filtered = config.filter(data)
        n = filtered.shape[0]
        correction = np.sqrt(n / (n - 1))
        calculator = w.WeightsCalculatorFactory(config, iterations, tolerance, correction, scheme)
        _, scores, weights = estimator.estimate(calculator, filtered, want_final_data=False)
        config = estimator.config()
        session, res = estimator.last_result()

        self.__scores = scores
        self.__inner_model = im.InnerModel(config.path(), scores, res["path_coefficients"], res["r_squared"])
        self.__outer_model = om.OuterModel(weights, res["loadings"], res["crossloadings"], session.lvs,
                                           session.blocks, self.__inner_model.r_squared())
        self.__inner_summary = pis.InnerSummary(config, self.__inner_model.r_squared(),
                                                self.__inner_model.r_squared_adj(), self.__outer_model.model())
        self.__unidimensionality = Unidimensionality(config, filtered, correction)
        self.__iterations = res["iterations"]
        self.__bootstrap = None
        if bootstrap:
            if n < 10:
                raise Exception("Bootstrapping could not be performed, at least 10 observations are required.")
            self.__bootstrap = Bootstrap(config, filtered, self.__inner_model, self.__outer_model, calculator,
                                         bootstrap_iterations, processes, seed=bootstrap_seed,
                                         indices=bootstrap_indices)

    def scores(self) -> pd.DataFrame:
        """Latent variable scores, one column per latent variable, indexed like the input data."""
        return self.__scores

    def outer_model(self) -> pd.DataFrame:
        """weight, loading, communality and redundancy of every manifest variable."""
        return self.__outer_model.model()

    def inner_model(self) -> pd.DataFrame:
        """estimate, std error, t and p>|t| for every path into an endogenous latent variable."""
        return self.__inner_model.inner_model()

    def path_coefficients(self) -> pd.DataFrame:
        """Path coefficient matrix, shaped like the path matrix given to Config."""
        return self.__inner_model.path_coefficients()

    def crossloadings(self) -> pd.DataFrame:
        """Correlations of every manifest variable (rows) with every latent variable score (columns)."""
        return self.__outer_model.crossloadings()

    def inner_summary(self) -> pd.DataFrame:
        """type, R squared, block communality, mean redundancy and AVE per latent variable."""
        return self.__inner_summary.summary()

    def goodness_of_fit(self) -> float:
        return self.__inner_summary.goodness_of_fit()

    def effects(self) -> pd.DataFrame:
        """direct, indirect and total effects for each path."""
        return self.__inner_model.effects()

    def unidimensionality(self) -> pd.DataFrame:
        """Cronbach's alpha, Dillon-Goldstein's rho and the first two eigenvalues per block."""
        return self.__unidimensionality.summary()

    def bootstrap(self) -> Bootstrap:
        if self.__bootstrap is None:
            raise Exception("To perform bootstrap validation, set the parameter bootstrap to True when calling Plspm")
        return self.__bootstrap

    def iterations(self) -> int:
        """Number of outer iterations the estimation used (not in the reference)."""
        return self.__iterations
