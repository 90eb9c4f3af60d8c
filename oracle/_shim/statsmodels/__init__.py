"""Minimal stand-in for the `statsmodels` package (TEST INFRASTRUCTURE ONLY).

statsmodels is an un-pinned third-party dependency of the reference
(/root/reference/requirements.txt:4, setup.py:36) that is not installed in this
image and cannot be fetched (no network).  The reference only uses
`statsmodels.api.OLS(...).fit()` and `statsmodels.api.add_constant`
(scheme.py:50, weights.py:140-142, inner_model.py:76-77).  This shim restates
the published semantics of those two calls (least squares through the
Moore-Penrose pseudo-inverse) so that the *unmodified* reference algorithm can
be executed here to generate golden vectors.  It is never imported by the
product package.
"""
