"""CPU oracle for the TVFC summary measures.  TEST INFRASTRUCTURE ONLY (see oracle/wishart_oracle.py header).

NumPy restatement of ``fcest/helpers/summary_measures.py:16-126``: nanmean / nanvar / nanstd over the time axis,
the rate of change (``:83-126``) and the AR(1) coefficient per edge (``:50-80``).

PARITY: mean / variance / std / rate of change are PINNED -- the reference functions run in the build container
(only their statsmodels import is stubbed) and ``tests/golden/make_golden.py`` stores their outputs
(``tests/golden/summary.npz``).  The AR(1) route is UNPINNED: it calls ``statsmodels.tsa.ar_model.AutoReg``
(un-vendored, not installed here; ``setup.py`` leaves the version open), whose ``lags=1`` fit with the default
constant trend is the ordinary least-squares regression of y[1:] on [1, y[:-1]]; that published definition is
restated with ``np.linalg.lstsq``.
"""
from __future__ import annotations

import numpy as np


def summarize(c: np.ndarray, metric: str) -> np.ndarray:
    if metric == "mean":
        return np.nanmean(c, axis=0)
    if metric == "variance":
        return np.nanvar(c, axis=0)
    if metric == "std":
        return np.nanstd(c, axis=0)
    if metric == "rate_of_change":
        return rate_of_change(c)
    if metric == "ar1":
        return ar1(c)
    raise ValueError(metric)


def rate_of_change(c: np.ndarray) -> np.ndarray:
    """summary_measures.py:83-126."""
    out = np.zeros(c.shape[1:])
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(c.shape[0] - 1):
            r = np.abs(c[t + 1] / c[t] - 1)
            r[r == np.inf] = 1
            r[r > 10] = 1
            out += r
    return out / (c.shape[0] - 1)


def ar1(c: np.ndarray) -> np.ndarray:
    """summary_measures.py:50-80 with AutoReg(lags=1).fit().params[-1] = OLS slope of y[1:] on [1, y[:-1]]."""
    d = c.shape[1]
    out = np.zeros((d, d))
    for i, j in zip(*np.tril_indices(d, k=-1)):
        y = c[:, i, j]
        X = np.stack([np.ones(len(y) - 1), y[:-1]], axis=1)
        beta = np.linalg.lstsq(X, y[1:], rcond=None)[0]
        out[i, j] = out[j, i] = beta[-1]
    return out
