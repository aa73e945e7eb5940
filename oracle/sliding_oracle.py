"""CPU oracle for the sliding-window path.  TEST INFRASTRUCTURE ONLY (see oracle/wishart_oracle.py header).

NumPy restatement of ``SlidingWindows.overlapping_windowed_cov_estimation`` and the cross-validated
window-length search (reference ``fcest/models/sliding_windows.py:109-174`` and ``:199-350``,
``fcest/helpers/array_operations.py:19-53``), including the third-party arithmetic they delegate to:
``np.cov`` (two-pass, multiply by 1/(w-1); numpy/lib/_function_base_impl.py:2890-2914) and
``scipy.stats.multivariate_normal.logpdf(..., allow_singular=True)`` (eigh-based pseudo-determinant,
eigenvalues <= 1e6 * eps * max|s| treated as zero, -inf outside the support;
scipy/stats/_multivariate.py:71-102,168-246,566-619).

PARITY PINNED: unlike the Wishart oracle this one is checked against the reference itself -- the
reference's SlidingWindows class is importable in the build container and
``tests/golden/make_golden.py`` stores its outputs as fixtures; ``tests/test_oracle_cpu.py`` holds this
restatement to those fixtures.
"""
from __future__ import annotations

import numpy as np

LOG_2PI = float(np.log(2 * np.pi))


def np_cov_rows(rows: np.ndarray) -> np.ndarray:
    """np.cov(rows.T): rows (w, D) -> (D, D), two-pass centred, scaled by 1/(w-1)."""
    w = rows.shape[0]
    avg = rows.mean(axis=0)
    xc = rows - avg
    c = xc.T @ xc
    c *= np.true_divide(1, w - 1)
    return c


def correlation_from_covariance(c: np.ndarray) -> np.ndarray:
    """array_operations.py:37-53."""
    v = np.sqrt(np.diag(c))
    with np.errstate(divide="ignore", invalid="ignore"):
        r = c / np.outer(v, v)
    r[c == 0] = 0
    return r


def overlapping_windowed_cov(y: np.ndarray, window_length: int, step_size: int = 1,
                             connectivity_metric: str = "covariance") -> np.ndarray:
    """sliding_windows.py:143-174 without the optional high-pass filter (repetition_time=None)."""
    n, d = y.shape
    pad = int(np.floor(window_length / 2))
    yp = np.concatenate((np.zeros((pad, d)), y, np.zeros((pad, d))))
    out = np.zeros((n, d, d))
    for i in range(int(n / step_size)):
        start = int(i * step_size)
        out[i] = np_cov_rows(yp[start:start + window_length])
    if connectivity_metric == "correlation":
        out = np.array([correlation_from_covariance(c) for c in out])
    return out


def cv_bounds(num_time_steps: int, repetition_time: float = 1.0):
    """sliding_windows.py:289-304 (the TR branch is always taken: TR=None is replaced by 1.0 at :63-66)."""
    lo = int(np.ceil(20.0 / repetition_time))
    hi = int(min(int(np.ceil(num_time_steps * repetition_time / 2 / repetition_time)),
                 int(np.ceil(180.0 / repetition_time))))
    if lo % 2 == 0:
        lo += 1
    if hi % 2 == 0:
        hi += 1
    return lo, hi


def mvn_logpdf_allow_singular(x: np.ndarray, cov: np.ndarray) -> float:
    """scipy.stats.multivariate_normal.logpdf(x, mean=0, cov=cov, allow_singular=True)."""
    s, u = np.linalg.eigh(cov)
    eps = 1e6 * np.finfo(np.float64).eps * np.max(np.abs(s))
    if np.min(s) < -eps:
        raise ValueError("The input matrix must be symmetric positive semidefinite.")
    keep = s > eps
    d = s[keep]
    s_pinv = np.array([0.0 if abs(v) <= eps else 1.0 / v for v in s])
    U = u * np.sqrt(s_pinv)
    rank = len(d)
    log_pdet = np.sum(np.log(d))
    maha = np.sum(np.square(x @ U))
    out = -0.5 * (rank * LOG_2PI + log_pdet + maha)
    if rank < len(s):
        null_basis = u[:, ~keep]
        residual = np.linalg.norm(x @ null_basis)
        if not residual < 1e3 * eps:
            return -np.inf
    return float(out)


def cv_grid(num_time_steps: int, step_w: int = 2, step_t: int = 1, repetition_time: float = 1.0):
    lo, hi = cv_bounds(num_time_steps, repetition_time)
    windows = np.arange(lo, hi, step_w)
    t_lo = int((hi - 1) / 2)
    t_hi = num_time_steps - int((hi + 1) / 2)
    locs = np.arange(t_lo, t_hi, step_t)
    return windows, locs


def cross_validated_table(y: np.ndarray, step_w: int = 2, step_t: int = 1, windows=None, locs=None) -> tuple:
    """sliding_windows.py:219-272 -> (window_lengths (W,), eval_locations (T,), table (T, W))."""
    n, d = y.shape
    if windows is None or locs is None:
        windows, locs = cv_grid(n, step_w, step_t)
    table = np.zeros((len(locs), len(windows)))
    for wi, w in enumerate(windows):
        for ti, t in enumerate(locs):
            start = int(t - np.floor(w / 2))
            end = int(t + np.ceil(w / 2))
            win = y[:end][start:]
            win = np.delete(win, int((w - 1) / 2), axis=0)
            table[ti, wi] = mvn_logpdf_allow_singular(y[t], np_cov_rows(win))
    return windows, locs, table


def optimal_window_length(windows: np.ndarray, table: np.ndarray) -> int:
    """sliding_windows.py:331-350: column mean -> argmax."""
    return int(windows[int(np.argmax(table.mean(axis=0)))])
