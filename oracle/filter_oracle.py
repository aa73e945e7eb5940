"""CPU oracle for the high-pass filter in front of the sliding-window estimator.  TEST INFRASTRUCTURE ONLY
(see oracle/wishart_oracle.py header): only tests/, __graft_entry__.smoke() and bench tools may import it.

NumPy / pure-Python restatement of ``highpass_filter_data`` (reference ``fcest/helpers/filtering.py:12-116``)
including the third-party arithmetic it delegates to -- SciPy 1.18's ``scipy.signal.butter(5, Wn, 'high')``
(buttap -> lp2hp_zpk -> bilinear_zpk at fs = 2 -> zpk2tf, scipy/signal/_filter_design.py) and
``scipy.signal.filtfilt`` with its defaults (odd extension by padlen = 3 * max(len(a), len(b)) samples,
``lfilter_zi`` initial state scaled by the first sample, forward pass, reversed pass;
scipy/signal/_signaltools.py), whose inner loop is the direct-form-II-transposed recurrence
``y = z0 + b0 x;  z_i = z_{i+1} + x b_{i+1} - y a_{i+1}``.

PARITY PINNED: ``filtering.py`` imports only numpy and scipy, so the reference function itself runs in the build
container; ``tests/golden/make_golden.py`` stores its outputs (``tests/golden/filtering.npz``) and
``tests/test_oracle_cpu.py`` holds this restatement to them bit for bit.
"""
from __future__ import annotations

import numpy as np


def butter_highpass(cutoff_low: float, nyquist_freq: float, order: int = 5):
    """filtering.py:75-95 -> scipy.signal.butter(order, cutoff_low / nyquist_freq, btype='high')."""
    from scipy.signal import butter  # the third-party routine the reference calls
    return butter(order, cutoff_low / nyquist_freq, btype="high", analog=False)


def lfilter_zi(b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """scipy.signal.lfilter_zi for a[0] == 1."""
    y_inf = b.sum() / a.sum()
    return np.cumsum((b - y_inf * a)[::-1])[::-1][1:]


def lfilter_df2t(b, a, x, z):
    """Direct form II transposed, one sample at a time; x (L,) or (L, C), z (order,) or (order, C)."""
    n = len(b) - 1
    z = np.array(z, dtype=np.float64, copy=True)
    y = np.empty_like(x)
    for k in range(x.shape[0]):
        xn = x[k]
        yn = z[0] + b[0] * xn
        for i in range(n - 1):
            z[i] = z[i + 1] + xn * b[i + 1] - yn * a[i + 1]
        z[n - 1] = xn * b[n] - yn * a[n]
        y[k] = yn
    return y


def filtfilt_columns(b, a, x):
    """scipy.signal.filtfilt(b, a, column) for every column of x (N, C) (or a single (N,) series)."""
    x = np.asarray(x, dtype=np.float64)
    edge = 3 * max(len(a), len(b))
    if x.shape[0] <= edge:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {edge}.")
    ext = np.concatenate((2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-(edge + 2):-1]))
    zi = lfilter_zi(b, a)
    if x.ndim == 2:
        zi = zi[:, None]
    y = lfilter_df2t(b, a, ext, zi * ext[0])
    y = lfilter_df2t(b, a, y[::-1].copy(), zi * y[-1])[::-1]
    return y[edge:-edge]


def highpass_filter_data(y_observed, window_length: int, repetition_time: float):
    """filtering.py:12-44."""
    nyquist = 1.0 / repetition_time / 2
    b, a = butter_highpass(1 / (window_length * repetition_time), nyquist)
    return filtfilt_columns(b, a, y_observed)
