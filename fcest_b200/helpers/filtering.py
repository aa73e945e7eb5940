"""``fcest.helpers.filtering`` mirror: Butterworth high-pass, zero-phase, every column on the GPU.

Reference: ``fcest/helpers/filtering.py:12-116`` (a Python loop over the D columns, each through
``scipy.signal.butter`` + ``scipy.signal.filtfilt``).  Here the filter is designed once on the host
(eleven numbers; the design below follows SciPy's published ``butter`` route -- analogue Butterworth
prototype, low-pass -> high-pass, bilinear transform at fs = 2, zeros / poles -> polynomials -- with the same
NumPy primitives in the same order, so the coefficients are bit-identical to SciPy's) and applied to all
columns by one launch of ``fcest_filtfilt`` (``csrc/filter.cu``).  Same function names and arguments as the
reference; NumPy in, NumPy out.  Device tensors are accepted too (and returned as device tensors), which is
what ``SlidingWindows.overlapping_windowed_cov_estimation`` uses to keep the data in HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib, ops

__all__ = [
    "highpass_filter_data",
    "butter_highpass_filter",
    "_butter_highpass",
    "_compute_lower_frequency_cutoff",
]


def highpass_filter_data(y_observed, window_length: int, repetition_time: float):
    """Remove frequencies below 1 / window length (in seconds) from every column of the (N, D) array
    (filtering.py:12-44)."""
    sampling_rate = 1.0 / repetition_time
    nyquist_frequency = sampling_rate / 2
    return butter_highpass_filter(
        y_observed,
        cutoff_low=_compute_lower_frequency_cutoff(window_length, repetition_time),
        nyquist_freq=nyquist_frequency,
    )


def butter_highpass_filter(data, cutoff_low, nyquist_freq):
    """Zero-phase high-pass of a (N,) series or of every column of a (N, C) array (filtering.py:47-72)."""
    b, a = _butter_highpass(cutoff_low, nyquist_freq=nyquist_freq)
    return _filtfilt_columns(b, a, data)


def _butter_highpass(cutoff_low, nyquist_freq, order=5) -> tuple[np.ndarray, np.ndarray]:
    """Digital Butterworth high-pass (b, a), ``scipy.signal.butter(order, cutoff_low / nyquist_freq, 'high')``
    (filtering.py:75-95)."""
    wn = cutoff_low / nyquist_freq
    if not 0.0 < wn < 1.0:
        raise ValueError("Digital filter critical frequencies must be 0 < Wn < 1")
    # analogue prototype: poles on the unit circle, left half plane (the middle one exactly real)
    m = np.arange(-order + 1, order, 2, dtype=np.float64)
    p = -np.exp(1j * np.pi * m / (2 * order))
    # pre-warp for the bilinear transform at fs = 2
    fs = 2.0
    wo = float(2 * fs * np.tan(np.pi * wn / fs))
    # low-pass -> high-pass: invert radially, the `order` zeros at infinity land on the origin
    p_hp = wo / p
    z_hp = np.zeros(order)
    k_hp = 1.0 * np.real(np.prod(-np.asarray([], dtype=np.float64)) / np.prod(-p))
    # bilinear transform
    fs2 = 2.0 * fs
    z_z = (fs2 + z_hp) / (fs2 - z_hp)
    p_z = (fs2 + p_hp) / (fs2 - p_hp)
    k_z = k_hp * np.real(np.prod(fs2 - z_hp) / np.prod(fs2 - p_hp))
    # zeros / poles / gain -> transfer function
    b = k_z * np.poly(z_z)
    a = np.poly(p_z)
    return np.real(b), np.real(a)


def _lfilter_zi(b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """Initial filter state of the step response's steady state (``scipy.signal.lfilter_zi`` for a[0] = 1)."""
    y_inf = np.sum(b) / np.sum(a)
    return np.flip(np.cumsum(np.flip(b - y_inf * a)))[1:]


def _compute_lower_frequency_cutoff(window_length, repetition_time) -> float:
    """1 / window length in seconds (filtering.py:98-116)."""
    window_length_in_seconds = window_length * repetition_time
    return 1 / window_length_in_seconds


def _host_doubles(v: np.ndarray):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return (C.c_double * v.size)(*v.tolist())


def _filtfilt_columns(b: np.ndarray, a: np.ndarray, data):
    """``scipy.signal.filtfilt(b, a, column)`` for every column, one kernel launch."""
    is_tensor = isinstance(data, torch.Tensor)
    x = data if is_tensor else torch.as_tensor(np.ascontiguousarray(data, dtype=np.float64))
    one_d = x.dim() == 1
    if one_d:
        x = x[:, None]
    if x.dim() != 2:
        raise ValueError("expected a (N,) series or a (N, C) array")
    x = x.to(device=x.device if x.is_cuda else "cuda", dtype=torch.float64).contiguous()
    n, c = x.shape
    order = len(a) - 1
    padlen = 3 * max(len(a), len(b))
    if n <= padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {padlen}.")
    out = torch.empty_like(x)
    if c > 0:
        nbytes = int(_lib.load_library().fcest_filtfilt_scratch_bytes(n, c, order))
        scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=x.device)
        ops.call("fcest_filtfilt", ops.ptr(x), n, c, _host_doubles(b), _host_doubles(a),
                 _host_doubles(_lfilter_zi(b, a)), order, ops.ptr(scratch), ops.ptr(out))
    if one_d:
        out = out[:, 0]
    return out if is_tensor else out.cpu().numpy()
