"""``fcest.helpers.summary_measures`` mirror: temporal summaries of a TVFC estimate on the GPU.

Reference: ``fcest/helpers/summary_measures.py:16-126``.  Same function names and arguments; NumPy (N, D, D) in,
NumPy (D, D) out (a CUDA tensor in gives a CUDA tensor out, so estimates that are already in HBM -- the output of
``fcest_window_cov`` or ``predict_cov`` -- are summarised without a trip through the host).  Every metric is one
launch of ``fcest_tvfc_summary`` (``csrc/summary.cu``); the reference's AR(1) route fits a statsmodels
``AutoReg(lags=1)`` per edge in a Python loop, which is the ordinary least-squares slope computed here.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import ops

__all__ = [
    "summarize_tvfc_estimates",
    "fit_and_extract_ar1_param",
    "compute_rate_of_change",
    "_rate_of_change",
]

_METRIC = {"mean": 0, "variance": 1, "std": 2, "rate_of_change": 3, "ar1": 4}


def _summary(full_covariance_structure, metric: str):
    is_tensor = isinstance(full_covariance_structure, torch.Tensor)
    c = full_covariance_structure if is_tensor else torch.as_tensor(
        np.ascontiguousarray(full_covariance_structure, dtype=np.float64))
    if c.dim() != 3 or c.shape[1] != c.shape[2]:
        raise ValueError("expected a TVFC array of shape (N, D, D)")
    c = c.to(device=c.device if c.is_cuda else "cuda", dtype=torch.float64).contiguous()
    n, d = c.shape[0], c.shape[1]
    out = torch.empty((d, d), dtype=torch.float64, device=c.device)
    ops.call("fcest_tvfc_summary", ops.ptr(c), n, d, _METRIC[metric], ops.ptr(out))
    return out if is_tensor else out.cpu().numpy()


def summarize_tvfc_estimates(full_covariance_structure, tvfc_summary_metric: str):
    """Summarize a full TVFC structure (N, D, D) over the temporal axis -> (D, D) (summary_measures.py:16-47)."""
    match tvfc_summary_metric:
        case 'ar1':
            return fit_and_extract_ar1_param(full_covariance_structure)
        case 'fourier_max':
            raise NotImplementedError
        case 'fourier_mean':
            raise NotImplementedError
        case 'mean' | 'std' | 'variance':
            return _summary(full_covariance_structure, tvfc_summary_metric)
        case 'rate_of_change':
            return compute_rate_of_change(full_covariance_structure)
        case _:
            logging.error(f"TVFC summary metric {tvfc_summary_metric:s} not recognized.")


def fit_and_extract_ar1_param(full_covariance_structure):
    """AR(1) coefficient of each edge's time series, diagonal zero (summary_measures.py:50-80)."""
    return _summary(full_covariance_structure, "ar1")


def compute_rate_of_change(full_covariance_structure):
    """Average relative step size across time (summary_measures.py:83-105)."""
    return _summary(full_covariance_structure, "rate_of_change")


def _rate_of_change(current_value: np.array, previous_value: np.array) -> np.array:
    """One step of the rate of change (summary_measures.py:108-126; plain NumPy, the kernel fuses it)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        rate_of_change = np.abs(current_value / previous_value - 1)
    rate_of_change[rate_of_change == np.inf] = 1
    rate_of_change[rate_of_change > 10] = 1
    return rate_of_change
