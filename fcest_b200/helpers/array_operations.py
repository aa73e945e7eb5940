"""``fcest.helpers.array_operations`` mirror -- the three helpers that sit on the hot path
(reference ``fcest/helpers/array_operations.py:19-103``)."""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import ops

__all__ = [
    "to_correlation_structure",
    "_correlation_from_covariance",
    "convert_tensor_to_correlation",
    "are_all_positive_definite",
]


def to_correlation_structure(covariance_structure: np.array) -> np.array:
    """(N, D, D) covariance structure -> correlation structure (array_operations.py:19-34)."""
    return np.array([_correlation_from_covariance(c) for c in covariance_structure])


def _correlation_from_covariance(covariance_matrix: np.array) -> np.array:
    """array_operations.py:37-53 (entries that are exactly zero stay zero)."""
    v = np.sqrt(np.diag(covariance_matrix))
    outer_v = np.outer(v, v)
    correlation_matrix = covariance_matrix / outer_v
    correlation_matrix[covariance_matrix == 0] = 0
    return correlation_matrix


def convert_tensor_to_correlation(covariance_matrix: torch.Tensor) -> torch.Tensor:
    """Sigma_ij / sqrt(Sigma_ii Sigma_jj) on the trailing two axes (array_operations.py:56-72).

    ``predict_corr`` does not call this (the conversion is fused into ``fcest_cov_moments``); it is
    kept for callers that convert sample tensors themselves."""
    diag = torch.diagonal(covariance_matrix, dim1=-2, dim2=-1).unsqueeze(-1)
    return covariance_matrix / torch.sqrt(diag @ diag.transpose(-1, -2))


def are_all_positive_definite(covariance_matrices) -> bool:
    """True iff every (D, D) matrix is symmetric (|A - A^T| <= 1e-8) and Cholesky-factorisable
    (array_operations.py:75-103), checked by one batched kernel instead of a Python loop."""
    t = covariance_matrices
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t), dtype=torch.float64)
    if t.dim() == 2:
        t = t[None]
    t = t.to(device=t.device if t.is_cuda else "cuda", dtype=torch.float64).contiguous()
    B, D = t.shape[0], t.shape[-1]
    info = torch.zeros(B, dtype=torch.int32, device=t.device)
    scratch = None
    if D * (D + 1) * 8 > 227 * 1024:
        scratch = torch.empty(B * D * (D + 1), dtype=torch.float64, device=t.device)
    ops.call("fcest_batched_chol_info", ops.ptr(t), B, D, 1e-8, ops.ptr(info), ops.ptr(scratch))
    info = info.cpu().numpy()
    if (info < 0).any():
        logging.warning('Matrix is not symmetric.')
        return False
    if (info > 0).any():
        print("Matrix is not positive definite")
        return False
    return True
