"""``fcest.helpers.data`` mirror: the on-disk layout of TVFC estimates (reference ``fcest/helpers/data.py:11-44``).

A (N, D, D) structure is stored as a 2-D table with one row per matrix entry and one column per time step; these
two helpers convert between the layouts (host-side NumPy: the data formats either side of the sliding-window path).
"""
from __future__ import annotations

import numpy as np

__all__ = ["to_2d_format", "to_3d_format"]


def to_2d_format(three_dimensional_cov_matrices_array: np.array) -> np.array:
    """(N, D, D) -> (D*D, N): row i*D + j holds the time series of entry (i, j) (data.py:11-24)."""
    a = np.asarray(three_dimensional_cov_matrices_array)
    return a.reshape(len(a), -1).T


def to_3d_format(r_formatted_array: np.array) -> np.array:
    """(D*D, N) -> (N, D, D) (data.py:27-44).  Entry (i, j) of the table lands at [n, j, i]: the transpose of
    what ``to_2d_format`` flattened, which is the same matrix for the symmetric structures this is used on."""
    assert len(r_formatted_array.shape) == 2
    num_time_series = int(np.sqrt(r_formatted_array.shape[0]))
    cube = np.reshape(r_formatted_array, (num_time_series, num_time_series, r_formatted_array.shape[1]))
    return np.transpose(cube, (2, 1, 0))
