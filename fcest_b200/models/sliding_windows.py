"""``fcest.models.sliding_windows`` mirror: sliding-window TVFC estimation on B200.

Same class, constructor and method names / defaults as the reference
(``fcest/models/sliding_windows.py:27-427``).  ``overlapping_windowed_cov_estimation`` and the
cross-validated window search run in ``fcest_window_cov`` / ``fcest_window_cv``; inputs and outputs stay
NumPy / pandas as in the reference.  The optional Butterworth high-pass filter the reference applies when
``repetition_time`` is passed to ``overlapping_windowed_cov_estimation`` runs on the device too
(``fcest_filtfilt``, ``helpers/filtering.py``), so the data makes one trip to HBM.  ``save_tvfc_estimates`` /
``load_tvfc_estimates`` keep the reference's CSV layout (one row per matrix entry, one column per time step).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import pandas as pd
import torch

from .. import ops
from ..helpers.data import to_3d_format
from ..helpers.filtering import highpass_filter_data

# scipy.stats._multivariate._eigvalsh_to_eps: 1e6 * eps relative eigenvalue threshold for "singular"
_SINGULAR_REL_TOL = 1e6 * float(np.finfo(np.float64).eps)


class SlidingWindows:
    """Main class for sliding-windows methods for time-varying functional connectivity (TVFC) estimation."""

    def __init__(
            self,
            x_train_locations: np.array,
            y_train_locations: np.array,
            repetition_time: float = None,
            window_shape: str = 'rectangle',
            device="cuda",
    ) -> None:
        self.x_train = x_train_locations
        self.y_train = y_train_locations
        assert len(x_train_locations) == len(y_train_locations)
        self.num_time_steps = y_train_locations.shape[0]  # N
        self.num_time_series = y_train_locations.shape[1]  # D
        self.device = device
        if repetition_time is None:
            logging.info("No repetition time found, assuming we are running simulations.")
            repetition_time = 1.0
        self.repetition_time = repetition_time
        self._set_min_max_proposed_window_length()
        if window_shape == 'tapered':
            logging.error("Tapered sliding windows not implemented yet.")
        elif window_shape != 'rectangle':
            logging.error("Other window types not implemented yet.")

    def _y_device(self) -> torch.Tensor:
        return torch.tensor(np.ascontiguousarray(self.y_train, dtype=np.float64), device=self.device)

    def estimate_static_functional_connectivity(self, connectivity_metric: str, return_structure: bool = True):
        """sliding_windows.py:74-107 (plain NumPy, not on the hot path)."""
        match connectivity_metric:
            case 'covariance':
                sfc_estimate = np.cov(self.y_train.T)
            case 'correlation':
                sfc_estimate = np.corrcoef(self.y_train.T)
            case _:
                raise NotImplementedError(f"Connectivity metric {connectivity_metric:s} not recognized.")
        if return_structure:
            sfc_estimate = np.tile(sfc_estimate, reps=(self.num_time_steps, 1, 1))
        return sfc_estimate

    def overlapping_windowed_cov_estimation(
            self,
            window_length: int,
            step_size: int = 1,
            repetition_time: float = None,
            connectivity_metric: str = 'covariance',
    ) -> np.array:
        """Overlapping sliding-windows estimate, (N, D, D) -- sliding_windows.py:109-174."""
        n, d = self.num_time_steps, self.num_time_series
        y = self._y_device()
        # Highpass filtering as recommended by Leonardi2015 (sliding_windows.py:147-154).
        if repetition_time is not None:
            y = highpass_filter_data(y, window_length=window_length, repetition_time=repetition_time)
        out = torch.empty((n, d, d), dtype=torch.float64, device=y.device)
        ops.call("fcest_window_cov", ops.ptr(y), n, d, int(window_length), int(step_size),
                 1 if connectivity_metric == 'correlation' else 0, ops.ptr(out))
        return self._to_numpy(out)

    _pinned = None

    @classmethod
    def _to_numpy(cls, t: torch.Tensor) -> np.ndarray:
        """Device -> NumPy through a cached pinned staging buffer (a pageable D2H copy of the 8 N D^2
        byte result would cost more than the estimator itself)."""
        if cls._pinned is None or cls._pinned.numel() < t.numel():
            cls._pinned = torch.empty(t.numel(), dtype=torch.float64, pin_memory=True)
        stage = cls._pinned[:t.numel()].view(t.shape)
        stage.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        # pinned staging -> caller-owned pageable memory with torch's multi-threaded copy (numpy's single-threaded
        # .copy() of the 96 MB result cost more than the kernel and the PCIe transfer together)
        return torch.empty(t.shape, dtype=torch.float64).copy_(stage).numpy()

    def compute_cross_validated_optimal_window_length(
            self,
            window_length_step_size: int = 2,
            eval_location_step_size: int = 1,
    ) -> int:
        results_df = self.find_cross_validated_optimal_window_length(
            window_length_step_size=window_length_step_size,
            eval_location_step_size=eval_location_step_size
        )
        return self.get_optimal_window_length(results_df)

    def find_cross_validated_optimal_window_length(
            self,
            window_length_step_size: int = 2,
            eval_location_step_size: int = 1,
    ) -> pd.DataFrame:
        """DataFrame (eval locations x proposed window lengths) of held-out log densities
        (sliding_windows.py:199-274)."""
        proposal_window_length_range = np.arange(
            self.minimum_proposal_window_length, self.maximum_proposal_window_length, window_length_step_size)
        num_w = len(proposal_window_length_range)
        logging.info(f"Proposing {num_w:d} window lengths (from {self.minimum_proposal_window_length:d} to "
                     f"{self.maximum_proposal_window_length:d}, in steps of {window_length_step_size:d}).")
        t_lo = int((self.maximum_proposal_window_length - 1) / 2)
        t_hi = self.num_time_steps - int((self.maximum_proposal_window_length + 1) / 2)
        eval_locations = np.arange(t_lo, t_hi, eval_location_step_size)
        num_t = len(eval_locations)
        logging.info(f"Evaluating on {num_t:d} observations (from position {t_lo:d} to {t_hi:d}).")
        y = self._y_device()
        table = torch.empty((num_t, num_w), dtype=torch.float64, device=y.device)
        if num_t > 0 and num_w > 0:
            ops.call("fcest_window_cv", ops.ptr(y), self.num_time_steps, self.num_time_series,
                     int(proposal_window_length_range[0]), int(window_length_step_size), num_w, int(t_lo),
                     int(eval_location_step_size), num_t, _SINGULAR_REL_TOL, ops.ptr(table), num_w)
        return pd.DataFrame(table.cpu().numpy(), index=eval_locations, columns=proposal_window_length_range)

    def _set_min_max_proposed_window_length(
            self,
            min_proposal_window_length_seconds: float = 20.0,
            max_proposal_window_length_seconds: float = 180.0,
    ) -> None:
        """sliding_windows.py:276-304 (repetition_time is never None here, see :63-66, so the per-N
        table of the reference, :306-329, is unreachable)."""
        lo = int(np.ceil(min_proposal_window_length_seconds / self.repetition_time))
        hi = int(np.minimum(
            int(np.ceil(self.num_time_steps * self.repetition_time / 2 / self.repetition_time)),
            int(np.ceil(max_proposal_window_length_seconds / self.repetition_time))))
        if lo % 2 == 0:
            lo += 1
        if hi % 2 == 0:
            hi += 1
        self.minimum_proposal_window_length = lo
        self.maximum_proposal_window_length = hi

    @staticmethod
    def get_optimal_window_length(results_df: pd.DataFrame) -> int:
        """sliding_windows.py:331-350."""
        mean_results_df = results_df.mean(axis=0)
        optimal_window_length = mean_results_df.index[mean_results_df.argmax()]
        logging.info(f"Obtained optimal window length: {optimal_window_length:d} TRs.")
        return optimal_window_length

    def windowed_cov_estimation(self, num_windows: int) -> np.array:
        """Segmented, non-overlapping windows (sliding_windows.py:352-372; plain NumPy)."""
        num_time_steps, num_time_series = self.y_train.shape
        cov_structure = np.zeros((num_time_steps, num_time_series, num_time_series))
        window_length = num_time_steps / num_windows
        for i_window in range(num_windows):
            idx = np.arange(int(i_window * window_length), int((i_window + 1) * window_length))
            cov_structure[idx, :, :] = np.cov(self.y_train[idx, :].T)
        return cov_structure

    def save_tvfc_estimates(
            self,
            optimal_window_length: int,
            savedir: str, model_name: str,
            repetition_time: float = None,
            connectivity_metric: str = 'correlation',
    ) -> None:
        """Estimate at the train locations and write the (D*D, N) table as CSV (sliding_windows.py:374-406)."""
        train_loc_cov_structure = self.overlapping_windowed_cov_estimation(
            window_length=optimal_window_length,
            repetition_time=repetition_time,
            connectivity_metric=connectivity_metric
        )  # (N_train, D, D)
        cov_structure_df = pd.DataFrame(
            train_loc_cov_structure.reshape(len(train_loc_cov_structure), -1).T
        )  # (D*D, N)
        if not os.path.exists(savedir):
            os.makedirs(savedir)
        cov_structure_df.to_csv(os.path.join(savedir, model_name))
        logging.info(f"Saved SW-CV model (train location) estimates to '{savedir:s}'.")

    @staticmethod
    def load_tvfc_estimates(savedir: str, model_name: str) -> np.array:
        """Read a table written by ``save_tvfc_estimates`` (sliding_windows.py:408-427).  As in the reference the
        file is read without ``index_col``, so the row index that ``to_csv`` wrote comes back as a leading column and
        the result has N + 1 leading slices (slice 0 holds the row numbers); callers of the reference drop it."""
        train_loc_cov_structure = pd.read_csv(os.path.join(savedir, model_name))  # (D*D, N + 1)
        return to_3d_format(train_loc_cov_structure)
