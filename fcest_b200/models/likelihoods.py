"""``fcest.models.likelihoods`` mirror: the Wishart-process likelihood on B200.

Same constructor, attributes and method names as the reference class
(``fcest/models/likelihoods.py:29-264``); the arithmetic of ``variational_expectations`` /
``_log_prob`` and of its gradient runs in the fused CUDA kernel ``fcest_wishart_loglik``.
Differences that a caller can observe:

* tensors are ``torch.float64`` CUDA tensors where the reference returns ``tf.Tensor``;
* the N(0,1) draw the reference takes from ``tf.random.normal`` (``:130-134``) comes from a
  counter-based Philox generator, or is *injected* with ``epsilon=`` (shape (S, N, D*nu)) -- the
  hook the parity tests use ("same seeds" = same epsilon tensor);
* a failed Cholesky raises ``RuntimeError`` (TensorFlow: ``InvalidArgumentError``, ``:184-189``).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import ops
from ..gp.parameter import Parameter

F64 = torch.float64

logging.basicConfig(
    format='%(asctime)s : %(levelname)s : %(message)s',
    datefmt='%d-%b-%y %H:%M:%S',
    level=logging.INFO
)


class MonteCarloLikelihood:
    """Placeholder for ``gpflow.likelihoods.MonteCarloLikelihood`` (only the dims are used)."""

    def __init__(self, input_dim: int, latent_dim: int, observation_dim: int) -> None:
        self.input_dim = input_dim
        self.latent_dim = latent_dim
        self.observation_dim = observation_dim


class WishartProcessLikelihood(MonteCarloLikelihood):
    """
    Class for Wishart process likelihoods.
    It specifies an observation model connecting the latent functions ('F') to the data ('Y').
    """

    def __init__(
            self,
            D: int,
            nu: int = None,
            num_mc_samples: int = 2,
            A_scale_matrix_option: str = 'train_full_matrix',
            train_additive_noise: bool = False,
            additive_noise_matrix_init: float = 0.01,
            verbose: bool = True,
            device="cuda",
    ) -> None:
        nu = D if nu is None else nu
        if nu < D:
            raise Exception("Wishart Degrees of Freedom must be >= D.")
        super().__init__(input_dim=1, latent_dim=D * nu, observation_dim=D)
        self.D = D
        self.nu = nu
        self.num_mc_samples = num_mc_samples
        self.device = torch.device(device)
        self.A_scale_matrix = self._set_A_scale_matrix(option=A_scale_matrix_option)
        # inverse softplus: the additive noise stays positive (likelihoods.py:75-86)
        additive_noise_matrix_init = np.log(np.exp(additive_noise_matrix_init) - 1)
        self.additive_part = Parameter(additive_noise_matrix_init * np.ones(D), transform=None,
                                       trainable=train_additive_noise, device=device,
                                       name="likelihood.additive_part")
        self._seed = 0
        self._draws = 0
        if verbose:
            logging.info(f"A scale matrix option is '{A_scale_matrix_option:s}'.")
            print('A_scale_matrix: ', self.A_scale_matrix.numpy())
            print('initial additive part: ', self.additive_part.numpy())

    # -- random draws ----------------------------------------------------------------------------
    def seed(self, seed: int) -> None:
        self._seed, self._draws = int(seed), 0

    def _draw(self, shape) -> torch.Tensor:
        out = torch.empty(shape, dtype=F64, device=self.device)
        ops.randn(out, self._seed, self._draws)
        self._draws += 1
        return out

    def _check_shapes(self) -> None:
        # tf.multiply(A, f_sample) broadcasts (D, D) or (D,) against (..., D, nu): needs nu == D
        if self.nu != self.D:
            raise ValueError(
                f"Incompatible shapes: A_scale_matrix {self.A_scale_matrix.shape} cannot be broadcast "
                f"against f_sample (..., {self.D}, {self.nu}); the shipped Hadamard form needs nu == D "
                "(reference likelihoods.py:175-177).")

    # -- reference API ---------------------------------------------------------------------------
    def variational_expectations(self, x_data, f_mean, f_variance, y_data, epsilon=None,
                                 need_grad: bool = False):
        """Expected log density (N,) -- likelihoods.py:93-142.  With ``need_grad`` returns the full
        dict from the fused kernel (ve, g_fmean, g_fvar, gA, glam, info)."""
        self._check_shapes()
        N = y_data.shape[0]
        R = self.D * self.nu
        if epsilon is None:
            epsilon = self._draw((self.num_mc_samples, N, R))
        epsilon = epsilon.reshape(epsilon.shape[0], N, R).contiguous()
        out = ops.wishart_loglik(f_mean, f_variance, epsilon, y_data, self.A_scale_matrix.unconstrained,
                                 self.additive_part.unconstrained, need_grad=need_grad)
        out["epsilon"] = epsilon
        self._last_info = out["info"]   # Cholesky status of this call (0 = ok), for callers that only take ve
        return out if need_grad else out["ve"]

    def check_info(self, info: torch.Tensor) -> None:
        bad = int(info.max().item())
        if bad != 0:
            raise RuntimeError(
                "Cholesky decomposition was not successful. The input might not be valid "
                f"(first non-positive pivot {bad}).")

    def _log_prob(self, x_data, f_sample, y_data):
        """MC estimate of the log likelihood given samples of F: (S, N, D, nu) -> (N,)
        (likelihoods.py:144-206)."""
        self._check_shapes()
        assert isinstance(f_sample, torch.Tensor)
        S, N = f_sample.shape[0], f_sample.shape[1]
        R = self.D * self.nu
        y = torch.as_tensor(y_data, dtype=F64, device=self.device).contiguous()
        zeros = ops.pitched(N, R, self.device, zero=True)
        ones = ops.pitched(N, R, self.device)
        ones.fill_(1.0)
        eps = f_sample.to(self.device, F64).reshape(S, N, R).contiguous()  # F = 0 + sqrt(1) * eps
        out = ops.wishart_loglik(zeros, ones, eps, y, self.A_scale_matrix.unconstrained,
                                 self.additive_part.unconstrained, need_grad=False)
        self.check_info(out["info"])
        return out["ve"]

    def _set_A_scale_matrix(self, option: str = 'identity') -> Parameter:
        """likelihoods.py:208-251."""
        if option == 'identity':
            return Parameter(np.ones(self.D), trainable=False, device=self.device, name="likelihood.A")
        elif option == 'scaled':
            return Parameter(np.ones(self.D) / self.nu, trainable=False, device=self.device,
                             name="likelihood.A")
        elif option == 'train_diagonal':
            return Parameter(np.ones(self.D), trainable=True, device=self.device, name="likelihood.A")
        elif option == 'train_full_matrix':
            return Parameter(np.eye(self.D), trainable=True, device=self.device, name="likelihood.A")
        else:
            raise NotImplementedError(f"Option '{option:s}' for 'A_scale_matrix' not recognized.")

    def _add_diagonal_additive_noise(self, cov_matrix):
        """likelihoods.py:253-264."""
        lam = torch.nn.functional.softplus(self.additive_part.unconstrained)
        return cov_matrix + torch.diag_embed(lam.expand(cov_matrix.shape[:-1]))


class FactoredWishartProcessLikelihood(MonteCarloLikelihood):

    def __init__(self):
        raise NotImplementedError("Factorized Wishart process not implemented yet.")
