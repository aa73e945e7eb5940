"""``fcest.models.wishart_process`` mirror: VWP / SVWP on B200.

Public surface follows the reference classes (``fcest/models/wishart_process.py:35-308`` and
``:311-602``): same constructor arguments and defaults, ``predict_cov`` / ``predict_corr`` /
``predict_cov_samples`` / ``predict_f``, ``save_model_params_dict`` / ``load_from_params_dict`` with the
same JSON schema, plus the slice of the GPflow model API that ``run_adam`` uses
(``training_loss_closure``, ``training_loss``, ``trainable_variables``, ``elbo``).

What the reference inherits from ``gpflow.models.VGP`` / ``SVGP`` (ELBO = sum of variational
expectations - KL, whitened conditional) is sequenced here on the device by
``fcest_b200.gp.conditionals`` and the fused likelihood kernel; gradients are analytic (no autograd
graph).  ``elbo_and_grad`` is the training primitive: one ELBO value plus d ELBO / d every trainable
(unconstrained) variable, stored on ``Parameter.grad``.
"""
from __future__ import annotations

import json
import logging
import os

import numpy as np
import torch

from .. import _lib, ops
from ..gp import conditionals as cond
from ..gp.kernels import Kernel, Matern52
from ..gp.parameter import Parameter
from ..helpers.array_operations import are_all_positive_definite
from .likelihoods import FactoredWishartProcessLikelihood, WishartProcessLikelihood

F64 = torch.float64

logging.basicConfig(
    format='%(asctime)s : %(levelname)s : %(message)s',
    datefmt='%d-%b-%y %H:%M:%S',
    level=logging.INFO
)


def _dev_tensor(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=F64).contiguous()
    return torch.tensor(np.ascontiguousarray(a, dtype=np.float64), dtype=F64, device=device)


def _init_variational(m: int, num_latent: int, device):
    """gpflow's initial q(u): q_mu zeros (M, L), q_sqrt identity (L, M, M) -- built on the device (a D = 200 model holds
    0.8 - 13 GB of q_sqrt; going through a NumPy array made constructing many subjects host-bound)."""
    q_mu = Parameter(np.zeros((1, 1)), device=device, name="q_mu")
    q_mu.unconstrained = torch.zeros((m, num_latent), dtype=F64, device=device)
    q_sqrt = Parameter(np.zeros((1, 1, 1)), transform="tril", device=device, name="q_sqrt")
    q_sqrt.unconstrained = torch.eye(m, dtype=F64, device=device).expand(num_latent, m, m).contiguous()
    return q_mu, q_sqrt


class InducingPoints:
    """``gpflow.inducing_variables.InducingPoints`` stand-in (only ``.Z`` is used)."""

    def __init__(self, Z, device):
        self.Z = Parameter(Z, device=device, name="inducing_variable.Z")


class _WishartProcessBase:
    """Shared ELBO / gradient / prediction machinery of VWP and SVWP."""

    _sparse: bool

    # ---- GPflow-model surface used by run_adam ---------------------------------------------------
    @property
    def parameters(self) -> list:
        ps = [self.kernel.variance, self.kernel.lengthscales]
        if self._sparse:
            ps.append(self.inducing_variable.Z)
        ps += [self.q_mu, self.q_sqrt, self.likelihood.A_scale_matrix, self.likelihood.additive_part]
        return ps

    @property
    def trainable_parameters(self) -> list:
        return [p for p in self.parameters if p.trainable]

    @property
    def trainable_variables(self) -> list:
        return [p.unconstrained for p in self.trainable_parameters]

    def _data(self, data):
        if data is None:
            data = self.data
        x, y = data
        return _dev_tensor(x, self.device), _dev_tensor(y, self.device)

    def predict_f(self, x_new, full_cov: bool = False):
        """(mean (N*, D nu), var (N*, D nu)) of the latent GPs -- gpflow predict_f (whitened)."""
        assert not full_cov, "full_cov is not used by FCEst"
        x_new = _dev_tensor(x_new, self.device)
        Z = self.inducing_variable.Z.unconstrained if self._sparse else self._x_train
        fmean, fvar, _, st = cond.conditional_forward(self.kernel, Z, x_new, self.q_mu.unconstrained,
                                                      self.q_sqrt.unconstrained, vgp=False, want_kl=False)
        self._check_chol(st.info)
        return fmean, fvar

    @staticmethod
    def _check_chol(info) -> None:
        bad = int(info.max().item())
        if bad != 0:
            raise RuntimeError(f"Cholesky decomposition of the kernel matrix was not successful (pivot {bad}).")

    def _moments(self, x, want_kl=True):
        if self._sparse:
            Z = self.inducing_variable.Z.unconstrained
            return cond.conditional_forward(self.kernel, Z, x, self.q_mu.unconstrained,
                                            self.q_sqrt.unconstrained, vgp=False, want_kl=want_kl)
        return cond.conditional_forward(self.kernel, x, x, self.q_mu.unconstrained,
                                        self.q_sqrt.unconstrained, vgp=True, want_kl=want_kl)

    def elbo(self, data=None, epsilon=None, check: bool = True) -> torch.Tensor:
        """Evidence lower bound (device scalar): sum_n var_exp[n] - KL  (gpflow SVGP.elbo / VGP.elbo).
        A failed Cholesky (kernel matrix or a Sigma) raises, as ``tf.linalg.cholesky`` does in the reference."""
        x, y = self._data(data)
        fmean, fvar, kl, st = self._moments(x)
        ve = self.likelihood.variational_expectations(x, fmean, fvar, y, epsilon=epsilon)
        if check:
            self._check_chol(st.info)
            self.likelihood.check_info(self.likelihood._last_info)
        out = torch.empty(1, dtype=F64, device=self.device)
        ops.sum_into(ve, ve.numel(), 1, 1.0, out, False)
        ops.axpby(-1.0, kl, 1.0, out)
        return out[0]

    def maximum_log_likelihood_objective(self, data=None):
        return self.elbo(data)

    def training_loss(self, data=None):
        return -self.elbo(data)

    def elbo_and_grad(self, data=None, epsilon=None, check: bool = True, kl_scale: float = 1.0,
                      gk_reduce=None, q_sqrt_adam: dict | None = None) -> torch.Tensor:
        """One ELBO evaluation + gradient w.r.t. every trainable variable (the `eval` of the metric).

        Gradients (d ELBO / d unconstrained) land in ``Parameter.grad``; returns the ELBO (device scalar).
        ``kl_scale`` / ``gk_reduce`` are the hooks of the time-sharded multi-GPU evaluation
        (``fcest_b200.parallel.TimeShardedSVWP``): the KL term is weighted 1 / world on every rank and the
        packed q_sqrt cotangent is all-reduced inside the backward pass.  ``q_sqrt_adam`` (see
        ``conditional_backward``) fuses the optimiser's q_sqrt update into the backward pass; ``q_sqrt.grad`` is then
        ``None`` and q_sqrt has already moved when this returns.
        """
        x, y = self._data(data)
        lik = self.likelihood
        if epsilon is None:
            # the Monte-Carlo draw does not depend on the conditional: generate it on its own stream underneath it
            main = torch.cuda.current_stream(self.device)
            rng_stream = cond._side_stream(self.device, 1)
            rng_stream.wait_stream(main)
            with torch.cuda.stream(rng_stream):
                epsilon = lik._draw((lik.num_mc_samples, y.shape[0], self.D * self.nu))
            fmean, fvar, kl, st = self._moments(x)
            main.wait_stream(rng_stream)
            epsilon.record_stream(main)
        else:
            fmean, fvar, kl, st = self._moments(x)
        out = lik.variational_expectations(x, fmean, fvar, y, epsilon=epsilon, need_grad=True)
        elbo = torch.empty(1, dtype=F64, device=self.device)
        ops.sum_into(out["ve"], out["ve"].numel(), 1, 1.0, elbo, False)
        ops.axpby(-float(kl_scale), kl, 1.0, elbo)
        need_Z = self._sparse and self.inducing_variable.Z.trainable
        g = cond.conditional_backward(st, out["g_fmean"], out["g_fvar"], need_Z=need_Z, kl_scale=float(kl_scale),
                                      gk_reduce=gk_reduce, q_sqrt_adam=q_sqrt_adam)
        self.q_mu.grad = g["q_mu"]
        self.q_sqrt.grad = g["q_sqrt"]
        kv, kl_ = self.kernel.variance, self.kernel.lengthscales
        kv.grad = ops.softplus_bwd(kv.unconstrained.reshape(1), g["kvar"], torch.empty_like(g["kvar"])).reshape(kv.shape)
        kl_.grad = ops.softplus_bwd(kl_.unconstrained.reshape(1), g["kls"], torch.empty_like(g["kls"])).reshape(kl_.shape)
        if self._sparse:
            self.inducing_variable.Z.grad = g["Z"]
        lik.A_scale_matrix.grad = out["gA"]
        lik.additive_part.grad = out["glam"]
        self._last_info = (out["info"], st.info)
        if check:
            lik.check_info(out["info"])
            self._check_chol(st.info)
        return elbo[0]

    def compile_evaluation(self, data=None):
        """ELBO + gradient evaluation captured in a CUDA graph (the analogue of ``tf.function`` on the evaluation
        path): returns a callable ``f(epsilon=None) -> elbo`` that refreshes ``Parameter.grad`` with ONE graph launch
        (plus the Philox draw).  The ~55 kernel launches of an evaluation dominate the small configurations
        (C1: N=200, D=3; C2: N=1200, D=15); parameters are read from their live tensors, so optimiser updates
        between calls are seen.  ``f.check()`` inspects the Cholesky status of the last call (a host sync)."""
        from ..helpers.inference import GraphedEvaluation
        return GraphedEvaluation(self, data)

    def training_loss_closure(self, data=None, compile: bool = True):
        """Callable returning the loss -ELBO (and refreshing ``Parameter.grad``), like
        ``gpflow.models.*.training_loss_closure`` used at ``fcest/helpers/inference.py:61-70``."""
        def closure():
            return -self.elbo_and_grad(data)
        return closure

    # ---- prediction (reference wishart_process.py:120-212 / :392-504) -----------------------------
    def _moment_pass(self, x_new, num_mc_samples, corr, want_samples=False, epsilon=None):
        lik = self.likelihood
        lik._check_shapes()
        f_mean_new, f_variance_new = self.predict_f(x_new)
        n, R = f_mean_new.shape
        D = self.D
        dev = self.device
        A = lik.A_scale_matrix.unconstrained
        a_mode = 0 if A.dim() == 2 else 1
        mean = torch.empty((n, D, D), dtype=F64, device=dev)
        m2 = torch.empty((n, D, D), dtype=F64, device=dev)
        out_mean = torch.empty((n, D, D), dtype=F64, device=dev)
        out_std = torch.empty((n, D, D), dtype=F64, device=dev)
        samples = torch.empty((num_mc_samples, n, D, D), dtype=F64, device=dev) if want_samples else None
        nscr = int(_lib.load_library().fcest_cov_moments_scratch_bytes(D, self.nu))
        scratch = torch.empty(nscr // 8, dtype=F64, device=dev) if nscr else None  # large-D tile scratch
        # stream the samples in chunks that keep the epsilon buffer around 1 GB
        chunk = num_mc_samples if epsilon is not None else max(1, min(num_mc_samples, (1 << 27) // max(n * R, 1)))
        done = 0
        while done < num_mc_samples:
            sc = min(chunk, num_mc_samples - done)
            if epsilon is not None:
                eps = epsilon.reshape(num_mc_samples, n, R)[done:done + sc].contiguous()
            else:
                eps = lik._draw((sc, n, R))
            smp = samples[done:done + sc] if want_samples else None
            last = done + sc == num_mc_samples
            ops.call("fcest_cov_moments", ops.ptr(f_mean_new), ops.ptr(f_variance_new), ops.ld(f_mean_new),
                     ops.ptr(eps), ops.ptr(A), a_mode, ops.ptr(lik.additive_part.unconstrained), sc, n, D,
                     self.nu, int(corr), done, ops.ptr(mean), ops.ptr(m2), ops.ptr(smp), int(last),
                     ops.ptr(out_mean), ops.ptr(out_std), ops.ptr(scratch))
            done += sc
        return out_mean, out_std, samples

    def predict_cov(self, x_new, num_mc_samples: int = 300, epsilon=None):
        """Mean and (population) std over MC samples of the covariance at ``x_new`` -> 2 x (N*, D, D)."""
        cov_mean, cov_stddev, _ = self._moment_pass(x_new, num_mc_samples, corr=False, epsilon=epsilon)
        assert are_all_positive_definite(cov_mean)
        return cov_mean, cov_stddev

    def predict_corr(self, x_new, num_mc_samples: int = 300, epsilon=None):
        """Mean and std over MC samples of the correlation matrices at ``x_new``."""
        corr_mean, corr_stddev, _ = self._moment_pass(x_new, num_mc_samples, corr=True, epsilon=epsilon)
        return corr_mean, corr_stddev

    def _get_cov_samples(self, x_new, num_mc_samples: int = 300, epsilon=None) -> torch.Tensor:
        """(S*, N*, D, D) covariance samples (A o F)(A o F)^T + diag(softplus(lam))."""
        _, _, samples = self._moment_pass(x_new, num_mc_samples, corr=False, want_samples=True,
                                          epsilon=epsilon)
        return samples

    # ---- parameter (de)serialisation: same JSON schema as the reference --------------------------
    def _params_dict(self) -> dict:
        d = {
            'D': self.D,
            'nu': self.nu,
            'A_scale_matrix': self.likelihood.A_scale_matrix.numpy().tolist(),
            'additive_noise': self.likelihood.additive_part.numpy().tolist(),
            'kernel_variance': float(self.kernel.variance.numpy()),
            'kernel_lengthscales': float(self.kernel.lengthscales.numpy()),
        }
        if self._sparse:
            d['Z'] = self.inducing_variable.Z.numpy().tolist()
        d['q_mu'] = self.q_mu.numpy().tolist()
        d['q_sqrt'] = self.q_sqrt.numpy().tolist()
        return d

    def save_model_params_dict(self, savedir: str, model_name: str) -> None:
        params_dict = self._params_dict()
        if not os.path.exists(savedir):
            os.makedirs(savedir)
        with open(os.path.join(savedir, model_name), 'w') as fp:
            json.dump(params_dict, fp)
        logging.info(f"Model '{model_name:s}' saved in '{savedir:s}'.")

    def load_from_params_dict(self, savedir: str, model_name: str) -> None:
        with open(os.path.join(savedir, model_name), 'r') as fp:
            params_dict = json.load(fp)
        A_scale_matrix = np.array(params_dict['A_scale_matrix'])
        if self._sparse and A_scale_matrix.ndim == 1:
            # reference wishart_process.py:579-581
            A_scale_matrix = A_scale_matrix * np.eye(len(A_scale_matrix))
            print('Squared A scale matrix:', A_scale_matrix)
        self.likelihood.A_scale_matrix.assign(A_scale_matrix)
        self.likelihood.additive_part.assign(np.array(params_dict['additive_noise']))
        self.kernel.variance.assign(params_dict['kernel_variance'])
        self.kernel.lengthscales.assign(params_dict['kernel_lengthscales'])
        if self._sparse:
            self.inducing_variable.Z.assign(np.array(params_dict['Z']))
        self.q_mu.assign(np.array(params_dict['q_mu']))
        self.q_sqrt.assign(np.array(params_dict['q_sqrt']))
        logging.info(f"{'SVWP' if self._sparse else 'VWP'} model loaded from '{savedir:s}'.")


class VariationalWishartProcess(_WishartProcessBase):
    """Variational Wishart process (VWP) -- reference ``wishart_process.py:35-308`` (gpflow VGP)."""

    _sparse = False

    def __init__(
            self,
            x_observed: np.array,
            y_observed: np.array,
            nu: int = None,
            kernel: Kernel = None,
            num_mc_samples: int = 5,
            A_scale_matrix_option: str = 'train_full_matrix',
            train_additive_noise: bool = True,
            kernel_lengthscale_init: float = 0.3,
            q_sqrt_init: float = 0.001,
            num_factors: int = None,
            device="cuda",
    ) -> None:
        self.device = torch.device(device)
        self.D = y_observed.shape[1]
        logging.info(f"Found {self.D:d} time series (D = {self.D:d}).")
        if num_factors is not None:
            raise NotImplementedError("Factorized Wishart process not implemented yet.")
        if nu is None:
            nu = self.D
        self.nu = nu
        if kernel is None:
            kernel = Matern52(device=device)
        self.kernel = kernel
        self.likelihood = WishartProcessLikelihood(
            D=self.D, nu=nu, num_mc_samples=num_mc_samples,
            A_scale_matrix_option=A_scale_matrix_option,
            train_additive_noise=train_additive_noise, device=device,
        )
        self.data = (x_observed, y_observed)
        self._x_train = _dev_tensor(x_observed, self.device)
        self.num_latent_gps = self.likelihood.latent_dim
        n = x_observed.shape[0]
        # gpflow VGP: q_mu zeros (N, L), q_sqrt identity (L, N, N)
        self.q_mu, self.q_sqrt = _init_variational(n, self.num_latent_gps, self.device)
        self._initialize_parameters(kernel_lengthscale_init, q_sqrt_init)

    def _initialize_parameters(self, kernel_lengthscale_init: float, q_sqrt_init: float) -> None:
        """wishart_process.py:214-237."""
        try:
            self.kernel.lengthscales.assign(np.ones_like(self.kernel.lengthscales.numpy()) * kernel_lengthscale_init)
        except Exception:
            print('Non-standard kernel detected.')
        self.q_sqrt.unconstrained.mul_(q_sqrt_init)


class SparseVariationalWishartProcess(_WishartProcessBase):
    """Sparse variational Wishart process (SVWP) -- reference ``wishart_process.py:311-602`` (gpflow SVGP)."""

    _sparse = True

    def __init__(
            self,
            D: int,
            Z: np.array,
            nu: int = None,
            kernel: Kernel = None,
            num_mc_samples: int = 5,
            A_scale_matrix_option: str = 'train_full_matrix',
            train_additive_noise: bool = True,
            kernel_lengthscale_init: float = 0.3,
            q_sqrt_init: float = 0.001,
            num_factors: int = None,
            verbose: bool = True,
            device="cuda",
    ) -> None:
        self.device = torch.device(device)
        self.D = D
        logging.info(f"Found {self.D:d} time series (D = {self.D:d}).")
        assert len(Z.shape) == 2
        if nu is None:
            nu = self.D
        self.nu = nu
        if num_factors is not None:
            likel = FactoredWishartProcessLikelihood()
        else:
            likel = WishartProcessLikelihood(
                D=self.D, nu=nu, num_mc_samples=num_mc_samples,
                A_scale_matrix_option=A_scale_matrix_option,
                train_additive_noise=train_additive_noise, verbose=verbose, device=device,
            )
            assert likel.latent_dim == self.D * nu
        self.likelihood = likel
        # the reference's default argument is one shared Matern52() instance (wishart_process.py:326);
        # a fresh kernel per model is used here
        self.kernel = Matern52(device=device) if kernel is None else kernel
        self.inducing_variable = InducingPoints(Z, device)
        self.num_latent_gps = likel.latent_dim
        self.data = None
        m = Z.shape[0]
        self.q_mu, self.q_sqrt = _init_variational(m, self.num_latent_gps, self.device)
        self._initialize_parameters(kernel_lengthscale_init, q_sqrt_init)

    def predict_cov_samples(self, x_new, num_mc_samples: int = 300):
        """wishart_process.py:417-434."""
        return self._get_cov_samples(x_new, num_mc_samples)

    def scale_matrix(self):
        A = self.likelihood.A_scale_matrix.unconstrained
        return A * A.T

    def _initialize_parameters(self, kernel_lengthscale_init: float, q_sqrt_init: float) -> None:
        """wishart_process.py:509-524."""
        self.kernel.lengthscales.assign(kernel_lengthscale_init)
        self.q_sqrt.unconstrained.mul_(q_sqrt_init)
