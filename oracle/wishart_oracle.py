"""CPU oracle for the Wishart-process hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the *checker* or the *timed CPU baseline* -- never on the
product path.  The product (``fcest_b200``) must not import anything from ``oracle/``.

What it is: a float64 torch-CPU restatement, op for op, of the arithmetic that FCEst's
TensorFlow/GPflow path executes for

* ``WishartProcessLikelihood.variational_expectations`` / ``_log_prob``
  (reference ``fcest/models/likelihoods.py:93-206`` and ``:253-264``),
* ``_get_cov_samples`` / ``predict_cov`` / ``predict_corr``
  (``fcest/models/wishart_process.py:120-212`` and ``:392-504``,
  ``fcest/helpers/array_operations.py:56-103``),
* the GPflow pieces FCEst subclasses (``gpflow.models.SVGP.elbo``, ``gpflow.models.VGP.elbo``,
  ``gpflow.conditionals.util.base_conditional``, ``gpflow.kullback_leiblers.gauss_kl``,
  ``gpflow.kernels.Matern52``) -- **third-party, un-vendored**: ``gpflow>=2.6.0`` /
  ``tensorflow>=2.10`` (``setup.py:27,33``, no lock file).  Their published algorithms are restated
  here from GPflow 2.6-2.9; gradients come from torch autograd, which plays
  the role of TF's reverse-mode autodiff,
* the Keras-flavoured Adam update that ``run_adam`` applies (``fcest/helpers/inference.py:72-74``).

PARITY UNPINNED: TensorFlow and GPflow are not installable in the build container (no network) and
the reference's own tests pin no numeric value on this path, so this
restatement is checked only against (i) the reference's algebraic identity tests
(``tests/fcest/models/test_likelihoods.py:54-76``), (ii) finite differences, (iii) an independent
closed-form gradient derivation and (iv) independent anchors for the GPflow formulas: scikit-learn's
Matern(nu=2.5), the un-whitened sparse-GP predictive algebra, the generic Gaussian KL, dense Gaussian algebra
for the VGP moments and torch.optim.Adam at epsilon = 0 (``tests/test_oracle_cpu.py``).  Noise draws are *injected* (``eps``) instead of drawn from TF's RNG.
"""
from __future__ import annotations

import math

import numpy as np
import torch

JITTER = 1e-6  # gpflow.config.default_jitter()
DT = torch.float64


def softplus(u):
    return torch.nn.functional.softplus(u)


def inv_softplus(x: float) -> float:
    """likelihoods.py:76-78 -- log(exp(x) - 1)."""
    return float(np.log(np.exp(x) - 1.0))


# --------------------------------------------------------------------------------------------
# GPflow restatements
# --------------------------------------------------------------------------------------------
def square_distance(X, X2):
    """gpflow.utilities.ops.square_distance: -2 X X2^T + |X|^2 + |X2|^2."""
    Xs = (X * X).sum(-1)
    X2s = (X2 * X2).sum(-1)
    return -2.0 * X @ X2.T + Xs[:, None] + X2s[None, :]


def matern52(X, X2, variance, lengthscale):
    """gpflow.kernels.Matern52.K: r = sqrt(max(r2, 1e-36)); var*(1+sqrt5 r+5/3 r^2) exp(-sqrt5 r)."""
    r2 = square_distance(X / lengthscale, X2 / lengthscale)
    r = torch.sqrt(torch.clamp(r2, min=1e-36))
    s5 = math.sqrt(5.0)
    return variance * (1.0 + s5 * r + 5.0 / 3.0 * r * r) * torch.exp(-s5 * r)


def base_conditional_white(Kmn, Kmm, knn, q_mu, q_sqrt, chunk=None):
    """gpflow.conditionals.util.base_conditional(white=True, full_cov=False), q_sqrt (R,M,M).

    Materialises LTA = tril(q_sqrt)^T A exactly as GPflow does (optionally in chunks over the
    latent axis to bound host memory -- same arithmetic per element).
    Returns (mean (N,R), var (N,R)).
    """
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)  # (M,N)
    fvar0 = knn - (A * A).sum(0)  # (N,)
    fmean = A.T @ q_mu  # (N,R)
    R = q_sqrt.shape[0]
    chunk = R if chunk is None else chunk
    cols = []
    for r0 in range(0, R, chunk):
        L = torch.tril(q_sqrt[r0:r0 + chunk])  # (c,M,M)
        LTA = L.transpose(-1, -2) @ A[None]  # (c,M,N)
        cols.append(fvar0[None, :] + (LTA * LTA).sum(-2))  # (c,N)
    fvar = torch.cat(cols, 0).T  # (N,R)
    return fmean, fvar


def gauss_kl_white(q_mu, q_sqrt):
    """gpflow.kullback_leiblers.gauss_kl(q_mu, q_sqrt, K=None)."""
    Lq = torch.tril(q_sqrt)
    mahalanobis = (q_mu * q_mu).sum()
    constant = -float(q_mu.numel())
    logdet_qcov = torch.log(torch.diagonal(q_sqrt, dim1=-2, dim2=-1) ** 2).sum()
    trace = (Lq * Lq).sum()
    return 0.5 * (mahalanobis + constant - logdet_qcov + trace)


def svgp_predict_f(x_new, p, chunk=None):
    """gpflow.models.SVGP.predict_f (whiten=True, zero mean function)."""
    var, ls = softplus(p["u_var"]), softplus(p["u_ls"])
    Z = p["Z"]
    Kmm = matern52(Z, Z, var, ls) + JITTER * torch.eye(Z.shape[0], dtype=DT)
    Kmn = matern52(Z, x_new, var, ls)
    knn = var * torch.ones(x_new.shape[0], dtype=DT)
    return base_conditional_white(Kmn, Kmm, knn, p["q_mu"], p["q_sqrt"], chunk)


def vgp_fmean_fvar(x, p):
    """gpflow.models.VGP.elbo: K=k(X,X)+jitter; L=chol K; fmean=L q_mu; fvar=sum((L tril(q_sqrt))^2)."""
    var, ls = softplus(p["u_var"]), softplus(p["u_ls"])
    K = matern52(x, x, var, ls) + JITTER * torch.eye(x.shape[0], dtype=DT)
    L = torch.linalg.cholesky(K)
    fmean = L @ p["q_mu"]
    LTA = L[None] @ torch.tril(p["q_sqrt"])  # (R,N,N)
    fvar = (LTA * LTA).sum(2).T
    return fmean, fvar


def vgp_predict_f(x_new, x_train, p):
    """gpflow.models.VGP.predict_f: base_conditional(k(X,X*), k(X,X)+jitter, kdiag, q_mu, q_sqrt, white)."""
    var, ls = softplus(p["u_var"]), softplus(p["u_ls"])
    Kmm = matern52(x_train, x_train, var, ls) + JITTER * torch.eye(x_train.shape[0], dtype=DT)
    Kmn = matern52(x_train, x_new, var, ls)
    knn = var * torch.ones(x_new.shape[0], dtype=DT)
    return base_conditional_white(Kmn, Kmm, knn, p["q_mu"], p["q_sqrt"])


# --------------------------------------------------------------------------------------------
# FCEst likelihood (likelihoods.py)
# --------------------------------------------------------------------------------------------
def add_diagonal_additive_noise(cov, additive_part):
    """likelihoods.py:253-264."""
    return cov + torch.diag_embed(softplus(additive_part).expand(cov.shape[:-1]))


def construct_cov(f_sample, A, additive_part, a_matmul=False):
    """likelihoods.py:176-180 / wishart_process.py:207-210: (A o F)(A o F)^T + diag(softplus(lam)).

    ``A`` is (D,D) or (D,) and is applied as a *broadcast Hadamard product* (the shipped code,
    ``tf.multiply``); ``a_matmul=True`` gives the commented-out intent ``tf.matmul(A, F)``.
    """
    af = A @ f_sample if a_matmul else A * f_sample
    affa = af @ af.transpose(-1, -2)
    return add_diagonal_additive_noise(affa, additive_part)


def log_prob(f_sample, y, A, additive_part, a_matmul=False):
    """likelihoods.py:144-206.  f_sample (S,N,D,nu), y (N,D) -> (N,)."""
    D = y.shape[1]
    constant_term = -D / 2 * math.log(2 * math.pi)
    affa = construct_cov(f_sample, A, additive_part, a_matmul)
    L = torch.linalg.cholesky(affa)
    log_det = 2 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(2)  # (S,N)
    yt = y[None, :, :, None].expand(f_sample.shape[0], -1, -1, -1)
    Lsy = torch.linalg.solve_triangular(L, yt, upper=False)
    yaffay = (Lsy ** 2).sum((2, 3))
    ll = constant_term - 0.5 * log_det - 0.5 * yaffay
    return ll.mean(0)


def variational_expectations(f_mean, f_var, y, A, additive_part, eps, a_matmul=False):
    """likelihoods.py:126-142 with the N(0,1) draw injected: eps (S,N,R)."""
    S, N, R = eps.shape
    D = y.shape[1]
    f_sample = eps * f_var ** 0.5 + f_mean
    f_sample = f_sample.reshape(S, N, D, -1)
    return log_prob(f_sample, y, A, additive_part, a_matmul)


def svgp_elbo(p, x, y, eps, chunk=None):
    """gpflow.models.SVGP.elbo with FCEst's likelihood: sum(var_exp) - KL (scale=1)."""
    fmean, fvar = svgp_predict_f(x, p, chunk)
    ve = variational_expectations(fmean, fvar, y, p["A"], p["lam"], eps)
    return ve.sum() - gauss_kl_white(p["q_mu"], p["q_sqrt"])


def vgp_elbo(p, x, y, eps):
    """gpflow.models.VGP.elbo with FCEst's likelihood."""
    fmean, fvar = vgp_fmean_fvar(x, p)
    ve = variational_expectations(fmean, fvar, y, p["A"], p["lam"], eps)
    return ve.sum() - gauss_kl_white(p["q_mu"], p["q_sqrt"])


TRAINABLE = ("u_var", "u_ls", "Z", "q_mu", "q_sqrt", "A", "lam")


def elbo_and_grad(kind, p, x, y, eps, names=TRAINABLE):
    """One ELBO + gradient evaluation through autograd (stand-in for TF autodiff).

    Gradient of q_sqrt is reported on the lower triangle only (GPflow's FillTriangular
    parametrisation has no upper-triangle variables).
    """
    q = {k: v.detach().clone().requires_grad_(k in names and (k != "Z" or kind == "svgp"))
         for k, v in p.items()}
    elbo = svgp_elbo(q, x, y, eps) if kind == "svgp" else vgp_elbo(q, x, y, eps)
    elbo.backward()
    g = {k: (torch.zeros_like(v) if v.grad is None else v.grad) for k, v in q.items() if v.requires_grad}
    if "q_sqrt" in g:
        g["q_sqrt"] = torch.tril(g["q_sqrt"])
    return float(elbo), g


def svgp_elbo_and_grad_chunked(p, x, y, eps, chunk=250):
    """Same value/gradient as ``elbo_and_grad('svgp', ...)`` with the (R,M,N) LTA tensor
    materialised ``chunk`` latents at a time (host-memory bound only; the arithmetic per element
    is unchanged).  Used for config-4-sized fixtures and the CPU baseline."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    with torch.no_grad():
        fmean, fvar = svgp_predict_f(x, q, chunk)
    fm = fmean.clone().requires_grad_(True)
    fv = fvar.clone().requires_grad_(True)
    ve = variational_expectations(fm, fv, y, q["A"], q["lam"], eps)
    ve.sum().backward()
    g_fm, g_fv = fm.grad, fv.grad
    var, ls = softplus(q["u_var"]), softplus(q["u_ls"])
    Z = q["Z"]
    M = Z.shape[0]
    R = q["q_sqrt"].shape[0]
    total = float(ve.sum())
    # conditional backward, chunk by chunk (retain kernel graph by rebuilding it per chunk)
    for r0 in range(0, R, chunk):
        Kmm = matern52(Z, Z, softplus(q["u_var"]), softplus(q["u_ls"])) + JITTER * torch.eye(M, dtype=DT)
        Kmn = matern52(Z, x, softplus(q["u_var"]), softplus(q["u_ls"]))
        knn = softplus(q["u_var"]) * torch.ones(x.shape[0], dtype=DT)
        sl = slice(r0, r0 + chunk)
        mu_c, v_c = base_conditional_white(Kmn, Kmm, knn, q["q_mu"][:, sl], q["q_sqrt"][sl])
        ((mu_c * g_fm[:, sl]).sum() + (v_c * g_fv[:, sl]).sum()).backward()
    kl = gauss_kl_white(q["q_mu"], q["q_sqrt"])
    (-kl).backward()
    g = {k: v.grad for k, v in q.items()}
    g["q_sqrt"] = torch.tril(g["q_sqrt"])
    return total - float(kl), g


# --------------------------------------------------------------------------------------------
# prediction (wishart_process.py)
# --------------------------------------------------------------------------------------------
def cov_samples(f_mean_new, f_var_new, A, additive_part, eps_new, D):
    """wishart_process.py:195-210 / :486-502.  eps_new (S*,N*,D,nu) -> (S*,N*,D,D)."""
    n = f_mean_new.shape[0]
    fm = f_mean_new.reshape(n, D, -1)
    fs = f_var_new.reshape(n, D, -1) ** 0.5
    f_sample = eps_new * fs + fm
    return construct_cov(f_sample, A, additive_part)


def convert_tensor_to_correlation(cov):
    """array_operations.py:56-72."""
    d = torch.diagonal(cov, dim1=-2, dim2=-1).unsqueeze(-1)
    return cov / torch.sqrt(d @ d.transpose(-1, -2))


def predict_cov(f_mean_new, f_var_new, A, additive_part, eps_new, D):
    """wishart_process.py:135-141: mean and population std (tf.math.reduce_std) over samples."""
    s = cov_samples(f_mean_new, f_var_new, A, additive_part, eps_new, D)
    return s.mean(0), s.std(0, unbiased=False)


def predict_corr(f_mean_new, f_var_new, A, additive_part, eps_new, D):
    """wishart_process.py:160-167."""
    s = convert_tensor_to_correlation(cov_samples(f_mean_new, f_var_new, A, additive_part, eps_new, D))
    return s.mean(0), s.std(0, unbiased=False)


def are_all_positive_definite(covs) -> bool:
    """array_operations.py:75-103."""
    for c in covs:
        c = np.asarray(c)
        if not np.allclose(c - c.T, 0, rtol=1e-6):
            return False
        try:
            np.linalg.cholesky(c)
        except np.linalg.LinAlgError:
            return False
    return True


# --------------------------------------------------------------------------------------------
# Keras Adam (tf.optimizers.Adam(learning_rate=1e-3), inference.py:72-74)
# --------------------------------------------------------------------------------------------
def keras_adam_step(theta, grad, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras-Adam update on an unconstrained variable; ``t`` is the 1-based step index."""
    m = m + (grad - m) * (1 - b1)
    v = v + (grad * grad - v) * (1 - b2)
    alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - alpha * m / (torch.sqrt(v) + eps), m, v


# --------------------------------------------------------------------------------------------
# independent closed-form likelihood gradient ("likelihood grad")
# --------------------------------------------------------------------------------------------
def likelihood_closed_form_grad(f_mean, f_var, y, A, lam, eps):
    """Returns d(sum_n var_exp[n]) / d(f_mean, f_var, A, lam) without autograd."""
    S, N, R = eps.shape
    D = y.shape[1]
    w = 1.0 / S
    sd = f_var ** 0.5
    F = (eps * sd + f_mean).reshape(S, N, D, -1)
    G = A * F
    Sig = G @ G.transpose(-1, -2) + torch.diag_embed(softplus(lam).expand(S, N, D))
    Sinv = torch.linalg.inv(Sig)
    alpha = (Sinv @ y[None, :, :, None])  # (S,N,D,1)
    Gbar = w * (-(Sinv @ G) + alpha @ (alpha.transpose(-1, -2) @ G))
    Fbar = (A * Gbar)
    if A.dim() == 2:
        Abar = (F * Gbar).sum((0, 1))
    else:
        Abar = (F * Gbar).sum((0, 1, 2))
    dSinv = torch.diagonal(Sinv, dim1=-2, dim2=-1)
    lam_bar = torch.sigmoid(lam) * (w * (-0.5 * dSinv + 0.5 * alpha[..., 0] ** 2)).sum((0, 1))
    Fbar = Fbar.reshape(S, N, R)
    mu_bar = Fbar.sum(0)
    v_bar = (Fbar * eps).sum(0) / (2 * sd)
    return mu_bar, v_bar, Abar, lam_bar


# --------------------------------------------------------------------------------------------
# canonical synthetic inputs
# --------------------------------------------------------------------------------------------
def make_inputs(N, D, S, M=None, nu=None, seed=0, perturbed=True, a_option="train_full_matrix"):
    """Synthetic data + parameters at the reference initialisation, optionally perturbed.

    M=None -> VGP layout (q_mu (N,R), q_sqrt (R,N,N)); else SVGP with Z = linspace(0,1,M).
    """
    nu = D if nu is None else nu
    R = D * nu
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 1, N)[:, None]
    y = rng.standard_normal((N, D))
    Mq = N if M is None else M
    Z = x.copy() if M is None else np.linspace(0, 1, M)[:, None]
    q_mu = np.zeros((Mq, R))
    q_sqrt = np.tile(1e-3 * np.eye(Mq)[None], (R, 1, 1))
    if a_option == "train_full_matrix":
        A = np.eye(D)
    elif a_option == "scaled":
        A = np.ones(D) / nu
    else:
        A = np.ones(D)
    lam = inv_softplus(0.01) * np.ones(D)
    if perturbed:
        q_mu = q_mu + 0.1 * rng.standard_normal(q_mu.shape)
        q_sqrt = q_sqrt + 1e-2 * np.tril(rng.standard_normal(q_sqrt.shape))
        A = A + 0.05 * rng.standard_normal(A.shape)
        lam = lam + 0.1 * rng.standard_normal(lam.shape)
    eps = rng.standard_normal((S, N, R))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=DT)
    p = {"u_var": t(inv_softplus(1.0)), "u_ls": t(inv_softplus(0.3)), "Z": t(Z), "q_mu": t(q_mu),
         "q_sqrt": t(q_sqrt), "A": t(A), "lam": t(lam)}
    return t(x), t(y), t(eps), p
