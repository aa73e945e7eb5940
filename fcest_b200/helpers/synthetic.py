"""Synthetic inputs of the shapes BASELINE.json names: data, inducing points and a parameter
point at the reference initialisation (``fcest/models/wishart_process.py:214-237``, ``likelihoods.py:75-86``,
``:208-251``), optionally perturbed -- at the initial point A = I makes every Sigma diagonal.

NumPy only.  The draw order is fixed (y, then the perturbations of q_mu, q_sqrt, A, lambda, then epsilon), so two
callers with the same seed hold the same numbers; ``tests/test_abi_cpu.py`` checks that the generator the CPU checker
uses produces the identical arrays."""
from __future__ import annotations

import numpy as np


def inv_softplus(x: float) -> float:
    return float(np.log(np.exp(x) - 1.0))


def make_synthetic(N: int, D: int, S: int, M: int | None = None, nu: int | None = None, seed: int = 0,
                   perturbed: bool = True, a_option: str = "train_full_matrix", want_eps: bool = True) -> dict:
    """dict(x (N,1), y (N,D), Z, q_mu, q_sqrt, A, lam, eps (S,N,R) or None).  ``M=None``: VWP layout (M = N, Z = x)."""
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
    eps = rng.standard_normal((S, N, R)) if want_eps else None
    return {"x": x, "y": y, "Z": Z, "q_mu": q_mu, "q_sqrt": q_sqrt, "A": A, "lam": lam, "eps": eps}
