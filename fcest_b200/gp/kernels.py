"""``gpflow.kernels`` mirror: only Matern52 (FCEst's default and the only kernel its save/load supports,
reference ``fcest/models/wishart_process.py:99-100,246,326``)."""
from __future__ import annotations

import torch

from .. import ops
from .parameter import Parameter


class Kernel:
    pass


class Matern52(Kernel):
    """k(x,x') = var (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = |x-x'|/lengthscale (scalar lengthscale)."""

    def __init__(self, variance: float = 1.0, lengthscales: float = 1.0, device="cuda"):
        self.variance = Parameter(variance, transform="softplus", device=device, name="kernel.variance")
        self.lengthscales = Parameter(lengthscales, transform="softplus", device=device,
                                      name="kernel.lengthscales")

    def constrained(self):
        """(variance, lengthscale) as 1-element device tensors."""
        return self.variance.value().reshape(1), self.lengthscales.value().reshape(1)

    def K(self, X: torch.Tensor, X2: torch.Tensor | None = None, jitter: float = 0.0) -> torch.Tensor:
        X2 = X if X2 is None else X2
        kvar, kls = self.constrained()
        out = ops.pitched(X.shape[0], X2.shape[0], X.device)
        return ops.matern52(X, X2, kvar, kls, out, jitter)

    def __call__(self, X, X2=None):
        return self.K(X, X2)
