"""Minimal stand-in for ``gpflow.Parameter`` / ``tf.Variable`` holding device-resident float64 state.

Mirrors the slice of the GPflow API that FCEst touches (``.numpy()``, ``.assign()``, ``trainable``):
the *unconstrained* tensor is what the optimiser updates (reference ``fcest/helpers/inference.py:113-116``
minimises over ``model.trainable_variables``), the constrained value is what ``.numpy()`` returns and
what ``save_model_params_dict`` serialises (``fcest/models/wishart_process.py:257-266``).
Transforms: ``None`` (identity), ``"softplus"`` (gpflow.utilities.positive()), ``"tril"``
(gpflow.utilities.triangular(): only the lower triangle is free).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

F64 = torch.float64


def _inv_softplus(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 30.0, x, np.log(np.expm1(np.minimum(x, 30.0))))


class Parameter:
    def __init__(self, value, transform: str | None = None, trainable: bool = True, device="cuda",
                 name: str = ""):
        self.transform = transform
        self.trainable = trainable
        self.name = name
        self.device = torch.device(device)
        self.unconstrained = None
        self.grad = None  # d ELBO / d unconstrained, filled by the model's backward pass
        self.assign(value)

    # -- tf.Variable / gpflow.Parameter surface -------------------------------------------------
    def assign(self, value) -> None:
        if isinstance(value, torch.Tensor) and value.is_cuda and self.transform in (None, "tril"):
            # device-resident value (e.g. the 0.8 GB q_sqrt of a D = 200 subject): no host round trip
            t = value.detach().to(device=self.device, dtype=F64)
            t = torch.tril(t) if self.transform == "tril" else t.clone()
            t = t.contiguous()
            if self.unconstrained is not None and self.unconstrained.shape == t.shape:
                self.unconstrained.copy_(t)
            else:
                self.unconstrained = t
            return
        v = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
        v = np.array(v, dtype=np.float64)
        if self.transform == "softplus":
            v = _inv_softplus(v)
        elif self.transform == "tril":
            v = np.tril(v)
        t = torch.tensor(v, dtype=F64, device=self.device).contiguous()
        if self.unconstrained is not None and self.unconstrained.shape == t.shape:
            self.unconstrained.copy_(t)
        else:
            self.unconstrained = t

    def numpy(self) -> np.ndarray:
        return self.value().detach().cpu().numpy()

    def value(self) -> torch.Tensor:
        """Constrained value as a device tensor."""
        if self.transform == "softplus":
            out = torch.empty_like(self.unconstrained)
            return ops.softplus_fwd(self.unconstrained, out)
        return self.unconstrained

    @property
    def shape(self):
        return tuple(self.unconstrained.shape)

    def __repr__(self):
        return f"Parameter(name={self.name!r}, shape={self.shape}, transform={self.transform}, trainable={self.trainable})"
