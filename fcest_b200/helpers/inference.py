"""``fcest.helpers.inference`` mirror: ``run_adam`` (reference ``fcest/helpers/inference.py:14-126``).

The optimiser is the Keras-flavoured Adam that ``tf.optimizers.Adam(learning_rate=0.001)`` applies
to the *unconstrained* variables (epsilon 1e-7, bias correction folded into the step size), run by
the ``fcest_adam_step`` kernel on the gradients that ``model.elbo_and_grad`` leaves on each Parameter.
"""
from __future__ import annotations

import logging

import torch

from .. import ops


class KerasAdam:
    """State (m, v, step) for a list of Parameters; ``step()`` consumes ``Parameter.grad`` (d ELBO / d theta)."""

    def __init__(self, parameters, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7):
        self.parameters = list(parameters)
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        self.m = [torch.zeros_like(p.unconstrained) for p in self.parameters]
        self.v = [torch.zeros_like(p.unconstrained) for p in self.parameters]

    def step(self) -> None:
        self.t += 1
        for p, m, v in zip(self.parameters, self.m, self.v):
            g = p.grad
            if g is None:
                continue
            g = g if g.is_contiguous() else g.contiguous()
            ops.adam_step(p.unconstrained, g.reshape(p.unconstrained.shape), m, v, self.t, grad_sign=-1.0,
                          lr=self.lr, b1=self.b1, b2=self.b2, eps=self.eps)


def run_adam(
    model_type: str,
    model,
    iterations: int,
    log_interval: int,
    log_dir: str = None,
    data: tuple = None,
    train_inducing_variables: bool = True,
    minibatch_size: int = None,
) -> list:
    """Run Adam on the model's ELBO; returns the ELBOs logged every ``log_interval`` steps.

    Same signature and return value as the reference.  ``log_dir`` (TensorBoard monitoring in the
    reference, ``inference.py:77-109``) is accepted and ignored with a log message.
    """
    if not train_inducing_variables:
        model.inducing_variable.Z.trainable = False

    if minibatch_size is not None:
        raise NotImplementedError("Minibatch training is not yet implemented.")

    logf = []
    match model_type:
        case "SVWP":
            step_data = data
        case "VWP":
            step_data = None
        case _:
            raise NotImplementedError(f"Model type '{model_type}' not recognized.")
    training_loss = model.training_loss_closure(step_data, compile=True)
    optimizer = KerasAdam(model.trainable_parameters, learning_rate=0.001)
    if log_dir is not None:
        logging.info("TensorBoard monitoring is not available in fcest_b200; ignoring log_dir.")

    for step in range(iterations):
        training_loss()          # ELBO + gradients (one fused device pass)
        optimizer.step()
        if step % log_interval == 0:
            elbo = float(model.elbo(step_data).item())  # a fresh MC draw, as in inference.py:122-125
            logf.append(elbo)
            logging.info(f"Epoch {step:04d}: ELBO (train) {elbo:.2f}")
    return logf
