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

    def fused_state(self, parameter) -> dict | None:
        """Optimiser state of ``parameter`` in the form ``elbo_and_grad(q_sqrt_adam=...)`` takes (the update of
        q_sqrt fused into the backward pass); needs the device-resident step size, i.e. ``advance()`` first."""
        for p, m, v in zip(self.parameters, self.m, self.v):
            if p is parameter:
                if getattr(self, "alpha_dev", None) is None:
                    self.alpha_dev = torch.empty(1, dtype=torch.float64, device=p.unconstrained.device)
                return {"m": m, "v": v, "alpha_dev": self.alpha_dev, "b1": self.b1, "b2": self.b2, "eps": self.eps}
        return None

    # -- CUDA-graph friendly form: the step size lives in a device scalar --------------------------------
    def advance(self) -> None:
        """Advance the step counter and refresh the device-resident bias-corrected step size."""
        self.t += 1
        alpha_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        if getattr(self, "alpha_dev", None) is None:
            self.alpha_dev = torch.empty(1, dtype=torch.float64, device=self.parameters[0].unconstrained.device)
        self.alpha_dev.fill_(alpha_t)

    def step_device(self) -> None:
        """The update of ``step()`` with the step size read from ``alpha_dev`` (call ``advance()`` first)."""
        for p, m, v in zip(self.parameters, self.m, self.v):
            g = p.grad
            if g is None:
                continue
            g = g if g.is_contiguous() else g.contiguous()
            ops.adam_step_dev(p.unconstrained, g.reshape(p.unconstrained.shape), m, v, self.alpha_dev,
                              grad_sign=-1.0, b1=self.b1, b2=self.b2, eps=self.eps)


class GraphedTrainingStep:
    """One training step (ELBO + gradients + Adam update) captured in a CUDA graph.

    The reference compiles its step with ``tf.function`` (``fcest/helpers/inference.py:111-116``); the analogue
    here is a captured graph of the ~60 kernel launches of a step, replayed with one launch.  Per step only the
    Monte-Carlo draw (one Philox launch into a static buffer) and the Adam step size (one fill) are issued
    outside the graph.  The first call runs eagerly (it is also the warm-up), the second captures, later calls
    replay.  Cholesky status is checked by ``check()`` (a host sync) instead of every step.
    """

    def __init__(self, model, data, optimizer: KerasAdam, fuse_q_sqrt: bool = True):
        self.model, self.opt = model, optimizer
        self.fuse = fuse_q_sqrt and getattr(model, "_sparse", False)   # VGP.elbo: M = N is rarely <= 200; keep it simple
        self.x, self.y = model._data(data)
        lik = model.likelihood
        self.eps = torch.empty((lik.num_mc_samples, self.y.shape[0], lik.D * lik.nu), dtype=torch.float64,
                               device=self.y.device)
        self.graph = None
        self.elbo = None
        self.calls = 0

    def _body(self):
        # q_sqrt (all but a few MB of the parameters) is updated inside the backward pass; the rest by step_device()
        fused = self.opt.fused_state(self.model.q_sqrt) if (self.fuse and self.model.q_sqrt.trainable) else None
        elbo = self.model.elbo_and_grad((self.x, self.y), epsilon=self.eps, check=False, q_sqrt_adam=fused)
        self.opt.step_device()
        return elbo

    def __call__(self) -> torch.Tensor:
        lik = self.model.likelihood
        ops.randn(self.eps, lik._seed, lik._draws)
        lik._draws += 1
        self.opt.advance()
        if self.calls == 0:
            self.elbo = self._body()
        elif self.calls == 1:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.elbo = self._body()
            self.graph.replay()
        else:
            self.graph.replay()
        self.calls += 1
        return self.elbo

    def check(self) -> None:
        info = getattr(self.model, "_last_info", None)
        if info is not None:
            self.model.likelihood.check_info(info[0])
            self.model._check_chol(info[1])


class GraphedEvaluation:
    """One ELBO + gradient evaluation captured in a CUDA graph (``model.compile_evaluation``)."""

    def __init__(self, model, data):
        self.model = model
        self.x, self.y = model._data(data)
        lik = model.likelihood
        self.eps = torch.empty((lik.num_mc_samples, self.y.shape[0], lik.D * lik.nu), dtype=torch.float64,
                               device=self.y.device)
        self.graph, self.elbo, self.calls = None, None, 0

    def __call__(self, epsilon=None) -> torch.Tensor:
        lik = self.model.likelihood
        if epsilon is None:
            ops.randn(self.eps, lik._seed, lik._draws)
            lik._draws += 1
        else:
            self.eps.copy_(epsilon.reshape(self.eps.shape))
        if self.calls == 0:      # eager: warm-up of allocator / lazy module loads
            self.elbo = self.model.elbo_and_grad((self.x, self.y), epsilon=self.eps, check=False)
        elif self.calls == 1:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.elbo = self.model.elbo_and_grad((self.x, self.y), epsilon=self.eps, check=False)
            # the graph writes its results into these tensors on every replay
            self._grads = [(p, p.grad) for p in self.model.parameters]
            self._info = self.model._last_info
            self.graph.replay()
        else:
            self.graph.replay()
        if self.graph is not None:   # an eager evaluation in between may have re-pointed Parameter.grad
            for p, g in self._grads:
                p.grad = g
            self.model._last_info = self._info
        self.calls += 1
        return self.elbo

    def check(self) -> None:
        info = getattr(self.model, "_last_info", None)
        if info is not None:
            self.model.likelihood.check_info(info[0])
            self.model._check_chol(info[1])


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

    import os
    use_graph = os.environ.get("FCEST_CUDA_GRAPH", "1") != "0" and iterations > 2
    graphed = GraphedTrainingStep(model, step_data, optimizer) if use_graph else None
    for step in range(iterations):
        if graphed is not None:
            graphed()            # ELBO + gradients + Adam update: one CUDA-graph launch
        else:
            training_loss()      # ELBO + gradients (one fused device pass)
            optimizer.step()
        if step % log_interval == 0:
            if graphed is not None:
                graphed.check()
            elbo = float(model.elbo(step_data).item())  # a fresh MC draw, as in inference.py:122-125
            logf.append(elbo)
            logging.info(f"Epoch {step:04d}: ELBO (train) {elbo:.2f}")
    return logf
