"""NumPy-backed stand-in for the slice of ``tensorflow`` / ``gpflow`` that FCEst's Wishart path calls.

TEST INFRASTRUCTURE (fixture generation only, build container only).  TensorFlow and GPflow cannot be
installed here, so the reference's *own source text* --

* ``fcest/models/likelihoods.py:35-264``  (``WishartProcessLikelihood.__init__``, ``variational_expectations``,
  ``_log_prob``, ``_set_A_scale_matrix``, ``_add_diagonal_additive_noise``),
* ``fcest/models/wishart_process.py:120-212, 239-308, 392-504, 526-602``  (``predict_cov``, ``predict_corr``,
  ``_get_cov_samples``, ``save_model_params_dict``, ``load_from_params_dict``),
* ``fcest/helpers/array_operations.py:56-103``  (``convert_tensor_to_correlation``, ``are_all_positive_definite``)

-- is imported from the read-only mount and executed on top of this module.  Every op below is the documented
TensorFlow semantics of an elementary function (multiply, matmul, cholesky, set_diag, softplus, triangular_solve,
tile, reduce_*, reshape, log, sqrt ...) evaluated in float64 by NumPy / LAPACK.  ``tf.random.normal`` returns
*injected* draws (``push_normal``), which is what "same seeds" means for this code base.

GPflow is only present as empty base classes: FCEst's own arithmetic never calls into it except
``self.predict_f`` (supplied by the harness as a prescribed (mean, variance) pair).
"""
from __future__ import annotations

import sys
import types

import numpy as np


class Tensor(np.ndarray):
    """float64 array that answers ``isinstance(x, tf.Tensor)`` and ``.numpy()``."""

    def numpy(self):
        return np.asarray(self)

    @property
    def T(self):  # tf.Variable has no .T; only used by SVWP.scale_matrix, which the harness does not call
        return np.asarray(self).T.view(Tensor)


class Variable(Tensor):
    def __new__(cls, value, dtype=None, trainable=True):
        obj = np.array(np.asarray(value), dtype=np.float64).view(cls)
        obj.trainable = trainable
        return obj

    def __array_finalize__(self, obj):
        self.trainable = getattr(obj, "trainable", True)

    def assign(self, value):
        value = np.asarray(value, dtype=np.float64)
        if value.shape != self.shape:
            raise ValueError(f"Shapes {self.shape} and {value.shape} are incompatible")   # tf.Variable.assign
        self[...] = value
        return self


def _t(a):
    return np.asarray(a, dtype=np.float64).view(Tensor)


_normal_queue: list = []


def push_normal(eps):
    """Queue the array the next ``tf.random.normal`` call returns (shape checked against the request)."""
    _normal_queue.append(np.asarray(eps, dtype=np.float64))


def _random_normal(shape, mean=0.0, stddev=1.0, dtype=None):
    eps = _normal_queue.pop(0)
    assert tuple(int(s) for s in shape) == eps.shape, (tuple(shape), eps.shape)
    return _t(eps * stddev + mean)


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = np.asarray(a), np.asarray(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return _t(np.matmul(a, b))


def _triangular_solve(matrix, rhs, lower=True, adjoint=False):
    """Forward substitution, batched over the leading axes (tf.linalg.triangular_solve, lower=True)."""
    assert lower and not adjoint
    L, b = np.asarray(matrix), np.asarray(rhs)
    x = np.zeros(np.broadcast_shapes(L.shape[:-2], b.shape[:-2]) + b.shape[-2:])
    n = L.shape[-1]
    for i in range(n):
        acc = b[..., i, :] - np.einsum("...j,...jk->...k", L[..., i, :i], x[..., :i, :])
        x[..., i, :] = acc / L[..., i, i][..., None]
    return _t(x)


def _set_diag(matrix, diagonal):
    out = np.array(matrix, dtype=np.float64)
    n = out.shape[-1]
    idx = np.arange(n)
    out[..., idx, idx] = np.asarray(diagonal)
    return _t(out)


def _softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return _t(np.logaddexp(0.0, x))


def _reduce(fn):
    def f(x, axis=None, keepdims=False):
        return _t(fn(np.asarray(x), axis=axis, keepdims=keepdims))
    return f


def install() -> types.ModuleType:
    """Put the shim (and empty gpflow / rpy2 / statsmodels stubs) into ``sys.modules``; returns the tf module."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tf = mod("tensorflow", Tensor=Tensor, Variable=Variable, float64=np.float64)
    tf.dtypes = mod("tensorflow.dtypes", float64=np.float64)
    tf.constant = lambda v, dtype=None: _t(v)
    tf.ones = lambda shape, dtype=None: _t(np.ones(shape))
    tf.eye = lambda n, dtype=None: _t(np.eye(n))
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), [int(s) for s in shape]))
    tf.multiply = lambda a, b: _t(np.multiply(np.asarray(a), np.asarray(b)))
    tf.matmul = _matmul
    tf.shape = lambda x: np.asarray(np.shape(x))
    tf.tile = lambda x, multiples: _t(np.tile(np.asarray(x), [int(m) for m in multiples]))
    tf.reduce_sum = _reduce(np.sum)
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(np.asarray(x), axis))
    tf.random = mod("tensorflow.random", normal=_random_normal)
    tf.math = mod("tensorflow.math", log=lambda x: _t(np.log(np.asarray(x))), sqrt=lambda x: _t(np.sqrt(np.asarray(x))),
                  softplus=_softplus, reduce_sum=_reduce(np.sum), reduce_mean=_reduce(np.mean),
                  reduce_std=_reduce(np.std))   # np.std: ddof = 0, like tf.math.reduce_std
    tf.linalg = mod("tensorflow.linalg", cholesky=lambda a: _t(np.linalg.cholesky(np.asarray(a))),
                    diag_part=lambda a: _t(np.diagonal(np.asarray(a), axis1=-2, axis2=-1)),
                    set_diag=_set_diag, triangular_solve=_triangular_solve)
    tf.optimizers = mod("tensorflow.optimizers")
    tf.function = lambda f: f
    mod("tensorflow.python"); mod("tensorflow.python.framework")
    mod("tensorflow.python.framework.errors", InvalidArgumentError=np.linalg.LinAlgError)

    class MonteCarloLikelihood:   # gpflow.likelihoods.MonteCarloLikelihood: FCEst only reads the three dims back
        def __init__(self, input_dim=None, latent_dim=None, observation_dim=None):
            self.input_dim, self.latent_dim, self.observation_dim = input_dim, latent_dim, observation_dim

    class _Base:
        def __init__(self, *a, **k):
            pass

    gp = mod("gpflow", set_trainable=lambda *a, **k: None)
    mod("gpflow.likelihoods", MonteCarloLikelihood=MonteCarloLikelihood)
    gp.kernels = mod("gpflow.kernels", Kernel=_Base, Matern52=_Base)
    gp.models = mod("gpflow.models")
    gp.models.vgp = mod("gpflow.models.vgp", VGP=_Base)
    gp.models.svgp = mod("gpflow.models.svgp", SVGP=_Base)
    mod("gpflow.monitor", ModelToTensorBoard=object, Monitor=object, MonitorTaskGroup=object,
        ScalarToTensorBoard=object)
    mod("rpy2"); mod("rpy2.robjects", pandas2ri=None, r=None)
    mod("rpy2.robjects.packages", importr=lambda *a, **k: None)
    mod("rpy2.robjects.conversion", localconverter=None)
    mod("statsmodels"); mod("statsmodels.tsa"); mod("statsmodels.tsa.ar_model", AutoReg=object)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    return tf
