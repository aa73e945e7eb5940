"""Whitened GP conditional + KL, forward and analytic backward, on the device.

Host-side mirror of the GPflow code FCEst's models inherit:
``gpflow.models.SVGP.elbo/predict_f``, ``gpflow.models.VGP.elbo/predict_f``,
``gpflow.conditionals.util.base_conditional(white=True, full_cov=False)``,
``gpflow.kullback_leiblers.gauss_kl`` (reached from ``fcest/models/wishart_process.py:109-114,381-386``
and ``:195,486``).  This module only sequences C-ABI calls; all arithmetic runs in
``libfcest_b200.so``.

Formulation ("packed quadratic form"):  with  P = Lm^-1 Kmn  (SVGP / predict) or
P = L^T (VGP.elbo),
    fmean = P^T q_mu ,   fvar[n,r] = c[n] + p_n^T S_r p_n ,  S_r = tril(q_sqrt_r) tril(q_sqrt_r)^T ,
    c[n] = kvar - sum_m P[m,n]^2 (SVGP / predict)  or 0 (VGP.elbo).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import ops

F64 = torch.float64
JITTER = 1e-6  # gpflow.config.default_jitter()


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


_side_streams: dict = {}

# The two independent backward chains (d q_sqrt: GEMM + tri_grad; d P: GEMM + sym_matvec + single-matrix chain) run
# on two streams.  bench.py switches this off for its per-kernel roofline pass, where every launch is timed alone.
OVERLAP_BACKWARD = True


def _side_stream(dev, index: int = 0) -> torch.cuda.Stream:
    """Auxiliary streams per device.  0: the latency-bound single-matrix chain (kernel matrix -> Cholesky -> inverse -> P)
    and the d q_sqrt chain of the backward pass; 1: the Monte-Carlo draw of an evaluation."""
    key = (str(dev), index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


@dataclass
class CondState:
    """Everything the backward pass needs from the forward pass."""
    vgp: bool
    X: torch.Tensor
    Z: torch.Tensor
    kvar: torch.Tensor
    kls: torch.Tensor
    Lm: torch.Tensor
    Li: torch.Tensor
    P: torch.Tensor
    Pk: torch.Tensor
    Sk: torch.Tensor
    q_mu: torch.Tensor
    q_sqrt: torch.Tensor
    info: torch.Tensor


def conditional_forward(kernel, Z, X, q_mu, q_sqrt, vgp: bool, want_kl: bool = True):
    """Returns (fmean (N,R), fvar (N,R), kl (1,) or None, state).

    vgp=False: SVGP conditional of points X on inducing points Z (also VGP.predict_f with Z := X_train).
    vgp=True : VGP.elbo moments at the training inputs (X is Z; P = chol(K)^T, no prior term).
    """
    dev = q_mu.device
    M = Z.shape[0]
    N = M if vgp else X.shape[0]
    R = q_mu.shape[1]
    kvar, kls = kernel.constrained()
    info = torch.zeros(1, dtype=torch.int32, device=dev)

    K = M * (M + 1) // 2
    ldk = _round_up(K, 16)
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        # single-matrix chain (one-CTA Cholesky is latency bound) -- overlapped with tri_syrk_pack below
        Lm = ops.pitched(M, M, dev)
        ops.matern52(Z, Z, kvar, kls, Lm, JITTER)
        Li = ops.pitched(M, M, dev)
        ops.chol_inv_single(Lm, Li, info)
        P = ops.pitched(M, N, dev)
        if vgp:
            ops.transpose(Lm, P)
        else:
            Kmn = ops.pitched(M, N, dev)
            ops.matern52(Z, X, kvar, kls, Kmn, 0.0)
            ops.dgemm(True, False, M, N, M, Li, Kmn, P)
        Pk = torch.empty((N, ldk), dtype=F64, device=dev)
        bias = torch.empty(N, dtype=F64, device=dev)
        ops.call("fcest_pack_outer", ops.ptr(P), ops.ld(P), M, N, ops.ptr(Pk), ldk, ops.ptr(bias),
                 ops.ptr(kvar), 0 if vgp else 1)
    Sk = torch.empty((R, ldk), dtype=F64, device=dev)
    klp = torch.empty((R, 2), dtype=F64, device=dev)
    q_sqrt = q_sqrt.contiguous()
    ops.call("fcest_tri_syrk_pack", ops.ptr(q_sqrt), R, M, ops.ptr(Sk), ldk, ops.ptr(klp))
    main.wait_stream(side)

    fvar = ops.pitched(N, R, dev)
    ops.dgemm(True, True, N, R, K, Pk, Sk, fvar, bias_row=bias)
    q_mu_p = ops.as_pitched(q_mu)
    fmean = ops.pitched(N, R, dev)
    ops.dgemm(False, False, N, R, M, P, q_mu_p, fmean)

    kl = None
    if want_kl:
        # gauss_kl (white): 0.5 (|q_mu|^2 - M R - sum log diag^2 + |tril q_sqrt|_F^2)
        kl = torch.empty(1, dtype=F64, device=dev)
        ops.sumsq_into(q_mu_p, 0.5, kl, False)
        ops.sum_into(klp, R, 2, 0.5, kl, True)
        ops.sum_into(klp.reshape(-1)[1:], R, 2, -0.5, kl, True)
        kl += -0.5 * M * R
    state = CondState(vgp, X, Z, kvar, kls, Lm, Li, P, Pk, Sk, q_mu_p, q_sqrt, info)
    return fmean, fvar, kl, state


def conditional_backward(st: CondState, g_fmean, g_fvar, need_Z: bool, with_kl: bool = True,
                         kl_scale: float = 1.0, gk_reduce=None, q_sqrt_adam: dict | None = None):
    """Adjoint of ``conditional_forward`` (+ minus the KL gradient): given d ELBO / d (fmean, fvar)
    returns d ELBO / d {q_mu, q_sqrt, kvar, kls, Z} (constrained kernel parameters).

    Time-sharded evaluation (``fcest_b200.parallel.TimeShardedSVWP``): every rank calls this on its own
    slice of time points.  ``gk_reduce`` (an in-place SUM all-reduce) is applied to the packed
    d ELBO / d S_r = g_var^T Pk *before* the per-latent ``tri_grad``, so d q_sqrt comes out complete
    and identical on every rank (402 MB on the wire at C4 instead of the 800 MB dense gradient);
    ``kl_scale`` = 1 / world weights the KL part of the *other* gradients, which the caller sums over ranks.

    ``q_sqrt_adam`` = dict(m, v, alpha_dev, b1, b2, eps): apply the optimiser's update to q_sqrt inside
    ``tri_grad``'s epilogue (``fcest_tri_grad_adam``) instead of returning the dense gradient -- SURVEY.md 8f row 1:
    the 8 M^2 R byte gradient never round-trips through HBM.  The returned ``"q_sqrt"`` is then ``None``."""
    dev = g_fmean.device
    M, N = st.P.shape
    R = st.q_mu.shape[1]
    K = M * (M + 1) // 2
    ldk = st.Pk.shape[1]
    kl = 1.0 if with_kl else 0.0

    # d/dS (packed) = g_var^T Pk  ->  d q_sqrt.  This chain (GEMM, optional all-reduce, tri_grad) is independent of
    # the d P chain below: it runs on the auxiliary stream so that the partial last waves of the two large GEMMs
    # and the latency-bound single-matrix kernels at the end of the d P chain overlap with it.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    Gk = torch.empty((R, ldk), dtype=F64, device=dev)
    fuse = q_sqrt_adam is not None and M <= 200 and M % 2 == 0   # what fcest_tri_grad_adam supports
    dq_sqrt = None if fuse else torch.empty_like(st.q_sqrt)
    chain_stream = side if OVERLAP_BACKWARD else main
    if OVERLAP_BACKWARD:
        side.wait_stream(main)
    with torch.cuda.stream(chain_stream):
        ops.dgemm(False, False, R, K, N, g_fvar, st.Pk, Gk, backward=True)
        if gk_reduce is not None:
            gk_reduce(Gk)
        if fuse:
            a = q_sqrt_adam
            ops.call("fcest_tri_grad_adam", ops.ptr(Gk), ldk, ops.ptr(st.q_sqrt), R, M, kl, ops.ptr(a["m"]),
                     ops.ptr(a["v"]), ops.ptr(a["alpha_dev"]), float(a["b1"]), float(a["b2"]), float(a["eps"]))
        else:
            ops.call("fcest_tri_grad", ops.ptr(Gk), ldk, ops.ptr(st.q_sqrt), R, M, ops.ptr(dq_sqrt), kl)
    Gk.record_stream(chain_stream)
    del Gk

    # T_n (packed) = g_var Sk  ->  dP = 2 T_n p_n (- prior term)
    Tk = torch.empty((N, ldk), dtype=F64, device=dev)
    ops.dgemm(True, False, N, K, R, g_fvar, st.Sk, Tk, backward=True)
    dP = ops.pitched(M, N, dev)
    gs = torch.empty(N, dtype=F64, device=dev)
    ops.call("fcest_sym_matvec", ops.ptr(Tk), ldk, ops.ptr(st.P), ops.ld(st.P), M, N, ops.ptr(g_fvar),
             ops.ld(g_fvar), R, 0 if st.vgp else 1, ops.ptr(dP), ops.ld(dP), ops.ptr(gs), 0)
    del Tk
    # mean path: dP += q_mu g_mean^T ; dq_mu = P g_mean - kl * q_mu
    ops.dgemm(True, True, M, N, R, st.q_mu, g_fmean, dP, beta=1.0)
    dq_mu = ops.pitched(M, R, dev)
    ops.dgemm(True, False, M, R, N, st.P, g_fmean, dq_mu)
    if with_kl:
        flat_q = st.q_mu if st.q_mu.is_contiguous() else None
        if flat_q is not None and dq_mu.is_contiguous():
            ops.axpby(-kl_scale, flat_q, 1.0, dq_mu)
        else:  # odd R: pitched views, do it row-block wise through a contiguous temp
            tmp = dq_mu.contiguous()
            ops.axpby(-kl_scale, st.q_mu.contiguous(), 1.0, tmp)
            dq_mu = tmp

    g_kvar = torch.zeros(1, dtype=F64, device=dev)
    g_kls = torch.zeros(1, dtype=F64, device=dev)
    if not st.vgp:
        ops.sum_into(gs, N, 1, 1.0, g_kvar, False)  # knn = kvar

    # chain through P
    X2 = ops.pitched(M, M, dev)
    dKmn = None
    if st.vgp:
        ops.transpose(dP, X2)  # dL = dP^T
    else:
        dKmn = ops.pitched(M, N, dev)
        ops.dgemm(False, False, M, N, M, st.Li, dP, dKmn)       # Li^T dP
        X1 = ops.pitched(M, M, dev)
        ops.dgemm(True, True, M, M, N, dP, st.P, X1)            # dP P^T
        ops.dgemm(False, False, M, M, M, st.Li, X1, X2, alpha=-1.0)  # dLm = -Li^T (dP P^T)
    ops.tril_(X2, 0)
    # Cholesky adjoint: dK = Li^T Phi(Lm^T dLm) Li, symmetrised inside matern52_bwd
    X3 = ops.pitched(M, M, dev)
    ops.dgemm(False, False, M, M, M, st.Lm, X2, X3)
    ops.tril_(X3, 1)
    X4 = ops.pitched(M, M, dev)
    ops.dgemm(False, False, M, M, M, st.Li, X3, X4)
    X5 = ops.pitched(M, M, dev)
    ops.dgemm(True, False, M, M, M, X4, st.Li, X5)

    gZ = None
    if need_Z and not st.vgp:
        gZ = torch.zeros_like(st.Z)
        ops.matern52_bwd(st.Z, st.Z, st.kvar, st.kls, X5, True, g_kvar, g_kls, gZ, True)
        gZ2 = torch.zeros_like(st.Z)
        ops.matern52_bwd(st.Z, st.X, st.kvar, st.kls, dKmn, False, g_kvar, g_kls, gZ2, True)
        ops.axpby(1.0, gZ2, 1.0, gZ)
    else:
        ops.matern52_bwd(st.Z, st.Z, st.kvar, st.kls, X5, True, g_kvar, g_kls, None, True)
        if not st.vgp:
            ops.matern52_bwd(st.Z, st.X, st.kvar, st.kls, dKmn, False, g_kvar, g_kls, None, True)
    if OVERLAP_BACKWARD:
        main.wait_stream(side)
    return {"q_mu": dq_mu[:, :R], "q_sqrt": dq_sqrt, "kvar": g_kvar, "kls": g_kls, "Z": gZ}
