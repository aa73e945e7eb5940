"""Tensor-level wrappers over the C ABI (device float64 tensors in, device tensors out).

Everything here is plumbing: allocate pitched device buffers with torch, pass raw pointers and the
current CUDA stream to ``libfcest_b200.so``.  No arithmetic happens in Python.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import call, ptr

F64 = torch.float64


def even(n: int) -> int:
    return n + (n & 1)


def pitched(rows: int, cols: int, device, zero: bool = False) -> torch.Tensor:
    """(rows, cols) float64 view with an even row pitch (GEMM operands need 16-byte aligned rows)."""
    buf = (torch.zeros if zero else torch.empty)((rows, even(cols)), dtype=F64, device=device)
    return buf[:, :cols]


def as_pitched(t: torch.Tensor) -> torch.Tensor:
    """Return ``t`` itself if it already is a legal GEMM operand, else an even-pitch copy."""
    assert t.dim() == 2 and t.dtype == F64
    if t.stride(1) == 1 and t.stride(0) % 2 == 0 and t.data_ptr() % 16 == 0:
        return t
    out = pitched(t.shape[0], t.shape[1], t.device, zero=True)
    out.copy_(t)
    return out


def ld(t: torch.Tensor) -> int:
    assert t.stride(-1) == 1, "innermost dimension must be contiguous"
    return t.stride(0)


_ws_cache: dict = {}
_ws_retired: list = []   # outgrown scratch buffers stay alive: a captured CUDA graph may hold their raw pointers


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)  # one scratch buffer per stream
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        if buf is not None:
            _ws_retired.append(buf)   # never hand memory a graph replay still writes back to the allocator
        buf = torch.empty((max(nbytes, 1 << 20) + 7) // 8, dtype=F64, device=device)
        _ws_cache[key] = buf
    return buf


# Large GEMMs go to the int8 tensor cores (tcgen05 Ozaki-style emulation, ``fcest_dgemm_ozaki``) unless
# FCEST_GEMM_IMPL=dmma; FCEST_OZAKI_SLICES (2..7, default 6 = 47.9 bits) sets the number of base-254 digits per operand.
# The two backward GEMMs (gradients, held to 1e-7 relative-to-max) use FCEST_OZAKI_SLICES_BWD digits (default 5 = 39.9
# bits: ~3e-11 of the largest entry).
GEMM_IMPL = os.environ.get("FCEST_GEMM_IMPL", "ozaki")
OZAKI_SLICES = int(os.environ.get("FCEST_OZAKI_SLICES", "6"))
OZAKI_SLICES_BWD = int(os.environ.get("FCEST_OZAKI_SLICES_BWD", "5"))
OZAKI_MIN_FLOPS = float(os.environ.get("FCEST_OZAKI_MIN_FLOPS", "4e9"))
OZAKI_MIN_DIM = int(os.environ.get("FCEST_OZAKI_MIN_DIM", "128"))   # smallest output dimension / contraction length
OZAKI_MIN_K = int(os.environ.get("FCEST_OZAKI_MIN_K", "192"))     # that still goes to the int8 path (C2: 461 -> 667 evals/s)


def dgemm_ozaki(a_kmajor: bool, b_kmajor: bool, M: int, N: int, K: int, A, B, C, alpha=1.0, bias_row=None,
                slices: int | None = None):
    """C[:M,:N] = alpha op(A) op(B) (+ bias_row[m]) on the int8 tensor cores; see ``fcest_dgemm_ozaki``."""
    lib = _lib.load_library()
    slices = OZAKI_SLICES if slices is None else slices
    nbytes = int(lib.fcest_dgemm_ozaki_workspace_bytes(M, N, K, slices))
    if nbytes <= 0:
        raise ValueError(f"fcest_dgemm_ozaki: unsupported problem M={M} N={N} K={K} slices={slices}")
    ws = _workspace(nbytes, C.device)
    call("fcest_dgemm_ozaki", int(a_kmajor), int(b_kmajor), M, N, K, float(alpha), ptr(A), ld(A), ptr(B), ld(B),
         ptr(C), ld(C), ptr(bias_row), int(slices), ptr(ws), ws.numel() * 8, flops=2.0 * M * N * K)
    return C


def dgemm(a_kmajor: bool, b_kmajor: bool, M: int, N: int, K: int, A, B, C, alpha=1.0, beta=0.0,
          bias_row=None, bias_col=None, backward: bool = False):
    """C[:M,:N] = alpha op(A) op(B) + beta C (+ biases); see ``fcest_dgemm`` in the header.  ``backward`` marks a
    gradient GEMM (fewer digits on the int8 path)."""
    if (GEMM_IMPL == "ozaki" and beta == 0.0 and bias_col is None and 2.0 * M * N * K >= OZAKI_MIN_FLOPS
            and min(M, N) >= OZAKI_MIN_DIM and K >= OZAKI_MIN_K):
        return dgemm_ozaki(a_kmajor, b_kmajor, M, N, K, A, B, C, alpha=alpha, bias_row=bias_row,
                           slices=OZAKI_SLICES_BWD if backward else OZAKI_SLICES)
    lib = _lib.load_library()
    nbytes = int(lib.fcest_dgemm_workspace_bytes(M, N, K, ld(C), 1))
    ws = _workspace(nbytes, C.device) if nbytes else None
    call("fcest_dgemm", int(a_kmajor), int(b_kmajor), M, N, K, float(alpha), ptr(A), ld(A), 0, ptr(B),
         ld(B), 0, float(beta), ptr(C), ld(C), 0, ptr(bias_row), ptr(bias_col), 1, ptr(ws),
         ws.numel() * 8 if ws is not None else 0, flops=2.0 * M * N * K)
    return C


def matern52(X1, X2, kvar, kls, out, jitter=0.0):
    call("fcest_matern52", ptr(X1), X1.shape[0], ptr(X2), X2.shape[0], X1.shape[1], ptr(kvar), ptr(kls),
         float(jitter), ptr(out), ld(out))
    return out


def matern52_bwd(X1, X2, kvar, kls, gK, sym, g_var, g_ls, gX1, accumulate):
    part = torch.empty(2 * X1.shape[0], dtype=F64, device=X1.device)
    call("fcest_matern52_bwd", ptr(X1), X1.shape[0], ptr(X2), X2.shape[0], X1.shape[1], ptr(kvar),
         ptr(kls), ptr(gK), ld(gK), int(sym), ptr(part), ptr(g_var), ptr(g_ls), ptr(gX1), int(accumulate))


def chol_single(A, info):
    call("fcest_chol_single", ptr(A), A.shape[0], ld(A), ptr(info))


def tri_inv_single(L, Li):
    call("fcest_tri_inv_single", ptr(L), L.shape[0], ld(L), ptr(Li), ld(Li))


def chol_inv_single(A, Li, info):
    """A -> chol(A) (lower, in place), Li = chol(A)^-1 (blocked single-CTA kernel)."""
    call("fcest_chol_inv_single", ptr(A), A.shape[0], ld(A), ptr(Li), ld(Li), ptr(info))


def tril_(X, mode=0):
    call("fcest_tril", ptr(X), X.shape[0], ld(X), mode)
    return X


def transpose(X, Y):
    call("fcest_transpose", ptr(X), X.shape[0], X.shape[1], ld(X), ptr(Y), ld(Y))
    return Y


def axpby(a, x, b, y):
    """y = a x + b y on contiguous buffers of equal size."""
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    call("fcest_axpby", x.numel(), float(a), ptr(x), float(b), ptr(y))
    return y


def sum_into(x, n, stride, scale, out, accumulate):
    call("fcest_sum", ptr(x), n, stride, float(scale), ptr(out), int(accumulate))


def sumsq_into(x, scale, out, accumulate):
    ws = torch.empty(296, dtype=F64, device=x.device)
    call("fcest_sumsq", ptr(x), x.shape[0], x.shape[1], ld(x), float(scale), ptr(out), int(accumulate), ptr(ws))


def softplus_fwd(u, out):
    call("fcest_softplus_fwd", ptr(u), ptr(out), u.numel())
    return out


def softplus_bwd(u, g_c, g_u):
    call("fcest_softplus_bwd", ptr(u), ptr(g_c), ptr(g_u), u.numel())
    return g_u


def randn(out, seed: int, offset: int = 0):
    assert out.is_contiguous()
    call("fcest_randn", ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1))
    return out


def adam_step(theta, grad, m, v, t, grad_sign=-1.0, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    assert theta.is_contiguous() and grad.is_contiguous() and theta.numel() == grad.numel()
    call("fcest_adam_step", ptr(theta), ptr(grad), ptr(m), ptr(v), theta.numel(), float(grad_sign), int(t),
         float(lr), float(b1), float(b2), float(eps))


def adam_step_dev(theta, grad, m, v, alpha_t, grad_sign=-1.0, b1=0.9, b2=0.999, eps=1e-7):
    """Adam update with the bias-corrected step size in a device scalar (CUDA-graph friendly)."""
    assert theta.is_contiguous() and grad.is_contiguous() and theta.numel() == grad.numel()
    call("fcest_adam_step_dev", ptr(theta), ptr(grad), ptr(m), ptr(v), theta.numel(), float(grad_sign), ptr(alpha_t),
         float(b1), float(b2), float(eps))


def wishart_loglik(fmean, fvar, eps, y, A, lam, need_grad=True):
    """Fused likelihood forward (+ backward).  Returns dict(ve, info[, g_fmean, g_fvar, gA, glam])."""
    N, R = fmean.shape
    S = eps.shape[0]
    D = y.shape[1]
    nu = R // D
    dev = fmean.device
    assert eps.is_contiguous() and eps.shape == (S, N, R) and y.is_contiguous() and A.is_contiguous()
    assert ld(fmean) == ld(fvar)
    lib = _lib.load_library()
    if D > 400:
        raise NotImplementedError(f"fcest_b200: the likelihood kernels support D <= 400 (got D={D}).")
    nscr = int(lib.fcest_wishart_loglik_scratch_bytes(D, nu))
    scratch = torch.empty(nscr // 8, dtype=F64, device=dev) if nscr else None  # large-D tile scratch
    a_mode = 0 if A.dim() == 2 else 1
    ve = torch.empty(N, dtype=F64, device=dev)
    info = torch.zeros(N, dtype=torch.int32, device=dev)
    out = {"ve": ve, "info": info}
    if need_grad:
        g_fmean = pitched(N, R, dev)
        g_fvar = pitched(N, R, dev)
        gA = torch.empty_like(A)
        glam = torch.empty_like(lam)
        ws_a = torch.empty((N, R), dtype=F64, device=dev)
        ws_l = torch.empty((N, D), dtype=F64, device=dev)
        call("fcest_wishart_loglik", ptr(fmean), ptr(fvar), ld(fmean), ptr(eps), ptr(y), ptr(A), a_mode,
             ptr(lam), S, N, D, nu, ptr(ve), ptr(g_fmean), ptr(g_fvar), ld(g_fmean), ptr(gA), ptr(glam),
             ptr(ws_a), ptr(ws_l), ptr(info), ptr(scratch), flops=(3.0 * D * D * nu + 2.0 / 3.0 * D ** 3 + D * D) * S * N)
        out.update(g_fmean=g_fmean, g_fvar=g_fvar, gA=gA, glam=glam)
    else:
        call("fcest_wishart_loglik", ptr(fmean), ptr(fvar), ld(fmean), ptr(eps), ptr(y), ptr(A), a_mode,
             ptr(lam), S, N, D, nu, ptr(ve), None, None, 0, None, None, None, None, ptr(info), ptr(scratch))
    return out
