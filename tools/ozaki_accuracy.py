#!/usr/bin/env python
"""CPU study: accuracy of an Ozaki-style int8-slice emulation of the three fp64 projection GEMMs at C4.

Operand rows are scaled by a power of two to (-1, 1) and cut into balanced 7-bit digits q_s in [-64, 64]
(x = 2^e sum_s q_s 2^(-6-7s)); the product keeps the digit pairs with s + t < ns, each an EXACT integer GEMM
(here: float64 matmul on integer-valued operands, exact below 2^53).  Prints, per GEMM and per ns, the error
against the float64 GEMM relative to the largest output entry and what it does to fvar / the gradients.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wishart_oracle as wo  # noqa: E402


BASE, FIRST = (254.0, 127.0) if os.environ.get("OZAKI_BASE", "254") == "254" else (128.0, 64.0)


def split(x, ns, axis):
    """x (rows, K) scaled per row (axis=1) -> exponents (rows,), list of ns integer-valued float64 slices."""
    m = np.abs(x).max(axis=axis, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))) + 1, 0.0)
    t = x * np.exp2(-e) * FIRST
    out = []
    for s in range(ns):
        q = np.rint(t)
        out.append(q)
        t = (t - q) * BASE
    return e, out


def ozaki_gemm(A, B, ns):
    """A (M,K) @ B (N,K)^T with per-row scaling of both, digit pairs s + t < ns."""
    ea, As = split(A, ns, 1)
    eb, Bs = split(B, ns, 1)
    acc = np.zeros((A.shape[0], B.shape[0]))
    At = [torch.from_numpy(a) for a in As]
    Bt = [torch.from_numpy(b).T.contiguous() for b in Bs]
    for d in range(ns - 1, -1, -1):          # least significant order first
        part = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64)
        for s in range(d + 1):
            part += At[s] @ Bt[d - s]
        acc += part.numpy() * BASE ** (-d) / FIRST ** 2
    return acc * np.exp2(ea) * np.exp2(eb).T


def main():
    N, D, S, M = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1200, 50, 8, 200)))
    x, y, eps, p = wo.make_inputs(N, D, S, M=M, seed=0)
    var, ls = wo.softplus(p["u_var"]), wo.softplus(p["u_ls"])
    Z = p["Z"]
    Kmm = wo.matern52(Z, Z, var, ls) + wo.JITTER * torch.eye(M, dtype=torch.float64)
    P = torch.linalg.solve_triangular(torch.linalg.cholesky(Kmm), wo.matern52(Z, x, var, ls), upper=False).numpy()
    iu = np.tril_indices(M)
    w = np.where(iu[0] == iu[1], 1.0, 2.0)
    Pk = (P[iu[0]] * P[iu[1]] * w[:, None]).T.copy()                   # (N, K)
    L = torch.tril(p["q_sqrt"])
    Sk = (L @ L.transpose(-1, -2)).numpy()[:, iu[0], iu[1]].copy()     # (R, K)
    fmean, fvar = wo.svgp_predict_f(x, p, chunk=125)
    _, g_var, _, _ = wo.likelihood_closed_form_grad(fmean, fvar, y, p["A"], p["lam"], eps)
    g = g_var.numpy()
    print(f"N={N} R={D*D} K={Pk.shape[1]}  |Pk| max/rms {np.abs(Pk).max()/np.sqrt((Pk**2).mean()):.1f}  "
          f"|Sk| max/rms {np.abs(Sk).max()/np.sqrt((Sk**2).mean()):.1f}  |g| max/rms {np.abs(g).max()/np.sqrt((g**2).mean()):.1f}")
    V = Pk @ Sk.T
    Gk = g.T @ Pk
    Tk = g @ Sk
    c = (float(var) - (P * P).sum(0))[:, None]
    for ns in (4, 5, 6, 7):
        t0 = time.time()
        V2 = ozaki_gemm(Pk, Sk, ns)
        G2 = ozaki_gemm(np.ascontiguousarray(g.T), np.ascontiguousarray(Pk.T), ns)
        T2 = ozaki_gemm(g, np.ascontiguousarray(Sk.T), ns)
        fv, fv2 = c + V, c + V2
        print(f"ns={ns}: V rel-to-max {np.abs(V2-V).max()/np.abs(V).max():.2e}  fvar max rel {np.abs((fv2-fv)/fv).max():.2e}  "
              f"Gk rel-to-max {np.abs(G2-Gk).max()/np.abs(Gk).max():.2e}  Tk rel-to-max {np.abs(T2-Tk).max()/np.abs(Tk).max():.2e}  "
              f"sum(g*dV)/|sum g V| {abs((g*(V2-V)).sum())/abs((g*V).sum()):.2e}  ({time.time()-t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
