#!/usr/bin/env python
"""Likelihood-kernel microbenchmark at the C4 shape (N=1200, D=nu=50, S=8): device time of
fcest_wishart_loglik (forward + backward) with CUDA events, inputs larger than L2 cycled between
iterations, against the fp64 roofline.  FCEST_LIK_IMPL=cta selects the CTA-per-time-step kernel.

  python tools/bench_lik.py [--N 1200 --D 50 --S 8 --iters 20] [--check]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1200)
    ap.add_argument("--D", type=int, default=50)
    ap.add_argument("--S", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fwd", action="store_true")
    args = ap.parse_args()
    from fcest_b200 import ops
    N, D, S = args.N, args.D, args.S
    R = D * D
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    nset = 3  # 3 x 192 MB of eps > 126 MB L2: consecutive iterations never re-use cached noise
    fm = ops.pitched(N, R, dev); fm.copy_(0.3 * torch.randn(N, R, dtype=torch.float64, device=dev, generator=g))
    fv = ops.pitched(N, R, dev); fv.copy_(0.05 + torch.rand(N, R, dtype=torch.float64, device=dev, generator=g))
    y = torch.randn(N, D, dtype=torch.float64, device=dev, generator=g)
    A = torch.eye(D, dtype=torch.float64, device=dev) + 0.2 * torch.randn(D, D, dtype=torch.float64, device=dev, generator=g)
    lam = torch.full((D,), -4.6, dtype=torch.float64, device=dev)
    eps = [torch.randn(S, N, R, dtype=torch.float64, device=dev, generator=g) for _ in range(nset)]
    for i in range(3):
        out = ops.wishart_loglik(fm, fv, eps[i % nset], y, A, lam, need_grad=not args.fwd)
    torch.cuda.synchronize()
    evs = []
    for i in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.wishart_loglik(fm, fv, eps[i % nset], y, A, lam, need_grad=not args.fwd)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    flops = 11.0 / 3.0 * D ** 3 * S * N
    peak = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")))["fp64_tflops"]
    res = {"impl": os.environ.get("FCEST_LIK_IMPL", "warp"), "N": N, "D": D, "S": S, "ms_median": float(np.median(ms)),
           "ms_min": float(ms.min()), "algorithmic_flops": flops,
           "tflops": flops / (np.median(ms) * 1e-3) * 1e-12, "frac_fp64_peak": flops / (np.median(ms) * 1e-3) * 1e-12 / peak,
           "note": "includes the two colsum reductions (A_bar, lam_bar); events around the C-ABI call"}
    if args.check:
        from oracle import wishart_oracle as wo
        n = min(N, 16)
        ve_ref = wo.variational_expectations(fm[:n].cpu(), fv[:n].cpu(), y[:n].cpu(), A.cpu(), lam.cpu(), eps[(args.iters - 1) % nset][:, :n].cpu())
        res["ve_rel_err_first16"] = float(((out["ve"][:n].cpu() - ve_ref).abs().max() / ve_ref.abs().max()).item())
        res["info_max"] = int(out["info"].max().item())
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
