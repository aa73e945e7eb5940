#!/usr/bin/env python
"""A few launches of the likelihood at a small-D shape (for an ncu launch list): python tools/lik_small_once.py N D S"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops  # noqa: E402

N, D, S = (int(v) for v in sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(0)
R = D * D
fm = ops.pitched(N, R, "cuda"); fm.copy_(0.3 * torch.randn(N, R, dtype=torch.float64, device="cuda", generator=g))
fv = ops.pitched(N, R, "cuda"); fv.copy_(0.05 + torch.rand(N, R, dtype=torch.float64, device="cuda", generator=g))
y = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
A = torch.eye(D, dtype=torch.float64, device="cuda") + 0.2 * torch.randn(D, D, dtype=torch.float64, device="cuda", generator=g)
lam = torch.full((D,), -4.6, dtype=torch.float64, device="cuda")
eps = torch.randn(S, N, R, dtype=torch.float64, device="cuda", generator=g)
for _ in range(4):
    out = ops.wishart_loglik(fm, fv, eps, y, A, lam, need_grad=True)
torch.cuda.synchronize()
print("done", float(out["ve"].sum()))
