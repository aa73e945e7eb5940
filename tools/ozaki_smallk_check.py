#!/usr/bin/env python
"""Regression probe for the int8 GEMM at short contractions: python tools/ozaki_smallk_check.py NS K N  (run under `timeout`).
Tk-shaped product (A k-major 1200 x K, B contraction-strided K x N)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops
ns, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(0)
M = 1200
B = ops.as_pitched(torch.randn(K, N, dtype=torch.float64, generator=g).cuda())
A = ops.as_pitched(torch.randn(M, K, dtype=torch.float64, generator=g).cuda())
ref = ops.pitched(M, N, "cuda"); got = ops.pitched(M, N, "cuda")
ops.GEMM_IMPL = "dmma"; ops.dgemm(True, False, M, N, K, A, B, ref); ops.GEMM_IMPL = "ozaki"
ops.dgemm_ozaki(True, False, M, N, K, A, B, got, slices=ns)
torch.cuda.synchronize()
print("ok ns", ns, "K", K, "N", N, float((got - ref).abs().max() / ref.abs().max()), flush=True)
