#!/usr/bin/env python
"""One shape of the int8-emulated GEMM, a few launches (for ncu): python tools/ozaki_one.py {V|Gk|Tk} [slices]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "V"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 7
N, R, K = 1200, 2500, 20100
g = torch.Generator(device="cuda").manual_seed(0)
ldk = (K + 15) // 16 * 16
Pk = torch.randn(N, ldk, dtype=torch.float64, device="cuda", generator=g)[:, :K]
Sk = torch.randn(R, ldk, dtype=torch.float64, device="cuda", generator=g)[:, :K]
gv = ops.as_pitched(torch.randn(N, R, dtype=torch.float64, device="cuda", generator=g))
args = {"V": (True, True, N, R, K, Pk, Sk), "Gk": (False, False, R, K, N, gv, Pk), "Tk": (True, False, N, K, R, gv, Sk)}[which]
out = torch.empty((args[2], (args[3] + 15) // 16 * 16), dtype=torch.float64, device="cuda")[:, :args[3]]
for _ in range(3):
    ops.dgemm_ozaki(*args, out, slices=ns)
torch.cuda.synchronize()
print("done", which, ns, float(out[0, 0]))
