#!/usr/bin/env python
"""Wait-cycle breakdown of the int8-emulated GEMM roles (producer / MMA issuer): python tools/ozaki_diag.py {V|Gk|Tk} [slices]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import _lib, ops  # noqa: E402

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
for _ in range(2):
    ops.dgemm_ozaki(*args, out, slices=ns)
torch.cuda.synchronize()
cnt = torch.zeros(8 * 148, dtype=torch.int64, device="cuda")
lib = _lib.load_library()
lib.fcest_dgemm_ozaki_debug(C.c_void_p(cnt.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ops.dgemm_ozaki(*args, out, slices=ns); b.record()
torch.cuda.synchronize()
lib.fcest_dgemm_ozaki_debug(None)
c = cnt.view(148, 8).double().cpu()
names = ["producer total", "producer wait empty", "issue busy (both issuers)", "mma total",
         "operands later than issue", "-", "mma wait tmem_empty", "wait full (both issuers)"]
print(which, "slices", ns, "ms", a.elapsed_time(b), "pair" if os.environ.get("FCEST_OZAKI_PAIR") == "1" else "")
c = c[c[:, 3] > 0]   # CTAs that issued MMAs (pair kernel: leaders only)
for i, n in enumerate(names):
    if n != "-":
        print(f"  {n:28s} mean {c[:, i].mean():12.0f}  min {c[:, i].min():12.0f}  max {c[:, i].max():12.0f} cycles")
