"""window_cov only (C3 shape): a few launches with an L2 flush in between; used under ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops
N, D, W = 1200, 100, 30
y = torch.tensor(np.random.default_rng(0).standard_normal((N, D)), device="cuda")
out = torch.empty((N, D, D), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.call("fcest_window_cov", ops.ptr(y), N, D, W, 1, 0, ops.ptr(out)); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("window_cov ms", ts)
