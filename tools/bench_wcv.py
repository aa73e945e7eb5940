"""window_cv only (C3: N=1200, D=100, reference default grid 80 window lengths x 1019 locations): kernel time
(CUDA events) and algorithmic rate.  Flops per non-singular problem: (w-1) D^2 (Gram, lower) + D^3/3 + D^2."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops
N, D = 1200, 100
y = torch.tensor(np.random.default_rng(0).standard_normal((N, D)), device="cuda")
w_min, w_max, w_step = 21, 181, 2
ws = np.arange(w_min, w_max, w_step)
t_lo, t_hi = (w_max - 1) // 2, N - (w_max + 1) // 2
num_t, num_w = t_hi - t_lo, len(ws)
table = torch.empty((num_t, num_w), dtype=torch.float64, device="cuda")
ts = []
for i in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.call("fcest_window_cv", ops.ptr(y), N, D, w_min, w_step, num_w, t_lo, 1, num_t, 1e6 * 2.220446049250313e-16, ops.ptr(table), num_w)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
flops = sum(num_t * ((w - 1) * D * D + D ** 3 / 3 + D * D) for w in ws if w - 2 >= D)
ms = float(np.median(ts[1:]))
print(json.dumps({"problems": num_t * num_w, "nonsingular": int(num_t * sum(1 for w in ws if w - 2 >= D)), "ms": ms,
                  "algorithmic_gflop": flops * 1e-9, "tflops": flops / (ms * 1e-3) * 1e-12,
                  "finite": int(torch.isfinite(table).sum().item())}))
