"""fcest_filtfilt timing (CUDA events around the C-ABI call, L2 flushed between launches) next to the
reference's per-column scipy loop on the host.  C3 shape (N=1200, D=100) and a 1000-subject batch of it."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200.helpers import filtering as ff
from oracle import filter_oracle as fo

res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tag, N, C in [("C3 one subject (N=1200, D=100)", 1200, 100), ("1000 subjects x D=100 (N=1200)", 1200, 100_000)]:
    y = torch.tensor(np.random.default_rng(0).standard_normal((N, C)), device="cuda")
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ff.highpass_filter_data(y, 30, 1.0); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    yh = y[:, :100].cpu().numpy()
    t0 = time.perf_counter()
    b, a = fo.butter_highpass(1 / 30, 0.5)
    from scipy.signal import filtfilt
    ref = np.stack([filtfilt(b, a, col) for col in yh.T], axis=1)   # the reference's loop (filtering.py:36-43)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    res[tag] = {"gpu_ms_median": float(np.median(ts[1:])), "gpu_ms_all": ts, "bytes": 4 * 8 * N * C,
                "gbs": 4 * 8 * N * C / (np.median(ts[1:]) * 1e-3) / 1e9,
                "cpu_scipy_ms_per_100_columns": cpu_ms,
                "bit_exact_vs_scipy": bool(np.array_equal(out[:, :100].cpu().numpy(), ref))}
print(json.dumps(res, indent=1))
