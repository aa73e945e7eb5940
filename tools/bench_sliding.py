"""Sliding-window kernels at config C3 (N=1200, D=100, w=30 + CV search): device time, achieved HBM
bandwidth of window_cov (algorithmic bytes = 8 N D^2 written + 8 N D read) against MEASURED_PEAKS.json,
and the CPU restatement of the reference timed beside it.  Writes gpurun_out/sliding_bench.json."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcest_b200 import ops
from fcest_b200.models import SlidingWindows
from oracle import sliding_oracle as so

N, D, W = 1200, 100, 30
rng = np.random.default_rng(0)
y = rng.standard_normal((N, D))
yd = torch.tensor(y, device="cuda")
out = torch.empty((N, D, D), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()  # evict L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


res = {"config": "C3: N=1200 D=100 window_length=30 step=1"}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
for corr in (0, 1):
    med, best = timed(lambda: ops.call("fcest_window_cov", ops.ptr(yd), N, D, W, 1, corr, ops.ptr(out)))
    bytes_alg = 8.0 * N * D * D + 8.0 * N * D
    res["window_cov_corr" if corr else "window_cov"] = {
        "ms_median": med, "ms_best": best, "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / med * 1e-6,
        "peak_gbs": peaks["hbm_gbs"], "frac": bytes_alg / med * 1e-6 / peaks["hbm_gbs"]}
# end-to-end through the Python API (host numpy in / out)
sw = SlidingWindows(np.linspace(0, 1, N)[:, None], y)
est = sw.overlapping_windowed_cov_estimation(W)  # first call allocates the pinned staging buffer
t0 = time.perf_counter(); est = sw.overlapping_windowed_cov_estimation(W); t1 = time.perf_counter()
res["window_cov_e2e_ms"] = 1e3 * (t1 - t0)
t0 = time.perf_counter(); ref = so.overlapping_windowed_cov(y, W); t1 = time.perf_counter()
res["window_cov_cpu_ms"] = 1e3 * (t1 - t0)
res["window_cov_max_abs_diff"] = float(np.abs(est - ref).max())
# CV search with the reference defaults (80 window lengths x 1019 locations at N=1200)
table = None
def cv():
    global table
    table = sw.find_cross_validated_optimal_window_length()
cv(); torch.cuda.synchronize()
t0 = time.perf_counter(); cv(); torch.cuda.synchronize(); t1 = time.perf_counter()
res["cv_search_ms"] = 1e3 * (t1 - t0)
res["cv_problems"] = int(table.shape[0] * table.shape[1])
res["cv_best_window"] = int(sw.get_optimal_window_length(table))
# CPU reference on a bounded sample of the same grid, extrapolated per problem
windows, locs = so.cv_grid(N)
sub_w, sub_t = windows[::10], locs[::100]
t0 = time.perf_counter(); _, _, tref = so.cross_validated_table(y, windows=sub_w, locs=sub_t); t1 = time.perf_counter()
res["cv_cpu_ms_per_problem"] = 1e3 * (t1 - t0) / (len(sub_w) * len(sub_t))
res["cv_cpu_extrapolated_s"] = res["cv_cpu_ms_per_problem"] * res["cv_problems"] * 1e-3
got = table.to_numpy()[::100][:, ::10]
fin = np.isfinite(tref)
res["cv_inf_pattern_equal"] = bool(np.array_equal(np.isinf(got), np.isinf(tref)))
res["cv_max_rel_diff"] = float(np.max(np.abs(got[fin] - tref[fin]) / np.abs(tref[fin]))) if fin.any() else None
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sliding_bench.json"), "w"), indent=1)
