"""cuBLAS DGEMM (torch.matmul f64) throughput: the measured fp64 roofline denominator."""
import json, torch, time
def t(m, n, k, reps=5):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * m * n * k / best * 1e-9, best
out = {}
for (m, n, k) in [(8192, 8192, 8192), (4096, 4096, 4096), (1200, 2500, 20100), (2500, 20100, 1200), (1200, 20100, 2500)]:
    tf, ms = t(m, n, k); out[f"{m}x{n}x{k}"] = {"tflops": tf, "ms": ms}; print(m, n, k, f"{tf:.2f} TFLOP/s {ms:.3f} ms", flush=True)
# sustained
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = a.clone()
torch.cuda.synchronize(); t0 = time.time(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    c = a @ b; n += 1
    if n % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["sustained_8192"] = {"tflops": 2.0 * 8192**3 * n / e0.elapsed_time(e1) * 1e-9}
print("sustained", out["sustained_8192"])
json.dump(out, open("gpurun_out/dgemm_cublas.json", "w"), indent=1)
