#!/usr/bin/env python
"""Headline benchmark: Wishart ELBO + gradient evaluations per second (BASELINE.json `metric`).

Workload (config.workload = "C4"): SparseVariationalWishartProcess, N=1200 time points, D=nu=50
(R=2500 latent GPs), S=8 MC samples, M=200 inducing points (M is our stated choice),
synthetic N(0,1) data, parameters at a perturbed point.  One "step" = one ELBO value + gradient
w.r.t. every trainable variable (q_mu, q_sqrt, Z, kernel variance / lengthscale, A, lambda), with a
fresh N(0,1) draw generated on the device (the reference draws inside the step too).

  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA)
  python bench.py --impl reference --steps K --warmup W  # CPU restatement of the reference path

Multi-GPU: one process per GPU (torchrun); each rank owns independent subjects (weak scaling, no
data-path collective); the per-step ELBO sum is combined with an NCCL all-reduce.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(N=1200, D=50, S=8, M=200)
WORKLOAD = "C4: SVWP N=1200 D=50 nu=50 S=8 M=200 ELBO+grad"   # identical string in both arms
METRIC = "Wishart ELBO+grad evals/s (N=1200,D=50,S=8)"
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "FP64_PEAK.json")
GEMM_KERNEL_NAME = "dgemm_dmma_kernel<120,128,40,32,32,3>"
LIK_KERNEL_NAME = "wishart_loglik_warp_kernel<7> (+ 2 colsum reductions)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the three projection GEMMs, from an ncu --set full
# capture (a constant from the named profile, NOT measured in the bench run)
TRAFFIC_BYTES_PER_LAUNCH = 777e6
TRAFFIC_SOURCE = "constant from profiles/r01_ncu_summary.md (ncu --set full of the round-1 build: 969 / 604 / 758 MB for V, Gk, Tk); not measured in this run"
# int8-emulated GEMM kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures of
# the final build (V: 495 + 361 MB, Gk: 148 + 376 MB, Tk: 297 + 177 MB; profiles/r02_ncu_ozaki.md) -- constants from those
# captures, NOT measured in the bench run.  Algorithmic bytes = digit arrays read once + the fp64 output (or the split-K
# partial tiles) written once: V 464 + 131 MB, Gk 139 + 402 MB, Tk 275 + 193 MB.
OZAKI_TRAFFIC_BYTES_PER_LAUNCH = (856e6 + 524e6 + 474e6) / 3.0
OZAKI_TRAFFIC_SOURCE = ("constant: mean of the three ncu --set full captures of the final build (V 856 MB, Gk 524 MB, Tk 474 MB per "
                        "launch; profiles/r02_ncu_ozaki.md); not measured in this run")
OZAKI_ALGORITHMIC_BYTES = (595e6 + 541e6 + 468e6) / 3.0
CPU_ARM_BUDGET_S = 200.0   # wall-clock budget of the --impl reference run (full-size evaluations only)


def algorithmic_flops(N, D, S, M):
    """projection 3 M^2 N R, mean 6 N M R, likelihood+grad (11/3) D^3 per (s, n)."""
    R = D * D
    return {"projection": 3.0 * M * M * N * R, "mean": 6.0 * N * M * R, "likelihood": 11.0 / 3.0 * D ** 3 * S * N}


def dist_env():
    from fcest_b200.parallel import dist_env as _env
    return _env()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's TF/GPflow path, timed on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """All the host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so the thread
    count is pinned explicitly instead of being inherited from the launcher."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_inputs():
    from oracle import wishart_oracle as wo
    return wo.make_inputs(CFG["N"], CFG["D"], CFG["S"], M=CFG["M"], seed=0)


def cpu_eval(inputs):
    """ONE full-size C4 evaluation (ELBO + gradient w.r.t. every trainable) of the CPU restatement; returns
    (seconds, elbo, grads).  Never a sub-sample of the time axis."""
    from oracle import wishart_oracle as wo
    x, y, eps, p = inputs
    t0 = time.perf_counter()
    elbo, g = wo.svgp_elbo_and_grad_chunked(p, x, y, eps, chunk=125)
    return time.perf_counter() - t0, elbo, g


CPU_SAMPLE = ("full C4 evaluations (N=1200, D=50, S=8, M=200; ELBO + gradient of every trainable) of the torch-CPU "
              "float64 restatement of the reference's TF/GPflow op sequence (oracle/wishart_oracle.py, LTA "
              "materialised 125 latents at a time); not TensorFlow")


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = host_threads()
    inputs = cpu_inputs()
    t_start = time.perf_counter()
    # warm-up: at least one full evaluation; more (up to --warmup) only while they fit a quarter of the budget
    warm_times = [cpu_eval(inputs)[0]]
    while len(warm_times) < args.warmup and (len(warm_times) + 1) * max(warm_times) < 0.25 * CPU_ARM_BUDGET_S:
        warm_times.append(cpu_eval(inputs)[0])
    # timed steps: full evaluations; the step COUNT shrinks if K of them would not fit the budget
    left = CPU_ARM_BUDGET_S - (time.perf_counter() - t_start)
    steps = int(max(1, min(args.steps, left // max(min(warm_times), 1e-3))))
    times = [cpu_eval(inputs)[0] for _ in range(steps)]
    t_step = float(np.mean(times))
    value = 1.0 / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "evals/s", "n_gpus": args.gpus, "steps": steps, "steps_requested": args.steps,
        "warmup": len(warm_times), "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU restatement of the reference path on the host cores (rank 0 only); not TensorFlow; "
                           "every timed step is a full-size evaluation, the step count is cut to fit "
                           f"{CPU_ARM_BUDGET_S:.0f} s of wall clock"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} timed + {len(warm_times)} warm-up " + CPU_SAMPLE},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def build_model(seed, device):
    from fcest_b200.helpers.synthetic import make_synthetic
    from fcest_b200.models import SparseVariationalWishartProcess
    p = make_synthetic(CFG["N"], CFG["D"], CFG["S"], M=CFG["M"], seed=seed, want_eps=False)
    m = SparseVariationalWishartProcess(CFG["D"], p["Z"], num_mc_samples=CFG["S"], verbose=False,
                                        device=device)
    m.q_mu.assign(p["q_mu"])
    m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"])
    m.likelihood.additive_part.assign(p["lam"])
    m.likelihood.seed(1234 + seed)
    return m, p["x"], p["y"]


def measure_likelihood_large(dev):
    """Likelihood kernels of the larger sizes (outside the timed region, rank 0 at N = 1): the shared-memory-tiled kernel at
    the C5 size D = 200 and the cluster-tiled kernel at D = 256, forward + backward, CUDA events around the C-ABI call."""
    from fcest_b200 import ops as _o
    out = []
    for D, N, S, kernel in ((200, 296, 8, "wishart_loglik_tiles_kernel<26> + wishart_g_fill_kernel (TMA panel ring, precomputed G)"),
                            (256, 296, 8, "wishart_loglik_cluster_kernel<34, 2> (tiles in the distributed shared memory of a CTA pair)")):
        R = D * D
        g = torch.Generator(device=dev).manual_seed(D)
        fm = _o.pitched(N, R, dev); fm.copy_(0.3 * torch.randn(N, R, dtype=torch.float64, device=dev, generator=g))
        fv = _o.pitched(N, R, dev); fv.copy_(0.05 + torch.rand(N, R, dtype=torch.float64, device=dev, generator=g))
        y = torch.randn(N, D, dtype=torch.float64, device=dev, generator=g)
        A = torch.eye(D, dtype=torch.float64, device=dev) + 0.2 * torch.randn(D, D, dtype=torch.float64, device=dev, generator=g)
        lam = torch.full((D,), -4.6, dtype=torch.float64, device=dev)
        eps = torch.randn(S, N, R, dtype=torch.float64, device=dev, generator=g)   # 0.76 / 1.24 GB: larger than L2
        for _ in range(2):
            o = _o.wishart_loglik(fm, fv, eps, y, A, lam)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); o = _o.wishart_loglik(fm, fv, eps, y, A, lam); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[1]
        fl = 11.0 / 3.0 * D ** 3 * S * N
        out.append({"D": D, "N": N, "S": S, "kernel": kernel, "ms": ms, "algorithmic_flops": fl,
                    "achieved": fl / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "cholesky_ok": bool(int(o["info"].max().item()) == 0)})
        del fm, fv, eps, o
        torch.cuda.empty_cache()
    return out


def measure_fp64_peak(dev):
    """cuBLAS DGEMM 8192^3 through torch.matmul (2 N^3 flop), best of 10 after 2 warm-up calls, CUDA events: the
    fp64 roofline denominator, measured in this run on this GPU (MEASURED_PEAKS.json carries no fp64 figure)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    best = float("inf")
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) * 1e-12


def measure_int8_peak(dev):
    """Issue-rate ceiling of the int8 tensor pipe measured in this run: back-to-back tcgen05.mma kind::i8 (128x128x32, 64
    cycles each) on every SM for ~0.7 ms, CUDA events, best of 5 (the power state of a real GEMM is not reproduced: zero
    operands).  MEASURED_PEAKS.json only has a bf16 figure; the int8 pipe is nominally twice as wide."""
    from fcest_b200 import _lib
    lib = _lib.load_library()
    best = 0.0
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops_issued = int(lib.fcest_int8_mma_peak(20000, _lib.stream(dev)))
        e1.record()
        torch.cuda.synchronize()
        if ops_issued > 0 and i >= 2:
            best = max(best, ops_issued / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
    return best


def strong_scaling_block(dev, rank, world, K, single_gpu_ms):
    """ONE C4 subject over all ranks (LatentShardedSVWP: conditional sharded by latent GP, likelihood by time, two
    all-to-alls + one small all-reduce per evaluation): parity against the unsharded evaluation on the same noise,
    then K timed evaluations (barrier + device events, max over ranks)."""
    import torch.distributed as dist
    from fcest_b200 import ops
    from fcest_b200.parallel import LatentShardedSVWP
    m, x_np, y_np = build_model(seed=0, device=dev)       # the same subject on every rank
    xd, yd = torch.tensor(x_np, device=dev), torch.tensor(y_np, device=dev)
    N, D, S = CFG["N"], CFG["D"], CFG["S"]
    eps = torch.empty((S, N, D * D), dtype=torch.float64, device=dev)
    ops.randn(eps, 4242, 0)                               # counter-based: identical on every rank
    params = {"q_mu": m.q_mu, "q_sqrt": m.q_sqrt, "kvar": m.kernel.variance, "kls": m.kernel.lengthscales,
              "A": m.likelihood.A_scale_matrix, "lam": m.likelihood.additive_part, "Z": m.inducing_variable.Z}
    ref_elbo = m.elbo_and_grad((xd, yd), epsilon=eps).item()
    ref = {k: q.grad.clone() for k, q in params.items()}
    sh = LatentShardedSVWP(m)
    elbo = sh.elbo_and_grad((xd, yd), epsilon=eps).item()
    lat = sh.latents
    errs = {"elbo": abs(elbo - ref_elbo) / abs(ref_elbo),
            "q_mu": float((sh.q_mu_grad_local - ref["q_mu"][:, lat]).abs().max() / ref["q_mu"].abs().max()),
            "q_sqrt": float((sh.q_sqrt_grad_local - ref["q_sqrt"][lat]).abs().max() / ref["q_sqrt"].abs().max())}
    for k in ("kvar", "kls", "A", "lam", "Z"):
        errs[k] = float((params[k].grad - ref[k]).abs().max() / ref[k].abs().max())
    e = torch.tensor([max(errs.values())], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    del ref
    for _ in range(5):
        sh.elbo_and_grad((xd, yd), check=False)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sh.elbo_and_grad((xd, yd), check=False)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    return {"what": "ONE C4 subject sharded over all GPUs (LatentShardedSVWP: conditional by latent GP, likelihood by "
                    "time block; strong scaling)", "n_gpus": world, "ms_per_step": ms, "evals_per_s": 1e3 / ms,
            "single_gpu_ms_per_step": single_gpu_ms, "speedup_vs_1gpu": single_gpu_ms / ms,
            "alltoall_bytes_per_step": int(2 * 2 * 8 * N * D * D), "allreduce_bytes_per_step": int(8 * (3 + D * D + D + CFG["M"])),
            "max_rel_err_vs_unsharded": float(e.item()), "rel_err_rank0": errs, "steps": K,
            "timing": "barrier + CUDA events on the launching stream, max over ranks"}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.

    NVML is initialised in the constructor -- call it BEFORE the warm-up: on a fresh box the first NVML attach
    (or the first ``nvidia-smi`` process) takes a few hundred milliseconds and stalls kernel submission while it
    runs, which once turned a 15 ms step into 39 ms when the sampler was created inside the timed region.
    ``start()`` / ``stop()`` bracket the timed region; a Python thread polls in-process NVML every 20 ms (cheap
    queries, no process spawn).  Fallback without pynvml: an ``nvidia-smi -lms`` process, likewise started before
    the warm-up, whose samples are cut to the timed region by timestamps."""
    REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        import threading
        self.samples, self._run, self._thread, self.proc, self.h = [], False, None, None, None
        self._threading = threading
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._poll()   # first query outside the timed region
            self.samples.clear()
        except Exception:
            self.h = None
            try:
                q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.proc = None

    def _poll(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        try:
            watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            watts = float("nan")
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.samples.append((mhz, watts, {k for k, b in bits.items() if mask & b}))

    def _loop(self):
        while self._run:
            try:
                self._poll()
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        self.t0 = time.time()
        if self.h is not None:
            self._run = True
            self._thread = self._threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self.t1 = time.time()
        if self.h is not None:
            self._run = False
            self._thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            reasons = set().union(*[s[2] for s in self.samples])
            return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "power_w_max": float(np.nanmax([s[1] for s in self.samples])),
                    "samples": len(self.samples), "source": "in-process NVML, 20 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        sm, mx, reasons, power = [], [], set(), []
        for ln in out.strip().splitlines():
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not (self.t0 - 0.05 <= ts <= self.t1 + 0.05):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(self.REASONS, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm), "source": "nvidia-smi -lms 50, timed region only"}


def run_cuda(args):
    import torch.distributed as dist
    from fcest_b200 import _lib
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)

    model, x_np, y_np = build_model(seed=rank, device=dev)
    x_dev = torch.tensor(x_np, device=dev)
    y_dev = torch.tensor(y_np, device=dev)
    x_pin = torch.tensor(x_np).pin_memory()
    y_pin = torch.tensor(y_np).pin_memory()
    elbo_sum = torch.zeros(1, dtype=torch.float64, device=dev)

    def step(data):
        e = model.elbo_and_grad(data, check=False)
        if world > 1:
            elbo_sum.copy_(e.reshape(1))
            dist.all_reduce(elbo_sum)  # ELBO over all subjects (NCCL, fp64 SUM)
        return e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None   # NVML attach happens here, outside the timed region
    # at least W steps, and at least 24 (~0.4 s of load at C4: clock ramp, allocator, lazy module loads); a fixed
    # count so that every rank issues the same number of collectives
    warm_steps = max(W, 24)
    for _ in range(warm_steps):
        step((x_dev, y_dev))
    barrier()

    # ---- timed region: K steps, inputs resident in HBM ------------------------------------------------
    flops = algorithmic_flops(**CFG)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    ev0.record()
    for _ in range(K):
        step((x_dev, y_dev))
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item())
    value = world * K / (t_ms * 1e-3)

    # ---- roofline pass: the same K steps with the two backward chains serialised on one stream, CUDA events
    #      around every projection GEMM launch (in the timed region above two of the three GEMMs share the GPU
    #      with each other on two streams, so a per-launch event pair would time both) ----------------------
    from fcest_b200.gp import conditionals as _cond
    _cond.OVERLAP_BACKWARD = False
    step((x_dev, y_dev))
    torch.cuda.synchronize()
    from fcest_b200 import ops as _ops
    use_ozaki = _ops.GEMM_IMPL == "ozaki"
    lib = _lib.load_library()
    if use_ozaki:
        lib.fcest_dgemm_ozaki_profile(1)   # event pair around the GEMM kernel proper inside every fcest_dgemm_ozaki call
    _lib.PROFILE = {"min_flops": 1e10, "records": []}
    for _ in range(K):
        step((x_dev, y_dev))
    torch.cuda.synchronize()
    records = _lib.PROFILE["records"]
    _lib.PROFILE = None
    _cond.OVERLAP_BACKWARD = True
    gemm_ms = [r[2].elapsed_time(r[3]) for r in records]     # whole call (digit extraction + GEMM kernel + split-K sum)
    gemm_fl = [r[1] for r in records]
    gemm_tflops = sum(gemm_fl) / (sum(gemm_ms) * 1e-3) * 1e-12 if gemm_ms else 0.0
    kern_ms = []
    if use_ozaki:
        import ctypes as _C
        buf = (_C.c_float * 512)()
        n = int(lib.fcest_dgemm_ozaki_profile_read(buf, 512))
        lib.fcest_dgemm_ozaki_profile(0)
        kern_ms = [float(buf[i]) for i in range(max(n, 0))]

    # ---- end-to-end: public API with HOST (pinned) inputs, ELBO read back every step --------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = 0.0
    for _ in range(K):
        last = float(step((x_pin, y_pin)).item())
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * K / (float(te.item()) * 1e-3)

    # ---- one profiled step (outside the timed regions, serialised chains): per-entry-point device time ---
    _cond.OVERLAP_BACKWARD = False
    _lib.PROFILE = {"min_flops": -1.0, "records": []}
    step((x_dev, y_dev))
    torch.cuda.synchronize()
    per = {}
    for name, fl, a, b in _lib.PROFILE["records"]:
        key = name if not (name == "fcest_dgemm" and fl >= 1e10) else "fcest_dgemm(projection)"
        per[key] = per.get(key, 0.0) + a.elapsed_time(b)
    _lib.PROFILE = None
    _cond.OVERLAP_BACKWARD = True
    info_ok = True
    try:
        model.elbo_and_grad((x_dev, y_dev), check=True)
    except RuntimeError:
        info_ok = False

    # ---- achieved HBM bandwidth of the sliding-window covariance at C3 (outside the timed
    #      regions; L2 flushed between launches, CUDA events on the launching stream) -------------------------
    wcov = None
    if rank == 0 and world == 1:
        from fcest_b200 import ops as _ops
        Nw, Dw, Ww = 1200, 100, 30
        yw = torch.tensor(np.random.default_rng(0).standard_normal((Nw, Dw)), device=dev)
        ow = torch.empty((Nw, Dw, Dw), dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for i in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _ops.call("fcest_window_cov", _ops.ptr(yw), Nw, Dw, Ww, 1, 0, _ops.ptr(ow)); b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = json.load(open(peaks_file))["hbm_gbs"] if os.path.exists(peaks_file) else 6543.0
        wbytes = 8.0 * Nw * Dw * Dw + 8.0 * Nw * Dw
        wms = float(np.median(ts))
        wcov = {"kernel": "window_cov_bulk_kernel (C3: N=1200 D=100 window_length=30)", "bound": "hbm", "us": wms * 1e3,
                "algorithmic_bytes": wbytes, "achieved": wbytes / wms * 1e-6, "peak": hbm, "unit": "GB/s",
                "frac": wbytes / wms * 1e-6 / hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"}
        del yw, ow, flush

    # ---- the small BASELINE configurations (launch-bound): ELBO + gradient evaluations/s through the CUDA-graph
    #      evaluation (model.compile_evaluation), outside the timed region, rank 0 at N = 1 only ------------------
    small = None
    if rank == 0 and world == 1:
        small = {}
        from fcest_b200.helpers.synthetic import make_synthetic
        from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
        for name, (n_, d_, s_, m_) in {"C1: VWP N=200 D=3 S=5": (200, 3, 5, None), "C2: SVWP N=1200 D=15 M=200 S=3": (1200, 15, 3, 200)}.items():
            q = make_synthetic(n_, d_, s_, M=m_, seed=0, want_eps=False)
            if m_ is None:
                mm = VariationalWishartProcess(q["x"], q["y"], num_mc_samples=s_, device=dev)
                data = None
            else:
                mm = SparseVariationalWishartProcess(d_, q["Z"], num_mc_samples=s_, verbose=False, device=dev)
                data = (q["x"], q["y"])
            mm.q_mu.assign(q["q_mu"]); mm.q_sqrt.assign(q["q_sqrt"])
            mm.likelihood.A_scale_matrix.assign(q["A"]); mm.likelihood.additive_part.assign(q["lam"])
            f = mm.compile_evaluation(data)
            for _ in range(5):
                f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100):
                f()
            b.record()
            torch.cuda.synchronize()
            f.check()
            small[name] = {"evals_per_s": 100.0 / (a.elapsed_time(b) * 1e-3), "ms_per_eval": a.elapsed_time(b) / 100.0,
                           "how": "100 CUDA-graph evaluations (one graph launch + one Philox launch each), CUDA events"}
            del mm, f

    # ---- fp64 roofline denominator measured in this run; strong-scaling leg (N > 1) ----------------------------
    peak_measured = measure_fp64_peak(dev) if rank == 0 else None
    int8_peak = measure_int8_peak(dev) if (rank == 0 and use_ozaki) else None
    del model
    torch.cuda.empty_cache()
    strong = strong_scaling_block(dev, rank, world, K, t_ms / K) if world > 1 else None
    lik_large = measure_likelihood_large(dev) if (rank == 0 and world == 1) else None

    if rank == 0:
        peak_file = json.load(open(FP64_PEAK_FILE))["fp64_tflops"] if os.path.exists(FP64_PEAK_FILE) else None
        peak = peak_measured
        lik_ms = per.get("fcest_wishart_loglik", 0.0)
        flops_gemm = flops["projection"] / 3.0
        if use_ozaki:
            ns, ns_b = _ops.OZAKI_SLICES, _ops.OZAKI_SLICES_BWD
            # per step: one forward GEMM with ns digits, two backward GEMMs with ns_b digits
            pairs = (ns * (ns + 1) // 2 + 2 * (ns_b * (ns_b + 1) // 2)) / 3.0
            k_ms = float(np.mean(kern_ms)) if kern_ms else float("nan")
            mb = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops") \
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
            achieved = pairs * flops_gemm / (k_ms * 1e-3) * 1e-12
            roofline = {
                "bound": "tensor",
                "kernel": f"ozaki_gemm_kernel<{ns}> / <{ns_b}> (tcgen05.mma kind::i8, TMEM accumulators, TMA-staged digit tiles): "
                          f"the 3 projection GEMMs per step, each an fp64 GEMM emulated by exact int8 digit-pair GEMMs "
                          f"({ns * (ns + 1) // 2} pairs forward, {ns_b * (ns_b + 1) // 2} pairs for the two gradient GEMMs)",
                "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s", "frac": achieved / int8_peak if int8_peak else None,
                "unit_note": "int8 tensor operations (multiply + add) per second, i.e. TOP/s; algorithmic = digit pairs x 2 N R K, "
                             "tile padding (1200 -> 1280, 2500 -> 2560, 20100 -> 20160) not counted",
                "algorithmic_int8_ops_per_launch": pairs * flops_gemm, "digit_pairs_mean": pairs,
                "digits": {"forward": ns, "backward": ns_b},
                "kernel_ms": k_ms, "launches_timed": len(kern_ms),
                "peak_source": "measured in THIS run: back-to-back tcgen05.mma kind::i8 128x128x32 on all SMs (fcest_int8_mma_peak, "
                               "64 cycles per instruction), CUDA events, best of 5; MEASURED_PEAKS.json has no int8 figure",
                "peak_alt_2x_measured_bf16": 2.0 * mb if mb else None,
                "traffic": OZAKI_TRAFFIC_BYTES_PER_LAUNCH, "traffic_unit": "bytes/launch", "traffic_source": OZAKI_TRAFFIC_SOURCE,
                "algorithmic_bytes_per_launch": OZAKI_ALGORITHMIC_BYTES,
                "fp64_equivalent": {
                    "what": "the fp64 GEMM each call replaces (2 N R K flop) over the duration of the WHOLE call (digit extraction "
                            "+ GEMM kernel + split-K sum), against the fp64 tensor peak measured in this run",
                    "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s", "ratio": gemm_tflops / peak,
                    "call_ms": float(np.mean(gemm_ms)) if gemm_ms else None,
                    "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul, best of 10, CUDA events, this run"},
                "how": "CUDA event pair recorded by the library around every GEMM kernel launch (fcest_dgemm_ozaki_profile), "
                       "in a separate pass of the same K steps with the two backward chains serialised on one stream"}
        else:
            roofline = {"bound": "tensor", "kernel": GEMM_KERNEL_NAME + " (3 projection GEMMs per step)",
                        "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tflops / peak,
                        "traffic": TRAFFIC_BYTES_PER_LAUNCH, "traffic_unit": "bytes/launch", "traffic_source": TRAFFIC_SOURCE,
                        "algorithmic_bytes_per_launch": 619e6,
                        "peak_source": "measured in THIS run on this GPU: cuBLAS DGEMM 8192^3 via torch.matmul, best of "
                                       "10, CUDA events (MEASURED_PEAKS.json has no fp64 figure)",
                        "algorithmic_flops_per_launch": flops_gemm, "launches_timed": len(gemm_ms),
                        "how": "CUDA events around every projection GEMM launch in a separate pass of the same K steps "
                               "with the two backward chains serialised on one stream"}
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": K, "warmup": warm_steps, "warmup_requested": args.warmup,
            "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "gemm": ("projection GEMMs: fp64 emulated on the int8 tensor cores with base-254 digits: %d for the forward "
                                "GEMM (47.9 bits, ~1e-12 of the largest entry), %d for the two gradient GEMMs (39.9 bits, "
                                "~3e-11); everything else native fp64" % (_ops.OZAKI_SLICES, _ops.OZAKI_SLICES_BWD))
                               if use_ozaki else "native fp64 (DMMA)",
                       "sharding": "one independent C4 subject per GPU (replicas; ELBO sum all-reduced)",
                       "l2": "per-step working set 2.8 GB (q_sqrt, Sk, Pk, Gk, Tk, eps) >> 126 MB L2, no flush needed",
                       "rng": "fresh Philox N(0,1) draw per step, generated on device inside the step"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(x_pin.numel() * 8 + y_pin.numel() * 8),
                    "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_step": {"note": "algorithmic fp64 flops of the whole evaluation over the step time, against the fp64 peak; "
                                      "above 1 when the projection GEMMs run on the int8 tensor pipe",
                              "algorithmic_flops": sum(flops.values()),
                              "achieved": sum(flops.values()) / (t_ms / K * 1e-3) * 1e-12, "peak": peak,
                              "unit": "TFLOP/s", "frac": sum(flops.values()) / (t_ms / K * 1e-3) * 1e-12 / peak},
            "roofline_likelihood": {"kernel": LIK_KERNEL_NAME, "ms": lik_ms,
                                    "achieved": (flops["likelihood"] / (lik_ms * 1e-3) * 1e-12) if lik_ms else None,
                                    "peak": peak, "unit": "TFLOP/s",
                                    "frac": (flops["likelihood"] / (lik_ms * 1e-3) * 1e-12 / peak) if lik_ms else None},
            "roofline_window_cov": wcov,
            "roofline_likelihood_large": ([dict(r, peak=peak, frac=r["achieved"] / peak) for r in lik_large]
                                          if lik_large else None),
            "per_call_ms": {k: round(v, 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
            "elbo_last": last, "cholesky_ok": info_ok,
        }
        if strong is not None:
            line["strong"] = strong
        if small is not None:
            line["small_configs"] = small
        if world == 1 and not args.no_cpu_baseline:
            # CPU baseline + parity at the metric's own configuration: the SAME epsilon goes through the CPU
            # restatement and through the CUDA path (fresh model at the same seed), ELBO and every gradient compared
            threads = host_threads()
            inputs = cpu_inputs()
            runs = [cpu_eval(inputs) for _ in range(3)]
            ts = [r[0] for r in runs[1:]]
            _, elbo_ref, g_ref = runs[-1]
            line["cpu_baseline"] = {
                "value": 1.0 / min(ts), "unit": "evals/s", "cores": threads, "kind": "port",
                "sample": "best of 2 after 1 warm-up: " + CPU_SAMPLE}
            m2, x_np2, y_np2 = build_model(seed=0, device=dev)
            eps_dev = inputs[2].to(dev)
            elbo_cuda = m2.elbo_and_grad((torch.tensor(x_np2, device=dev), torch.tensor(y_np2, device=dev)),
                                         epsilon=eps_dev).item()
            got = {"q_mu": m2.q_mu.grad, "q_sqrt": m2.q_sqrt.grad, "A": m2.likelihood.A_scale_matrix.grad,
                   "lam": m2.likelihood.additive_part.grad, "u_var": m2.kernel.variance.grad,
                   "u_ls": m2.kernel.lengthscales.grad, "Z": m2.inducing_variable.Z.grad}
            grel = {k: float((v.cpu().reshape(-1) - g_ref[k].reshape(-1)).abs().max() / g_ref[k].abs().max())
                    for k, v in got.items()}
            line["parity_c4"] = {"elbo_rel": abs(elbo_cuda - elbo_ref) / abs(elbo_ref), "grad_rel_max": max(grel.values()),
                                 "grad_rel": grel, "elbo_cuda": elbo_cuda, "elbo_cpu": elbo_ref,
                                 "against": "oracle/wishart_oracle.py (pinned to the reference's own likelihood source, "
                                            "tests/golden/wishart_reference.npz) on the same epsilon, full C4 size",
                                 "tolerance": {"elbo_rel": 1e-8, "grad_rel_max": 1e-7}}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def main():
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version
    # banner there when NCCL_DEBUG is set on the box) are sent to stderr; the JSON goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
