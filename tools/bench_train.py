#!/usr/bin/env python
"""Training-loop throughput (ELBO + gradients + Keras-Adam update per iteration) on the small reference
configs, eager launches vs. the CUDA-graph step of fcest_b200.helpers.inference.GraphedTrainingStep.
  C1: VWP  N=200  D=3  S=5 (README dummy data)      C2: SVWP N=1200 D=15 M=200 S=3
  C4: SVWP N=1200 D=50 M=200 S=8 -- the graph step with and without the Adam update of q_sqrt fused into tri_grad"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fcest_b200.helpers.inference import GraphedTrainingStep, KerasAdam
from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess


def build(cfg):
    rng = np.random.default_rng(0)
    if cfg == "C1":
        N, D = 200, 3
        x = np.linspace(0, 1, N)[:, None]; y = rng.standard_normal((N, D))
        return VariationalWishartProcess(x, y, num_mc_samples=5), None
    N, D, M, S = (1200, 15, 200, 3) if cfg == "C2" else (1200, 50, 200, 8)
    x = np.linspace(0, 1, N)[:, None]; y = rng.standard_normal((N, D))
    m = SparseVariationalWishartProcess(D, np.linspace(0, 1, M)[:, None], num_mc_samples=S, verbose=False)
    return m, (x, y)


def run(cfg, graphed, iters=200, fuse=True):
    m, data = build(cfg)
    opt = KerasAdam(m.trainable_parameters)
    if graphed:
        step = GraphedTrainingStep(m, data, opt, fuse_q_sqrt=fuse)
    else:
        xd_yd = m._data(data)
        def step():
            e = m.elbo_and_grad(xd_yd, check=False)
            opt.step()
            return e
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        e = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return iters / dt, float(e.item())


def run_eval(cfg, graphed, iters=200):
    """ELBO + gradient evaluations per second (no optimiser): eager launches vs model.compile_evaluation."""
    m, data = build(cfg)
    if graphed:
        f = m.compile_evaluation(data)
        step = lambda: f()
    else:
        xd_yd = m._data(data)
        step = lambda: m.elbo_and_grad(xd_yd, check=False)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return iters / (time.perf_counter() - t0)


res = {}
for cfg in ("C1", "C2"):
    eager, _ = run(cfg, False)
    graph, _ = run(cfg, True)
    res[cfg] = {"eager_it_per_s": eager, "cuda_graph_it_per_s": graph, "speedup": graph / eager,
                "eval_eager_per_s": run_eval(cfg, False), "eval_cuda_graph_per_s": run_eval(cfg, True)}
unfused, _ = run("C4", True, iters=40, fuse=False)
fused, _ = run("C4", True, iters=40, fuse=True)
res["C4"] = {"cuda_graph_it_per_s_unfused_adam": unfused, "cuda_graph_it_per_s_fused_q_sqrt_adam": fused,
             "ms_per_iteration": [1e3 / unfused, 1e3 / fused], "speedup": fused / unfused}
print(json.dumps(res))
