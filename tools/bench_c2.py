#!/usr/bin/env python
"""C2 (SVWP N=1200 D=15 M=200 S=3): ELBO + gradient evaluations/s through the CUDA-graph evaluation; A/B tool for the
int8-GEMM dispatch thresholds (FCEST_OZAKI_MIN_DIM / FCEST_OZAKI_MIN_K / FCEST_OZAKI_MIN_FLOPS)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops  # noqa: E402
from fcest_b200.helpers.synthetic import make_synthetic  # noqa: E402
from fcest_b200.models import SparseVariationalWishartProcess  # noqa: E402

dev = torch.device("cuda")
n_, d_, s_, m_ = 1200, 15, 3, 200
q = make_synthetic(n_, d_, s_, M=m_, seed=0, want_eps=False)
mm = SparseVariationalWishartProcess(d_, q["Z"], num_mc_samples=s_, verbose=False, device=dev)
mm.q_mu.assign(q["q_mu"]); mm.q_sqrt.assign(q["q_sqrt"])
mm.likelihood.A_scale_matrix.assign(q["A"]); mm.likelihood.additive_part.assign(q["lam"])
f = mm.compile_evaluation((q["x"], q["y"]))
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
print(json.dumps({"min_dim": ops.OZAKI_MIN_DIM, "min_k": ops.OZAKI_MIN_K, "evals_per_s": 100.0 / (a.elapsed_time(b) * 1e-3),
                  "ms_per_eval": a.elapsed_time(b) / 100.0}))
