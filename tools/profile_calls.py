#!/usr/bin/env python
"""Per-entry-point device time of one SVWP ELBO + gradient evaluation at a given shape (CUDA events around
every C-ABI call).   python tools/profile_calls.py --D 200 --M 100 --S 8 --N 1200"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(N=1200, D=50, S=8, M=200).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from fcest_b200 import _lib
    from fcest_b200.models import SparseVariationalWishartProcess
    from oracle import wishart_oracle as wo  # input generator only
    dev = torch.device("cuda")
    x, y, _, p = wo.make_inputs(a.N, a.D, 1, M=a.M, seed=0)
    m = SparseVariationalWishartProcess(a.D, p["Z"].numpy(), num_mc_samples=a.S, verbose=False, device=dev)
    m.q_mu.assign(p["q_mu"]); m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"]); m.likelihood.additive_part.assign(p["lam"])
    xd, yd = x.to(dev), y.to(dev)
    for _ in range(2):
        m.elbo_and_grad((xd, yd), check=False)
    torch.cuda.synchronize()
    _lib.PROFILE = {"min_flops": -1.0, "records": []}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.elbo_and_grad((xd, yd), check=False); e1.record()
    torch.cuda.synchronize()
    per = {}
    for name, fl, s, e in _lib.PROFILE["records"]:
        key = name if not (name == "fcest_dgemm" and fl >= 1e10) else "fcest_dgemm(projection)"
        per[key] = per.get(key, 0.0) + s.elapsed_time(e)
    _lib.PROFILE = None
    print(json.dumps({"shape": vars(a), "step_ms": e0.elapsed_time(e1),
                      "per_call_ms": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])},
                      "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
