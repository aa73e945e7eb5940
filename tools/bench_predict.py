#!/usr/bin/env python
"""predict_cov / predict_corr (300 MC samples) at the C4 model size: device time per call (events), the
streaming-moments kernel's share, and achieved fp64 rate (S* N* D^2 nu FMA = 2 flops each)."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(N=1200, D=50, M=200, S=300).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from fcest_b200 import _lib
    from fcest_b200.models import SparseVariationalWishartProcess
    from oracle import wishart_oracle as wo  # input generator only
    dev = torch.device("cuda")
    x, y, _, p = wo.make_inputs(a.N, a.D, 1, M=a.M, seed=0)
    m = SparseVariationalWishartProcess(a.D, p["Z"].numpy(), verbose=False, device=dev)
    m.q_mu.assign(p["q_mu"]); m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"]); m.likelihood.additive_part.assign(p["lam"])
    xd = x.to(dev)
    res = {"shape": vars(a)}
    for name, fn in (("predict_corr", m.predict_corr), ("predict_cov", m.predict_cov)):
        fn(xd, a.S); torch.cuda.synchronize()
        _lib.PROFILE = {"min_flops": -1.0, "records": []}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(xd, a.S); e1.record(); torch.cuda.synchronize()
        per = {}
        for nm, fl, s, e in _lib.PROFILE["records"]:
            per[nm] = per.get(nm, 0.0) + s.elapsed_time(e)
        _lib.PROFILE = None
        flops = 2.0 * a.S * a.N * a.D * a.D * a.D
        mom = per.get("fcest_cov_moments", 0.0)
        res[name] = {"ms": e0.elapsed_time(e1), "cov_moments_ms": mom, "randn_ms": per.get("fcest_randn", 0.0),
                     "cov_moments_tflops": flops / (mom * 1e-3) * 1e-12 if mom else None,
                     "top": {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:5]}}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
