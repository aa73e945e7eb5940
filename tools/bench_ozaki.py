#!/usr/bin/env python
"""Times the three C4 projection GEMM shapes: native DMMA (fcest_dgemm) vs the int8 tcgen05 emulation
(fcest_dgemm_ozaki) for several digit counts; prints error vs the DMMA result and effective fp64 TFLOP/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops  # noqa: E402


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    N, R, K = 1200, 2500, 20100
    g = torch.Generator(device="cuda").manual_seed(0)
    ldk = (K + 15) // 16 * 16
    Pk = torch.randn(N, ldk, dtype=torch.float64, device="cuda", generator=g)[:, :K]
    Sk = torch.randn(R, ldk, dtype=torch.float64, device="cuda", generator=g)[:, :K]
    gv = ops.as_pitched(torch.randn(N, R, dtype=torch.float64, device="cuda", generator=g))
    shapes = {"V = Pk Sk^T": (True, True, N, R, K, Pk, Sk), "Gk = g^T Pk": (False, False, R, K, N, gv, Pk),
              "Tk = g Sk": (True, False, N, K, R, gv, Sk)}
    out = {}
    for name, args in shapes.items():
        M_, N_, K_ = args[2], args[3], args[4]
        ref = torch.empty((M_, (N_ + 15) // 16 * 16), dtype=torch.float64, device="cuda")[:, :N_]
        got = torch.empty_like(ref)
        ops.GEMM_IMPL = "dmma"
        t_d = timeit(lambda: ops.dgemm(*args, ref))
        rec = {"dmma_ms": t_d[0], "dmma_tflops": 2.0 * M_ * N_ * K_ / t_d[0] * 1e-9}
        for ns in (5, 6, 7):
            t_o = timeit(lambda: ops.dgemm_ozaki(*args, got, slices=ns))
            rel = float((got - ref).abs().max() / ref.abs().max())
            rec[f"ozaki{ns}"] = {"ms_best": t_o[0], "ms_mean": t_o[1], "fp64_equiv_tflops": 2.0 * M_ * N_ * K_ / t_o[0] * 1e-9,
                                 "rel_to_max_vs_dmma": rel}
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    json.dump(out, open(os.path.join("gpurun_out", "r02_bench_ozaki.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
