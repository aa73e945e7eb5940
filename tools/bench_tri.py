#!/usr/bin/env python
"""Per-latent triangular kernels at the C4 shape (R = 2500 latents, M = 200): fcest_tri_syrk_pack and fcest_tri_grad,
device time with CUDA events (operands of 0.8 GB cycle through L2), against the fp64 peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcest_b200 import ops  # noqa: E402


def main():
    R, M = 2500, 200
    g = torch.Generator(device="cuda").manual_seed(0)
    K = M * (M + 1) // 2
    ldk = (K + 15) // 16 * 16
    q = torch.tril(torch.randn(R, M, M, dtype=torch.float64, device="cuda", generator=g)).contiguous()
    Gk = torch.randn(R, ldk, dtype=torch.float64, device="cuda", generator=g)
    Sk = torch.empty((R, ldk), dtype=torch.float64, device="cuda")
    klp = torch.empty((R, 2), dtype=torch.float64, device="cuda")
    dq = torch.empty_like(q)

    def timeit(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    t_s = timeit(lambda: ops.call("fcest_tri_syrk_pack", ops.ptr(q), R, M, ops.ptr(Sk), ldk, ops.ptr(klp)))
    t_g = timeit(lambda: ops.call("fcest_tri_grad", ops.ptr(Gk), ldk, ops.ptr(q), R, M, ops.ptr(dq), 1.0))
    # reference values for a change check: checksums of the outputs
    out = {"tri_syrk_ms": t_s, "tri_syrk_tflops": R * M ** 3 / 3 / t_s * 1e-9, "tri_grad_ms": t_g,
           "tri_grad_tflops": 2 * R * M ** 3 / 3 / t_g * 1e-9,
           "checksum_Sk": float(Sk[:, :K].double().sum()), "checksum_dq": float(dq.sum()), "absmax_dq": float(dq.abs().max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
