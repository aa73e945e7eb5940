#!/usr/bin/env python
"""One C4 subject (SVWP N=1200 D=50 S=8 M=200) sharded over the ranks -- by time points
(fcest_b200.parallel.TimeShardedSVWP, --mode time) or by latent GP for the conditional and by time for the
likelihood (LatentShardedSVWP, --mode latent): parity of the sharded ELBO + gradients against the unsharded
evaluation on the same noise, then strong-scaling timing (device events, max over ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_timeshard.py --steps 5
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--N", type=int, default=1200)
    ap.add_argument("--D", type=int, default=50)
    ap.add_argument("--S", type=int, default=8)
    ap.add_argument("--M", type=int, default=200)
    ap.add_argument("--mode", choices=["time", "latent"], default="time")
    args = ap.parse_args()
    from fcest_b200.models import SparseVariationalWishartProcess
    from fcest_b200.parallel import LatentShardedSVWP, TimeShardedSVWP, dist_env
    from oracle import wishart_oracle as wo  # input generator only

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, D, S, M = args.N, args.D, args.S, args.M
    x, y, _, p = wo.make_inputs(N, D, 1, M=M, seed=0)
    m = SparseVariationalWishartProcess(D, p["Z"].numpy(), num_mc_samples=S, verbose=False, device=dev)
    m.q_mu.assign(p["q_mu"]); m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"]); m.likelihood.additive_part.assign(p["lam"])
    xd, yd = x.to(dev), y.to(dev)
    g = torch.Generator(device=dev).manual_seed(123)  # the same noise on every rank
    eps = torch.randn(S, N, D * D, dtype=torch.float64, device=dev, generator=g)

    names = ["q_mu", "q_sqrt", "kvar", "kls", "A", "lam", "Z"]
    params = [m.q_mu, m.q_sqrt, m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix,
              m.likelihood.additive_part, m.inducing_variable.Z]
    ref_elbo = m.elbo_and_grad((xd, yd), epsilon=eps).item()
    ref = [q.grad.clone() for q in params]
    sh = TimeShardedSVWP(m) if args.mode == "time" else LatentShardedSVWP(m)
    elbo = sh.elbo_and_grad((xd, yd), epsilon=eps).item()
    errs = {"elbo": abs(elbo - ref_elbo) / abs(ref_elbo)}
    if args.mode == "latent":   # q_mu / q_sqrt gradients stay sharded: compare the own block
        lat = sh.latents
        errs["q_mu"] = float((sh.q_mu_grad_local - ref[0][:, lat]).abs().max() / ref[0].abs().max())
        errs["q_sqrt"] = float((sh.q_sqrt_grad_local - ref[1][lat]).abs().max() / ref[1].abs().max())
    for nm, q, r in list(zip(names, params, ref))[2 if args.mode == "latent" else 0:]:
        errs[nm] = float((q.grad - r).abs().max() / r.abs().max())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sh.elbo_and_grad((xd, yd), check=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sh.elbo_and_grad((xd, yd), check=False)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    if rank == 0:
        label = ("time-sharded" if args.mode == "time" else "latent-sharded conditional + time-sharded likelihood")
        print(json.dumps({"mode": label + " single subject (strong scaling)", "n_gpus": world, "N": N, "D": D, "S": S,
                          "M": M, "ms_per_step": ms, "evals_per_s": 1e3 / ms, "max_rel_err_vs_unsharded": errs,
                          "collective_bytes_per_step": (int(8 * (D * D * ((M * (M + 1) // 2 + 15) // 16 * 16))) if args.mode == "time"
                                                        else int(2 * 2 * 8 * N * D * D))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
