#!/usr/bin/env python
"""Config C5 slice: independent SVWP subjects with D = 200 (N=1200, S=8, M=100 inducing points -- M and S are our
stated choices, BASELINE.json does not fix them), `--per-gpu` subjects per rank, sharded over the ranks with
fcest_b200.parallel.SubjectBatch and an NCCL all-reduce of the summed ELBO.  Reports ms per batch step (max
over ranks) and subject evaluations per second; 1000 subjects = 125 such steps' worth of subjects per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
      tools/bench_c5.py --per-gpu 2 --steps 3
"""
import argparse, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=2)
    ap.add_argument("--subjects", type=int, default=0, help="total number of subjects (overrides --per-gpu)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--D", type=int, default=200)
    ap.add_argument("--M", type=int, default=100)
    a = ap.parse_args()
    from fcest_b200.models import SparseVariationalWishartProcess
    from fcest_b200.parallel import SubjectBatch, dist_env
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, D, S, M = 1200, a.D, 8, a.M

    def make(sid):
        rng = np.random.default_rng(sid)  # per-subject seed
        x = np.linspace(0, 1, N)[:, None]
        y = rng.standard_normal((N, D))
        m = SparseVariationalWishartProcess(D, np.linspace(0, 1, M)[:, None], num_mc_samples=S, verbose=False, device=dev)
        m.likelihood.A_scale_matrix.assign(np.eye(D) + 0.05 * rng.standard_normal((D, D)))
        m.likelihood.seed(1000 + sid)
        return m, (torch.tensor(x, device=dev), torch.tensor(y, device=dev))

    n_subjects = a.subjects if a.subjects > 0 else a.per_gpu * world
    import time
    t_build = time.perf_counter()
    batch = SubjectBatch(n_subjects, make)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    # warm-up: one evaluation of the first local subject (module loads, allocator, workspaces), not a full sweep
    batch.models[0].elbo_and_grad(batch.data[0], check=False)
    for prm in batch.models[0].parameters:
        prm.grad = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        total = batch.step(dev, keep_grads=False)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({"config": f"C5: {n_subjects} independent SVWP subjects over {world} GPU(s), N=1200 D={D} S=8 M={M} "
                                    "(M and S are our stated choices), ELBO + gradient of every subject per sweep, "
                                    "ELBO sum all-reduced (NCCL)", "n_gpus": world,
                          "subjects": n_subjects, "subjects_per_gpu_max": len(batch.models), "sweeps_timed": a.steps,
                          "ms_per_sweep": ms, "ms_per_subject_evaluation": ms / max(len(batch.models), 1),
                          "subject_evals_per_s": n_subjects / (ms * 1e-3), "elbo_sum": float(total.item()),
                          "model_build_s": t_build,
                          "mem_gb_per_gpu": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
