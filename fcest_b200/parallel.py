"""Multi-GPU plumbing: one process per GPU (torchrun), independent subjects per rank.

The path shards over *subjects* (SURVEY.md 8e (1)): every subject is its own model with its own
parameters and optimiser state, so there is no data-path collective; ranks only combine the per-step
ELBO sum (fp64 SUM all-reduce; NCCL over NVLink on GPUs, gloo in the CPU tests).  Kernels always see a
rank-local slice.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str | None = None, device=None) -> tuple:
    """Initialise torch.distributed from the environment if WORLD_SIZE > 1.  Returns dist_env()."""
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def subject_partition(num_subjects: int, world: int, rank: int) -> range:
    """Contiguous block partition of subject ids; sizes differ by at most one, every id owned once."""
    base, extra = divmod(num_subjects, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


class SubjectBatch:
    """Owns this rank's share of a batch of independent subjects and steps them together.

    ``make_model(subject_id) -> (model, data)`` builds one subject's model on this rank's device;
    ``step()`` runs one ELBO + gradient evaluation (and optionally one Adam update) for every local
    subject and returns the ELBO summed over *all* subjects of *all* ranks.
    """

    def __init__(self, num_subjects: int, make_model, world: int | None = None, rank: int | None = None,
                 optimizer_factory=None):
        r, w, _ = dist_env()
        self.world = w if world is None else world
        self.rank = r if rank is None else rank
        self.subject_ids = list(subject_partition(num_subjects, self.world, self.rank))
        self.models, self.data = [], []
        for sid in self.subject_ids:
            m, d = make_model(sid)
            self.models.append(m)
            self.data.append(d)
        self.optimizers = [optimizer_factory(m) for m in self.models] if optimizer_factory else None

    def step(self, device=None) -> torch.Tensor:
        total = None
        for i, (m, d) in enumerate(zip(self.models, self.data)):
            e = m.elbo_and_grad(d, check=False).reshape(1)
            total = e.clone() if total is None else total + e
            if self.optimizers is not None:
                self.optimizers[i].step()
        if total is None:
            total = torch.zeros(1, dtype=torch.float64, device=device or "cpu")
        return allreduce_sum_(total)
