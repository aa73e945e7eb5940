"""Multi-GPU plumbing: one process per GPU (torchrun).  Two ways the path shards:

* independent subjects per rank (``SubjectBatch``) -- no data-path collective, ELBO sum only;
* one subject, time points sharded over the ranks (``TimeShardedSVWP``): every rank evaluates the
  likelihood + projection on its slice of the N x S batch and the ELBO and gradients are combined with
  SUM all-reduces (NCCL over NVLink on GPUs, gloo in the CPU tests).
Kernels always see a rank-local slice; the collectives are issued from the host between kernel launches.
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


def time_partition(num_time_steps: int, world: int, rank: int) -> slice:
    """Contiguous block of time points owned by ``rank`` (sizes differ by at most one)."""
    r = subject_partition(num_time_steps, world, rank)
    return slice(r.start, r.stop)


class TimeShardedSVWP:
    """Data-parallel evaluation of ONE sparse Wishart-process model over the time axis.

    Every rank holds the full parameter set and the full data (x (N,1), y (N,D): < 1 MB); a step evaluates
    the conditional, the likelihood and their adjoints on this rank's block of time points, then
      * the packed q_sqrt cotangent Gk = g_var^T Pk (R x M(M+1)/2) is SUM all-reduced inside the backward
        pass (before the per-latent ``tri_grad``), so ``q_sqrt.grad`` is complete on every rank;
      * ELBO and the remaining gradients (q_mu, Z, kernel, A, lambda) travel in one flat SUM all-reduce.
    The KL term carries weight 1 / world on every rank so that the sums count it once.  Results equal the
    single-GPU evaluation up to the order of the floating-point sums over time points.
    Only the sparse model shards this way (VGP couples all time points through chol(K(X, X))).
    """

    def __init__(self, model, world: int | None = None, rank: int | None = None, reduce_=None):
        if not getattr(model, "_sparse", False):
            raise NotImplementedError("time sharding needs the sparse (SVWP) model")
        r, w, _ = dist_env()
        self.model = model
        self.world = w if world is None else world
        self.rank = r if rank is None else rank
        self.reduce_ = allreduce_sum_ if reduce_ is None else reduce_
        lik = model.likelihood  # decorrelate the ranks' Monte-Carlo draws (each rank draws its own time block)
        lik.seed(lik._seed + 7919 * (self.rank + 1))

    def small_grad_parameters(self) -> list:
        m = self.model
        ps = [m.q_mu, m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix, m.likelihood.additive_part]
        if m.inducing_variable.Z.trainable:
            ps.append(m.inducing_variable.Z)
        return ps

    def elbo_and_grad(self, data, epsilon=None, check: bool = True) -> torch.Tensor:
        x, y = data
        sl = time_partition(len(x), self.world, self.rank)
        eps = None if epsilon is None else epsilon[:, sl]
        m = self.model
        gk_reduce = self.reduce_ if self.world > 1 else None
        elbo = m.elbo_and_grad((x[sl], y[sl]), epsilon=eps, check=check, kl_scale=1.0 / self.world,
                               gk_reduce=gk_reduce)
        if self.world > 1:
            ps = self.small_grad_parameters()
            flat = torch.cat([elbo.reshape(1)] + [p.grad.reshape(-1) for p in ps])
            self.reduce_(flat)
            elbo = flat[0].clone()
            o = 1
            for p in ps:
                n = p.grad.numel()
                p.grad = flat[o:o + n].reshape(p.grad.shape).clone()
                o += n
        return elbo
