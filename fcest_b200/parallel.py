"""Multi-GPU plumbing: one process per GPU (torchrun).  Two ways the path shards:

* independent subjects per rank (``SubjectBatch``) -- no data-path collective, ELBO sum only;
* one subject, time points sharded over the ranks (``TimeShardedSVWP``): every rank evaluates the
  likelihood + projection on its slice of the N x S batch and the ELBO and gradients are combined with
  SUM all-reduces (NCCL over NVLink on GPUs, gloo in the CPU tests);
* one subject, latent GPs sharded for the conditional and time points sharded for the likelihood
  (``LatentShardedSVWP``, SURVEY.md 8e option B): q_sqrt, its gradient and the three projection GEMMs split by
  latent, the latent moments and their cotangents change layout in one all-to-all each way (48 MB in total
  at C4), and only the small gradients are all-reduced.
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

    def step(self, device=None, keep_grads: bool = True) -> torch.Tensor:
        """``keep_grads=False`` drops every subject's gradients once its optimiser (if any) has consumed them: with
        many resident subjects the dense q_sqrt gradients (8 M^2 R bytes each) would otherwise double the footprint."""
        total = None
        for i, (m, d) in enumerate(zip(self.models, self.data)):
            e = m.elbo_and_grad(d, check=False).reshape(1)
            total = e.clone() if total is None else total + e
            if self.optimizers is not None:
                self.optimizers[i].step()
            if not keep_grads:
                for prm in m.parameters:
                    prm.grad = None
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


def alltoall_chunks(send: list, recv: list) -> None:
    """``recv[q] <- send[rank]`` of rank q, for contiguous tensors of per-pair sizes (in place).

    NCCL: one ``all_to_all`` (NVLink / NVSwitch).  Other backends (gloo in the CPU tests has no all-to-all):
    batched point-to-point sends and receives."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        recv[0].copy_(send[0])
        return
    if dist.get_backend() == "nccl":
        dist.all_to_all(recv, send)
        return
    rank = dist.get_rank()
    recv[rank].copy_(send[rank])
    reqs = []
    for q in range(dist.get_world_size()):
        if q != rank:
            reqs.append(dist.P2POp(dist.isend, send[q], q))
            reqs.append(dist.P2POp(dist.irecv, recv[q], q))
    for w in dist.batch_isend_irecv(reqs):
        w.wait()


class LatentShardedSVWP:
    """ONE sparse Wishart-process model over several GPUs: conditional sharded by latent GP, likelihood by time.

    The SVGP projection is independent per latent r and the likelihood per time point n (SURVEY.md 8e,
    option B).  Rank p owns the latents ``latents`` (a contiguous block of the R = D nu columns of q_mu / rows
    of q_sqrt) and the time points ``times``.  One evaluation:

      1. conditional on the own latents, all N time points: Sk_p, fvar_p = Pk Sk_p^T, fmean_p  (N, R_p);
      2. all-to-all  (N, R_p) -> (N_p, R): every rank receives the moments of all latents on its time block;
      3. Wishart likelihood + analytic cotangents on the own time block                        (N_p, R);
      4. all-to-all back (N_p, R) -> (N, R_p);
      5. backward of the conditional on the own latents: d q_sqrt[latents], d q_mu[:, latents] are complete
         and stay local (with their Adam state, if any); d P and what follows from it are partial sums;
      6. one flat SUM all-reduce of the ELBO and the small gradients (kernel, Z, A, lambda).

    No gradient of the size of q_sqrt ever travels (TimeShardedSVWP all-reduces 402 MB at C4) and the per-latent
    kernels (tri_syrk, tri_grad) shard too.  The single-matrix chain (Kmm, Cholesky, P, Pk) is replicated.
    The phases are separate methods so that a single process can step several emulated ranks in lock-step
    (tests); ``elbo_and_grad`` chains them with the real collectives.
    """

    def __init__(self, model, world: int | None = None, rank: int | None = None):
        if not getattr(model, "_sparse", False):
            raise NotImplementedError("latent sharding needs the sparse (SVWP) model")
        r, w, _ = dist_env()
        self.model = model
        self.world = w if world is None else world
        self.rank = r if rank is None else rank
        self.R = model.q_mu.unconstrained.shape[1]
        self.latent_blocks = [subject_partition(self.R, self.world, q) for q in range(self.world)]
        lat = self.latent_blocks[self.rank]
        self.latents = slice(lat.start, lat.stop)
        self.q_mu_grad_local = None      # (M, R_p)
        self.q_sqrt_grad_local = None    # (R_p, M, M)
        lik = model.likelihood
        lik.seed(lik._seed + 7919 * (self.rank + 1))
        self._st = None

    # ---- phases ----------------------------------------------------------------------------------
    def small_grad_parameters(self) -> list:
        m = self.model
        ps = [m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix, m.likelihood.additive_part]
        if m.inducing_variable.Z.trainable:
            ps.append(m.inducing_variable.Z)
        return ps

    def forward_local(self, x) -> list:
        """Phase 1: moments of the own latents at all time points -> one (N_q, 2, R_p) chunk per rank q."""
        from .gp import conditionals as cond
        m = self.model
        fmean, fvar, kl, st = cond.conditional_forward(
            m.kernel, m.inducing_variable.Z.unconstrained, x, m.q_mu.unconstrained[:, self.latents],
            m.q_sqrt.unconstrained[self.latents], vgp=False, want_kl=True)
        self._st, self._kl = st, kl
        n = x.shape[0]
        self.time_blocks = [time_partition(n, self.world, q) for q in range(self.world)]
        both = torch.stack([fmean, fvar], dim=1)                              # (N, 2, R_p)
        return [both[sl].contiguous() for sl in self.time_blocks]

    def recv_buffers_forward(self, dev) -> list:
        sl = self.time_blocks[self.rank]
        return [torch.empty((sl.stop - sl.start, 2, len(b)), dtype=torch.float64, device=dev)
                for b in self.latent_blocks]

    def likelihood_local(self, x, y, recv: list, epsilon=None) -> list:
        """Phase 3: likelihood on the own time block with all latents -> one (N_p, 2, R_q) cotangent chunk per
        rank q."""
        from . import ops
        m = self.model
        sl = self.time_blocks[self.rank]
        fmean = ops.as_pitched(torch.cat([c[:, 0, :] for c in recv], dim=1))  # (N_p, R)
        fvar = ops.as_pitched(torch.cat([c[:, 1, :] for c in recv], dim=1))
        eps = None if epsilon is None else epsilon[:, sl]
        out = m.likelihood.variational_expectations(x[sl], fmean, fvar, y[sl].contiguous(), epsilon=eps,
                                                    need_grad=True)
        self._lik = out
        g = torch.stack([out["g_fmean"], out["g_fvar"]], dim=1)                # (N_p, 2, R)
        return [g[:, :, b.start:b.stop].contiguous() for b in self.latent_blocks]

    def recv_buffers_backward(self, dev) -> list:
        rp = len(self.latent_blocks[self.rank])
        return [torch.empty((sl.stop - sl.start, 2, rp), dtype=torch.float64, device=dev)
                for sl in self.time_blocks]

    def backward_local(self, recv: list) -> torch.Tensor:
        """Phase 5: backward of the conditional on the own latents; returns the flat buffer of partial sums
        [ELBO, small gradients] for the all-reduce."""
        from . import ops
        from .gp import conditionals as cond
        m = self.model
        g_fmean = ops.as_pitched(torch.cat([c[:, 0, :] for c in recv], dim=0))  # (N, R_p)
        g_fvar = ops.as_pitched(torch.cat([c[:, 1, :] for c in recv], dim=0))
        need_Z = m.inducing_variable.Z.trainable
        gr = cond.conditional_backward(self._st, g_fmean, g_fvar, need_Z=need_Z)
        self.q_mu_grad_local, self.q_sqrt_grad_local = gr["q_mu"], gr["q_sqrt"]
        out = self._lik
        elbo = torch.empty(1, dtype=torch.float64, device=g_fmean.device)
        ops.sum_into(out["ve"], out["ve"].numel(), 1, 1.0, elbo, False)
        ops.axpby(-1.0, self._kl, 1.0, elbo)
        parts = [elbo, gr["kvar"].reshape(-1), gr["kls"].reshape(-1), out["gA"].reshape(-1), out["glam"].reshape(-1)]
        if need_Z:
            parts.append(gr["Z"].reshape(-1))
        self._infos = (out["info"], self._st.info)
        return torch.cat(parts)

    def finish(self, flat: torch.Tensor, check: bool = True) -> torch.Tensor:
        """Phase 6 (after the all-reduce of ``flat``): scatter the small gradients into ``Parameter.grad``."""
        from . import ops
        m = self.model
        lik = m.likelihood
        kv, kls = m.kernel.variance, m.kernel.lengthscales
        g_kvar, g_kls = flat[1:2].clone(), flat[2:3].clone()
        kv.grad = ops.softplus_bwd(kv.unconstrained.reshape(1), g_kvar, torch.empty_like(g_kvar)).reshape(kv.shape)
        kls.grad = ops.softplus_bwd(kls.unconstrained.reshape(1), g_kls, torch.empty_like(g_kls)).reshape(kls.shape)
        o = 3
        for prm in [lik.A_scale_matrix, lik.additive_part] + ([m.inducing_variable.Z] if m.inducing_variable.Z.trainable else []):
            n = prm.unconstrained.numel()
            prm.grad = flat[o:o + n].reshape(prm.unconstrained.shape).clone()
            o += n
        if check:
            lik.check_info(self._infos[0])
            m._check_chol(self._infos[1])
        return flat[0].clone()

    # ---- one evaluation with the real collectives ------------------------------------------------------
    def elbo_and_grad(self, data, epsilon=None, check: bool = True) -> torch.Tensor:
        """ELBO (identical on every rank); ``q_mu_grad_local`` / ``q_sqrt_grad_local`` hold the gradients of the
        own latents, the small parameters' ``.grad`` are complete on every rank."""
        x, y = self.model._data(data)
        send = self.forward_local(x)
        recv = self.recv_buffers_forward(x.device)
        alltoall_chunks(send, recv)
        send = self.likelihood_local(x, y, recv, epsilon=epsilon)
        recv = self.recv_buffers_backward(x.device)
        alltoall_chunks(send, recv)
        flat = self.backward_local(recv)
        allreduce_sum_(flat)
        return self.finish(flat, check=check)

    # ---- training with sharded optimiser state ------------------------------------------------------------
    def adam_step(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999,
                  epsilon: float = 1e-7, update_small: bool = True) -> None:
        """One Keras-flavoured Adam update (``helpers.inference.KerasAdam``) after ``elbo_and_grad``: the own block of
        q_sqrt / q_mu with LOCAL first / second moments (1 / world of the 2.4 GB of Adam state at C4), the small
        parameters replicated (their gradients are identical on every rank after the all-reduce, so the replicas
        stay bit-identical).  ``update_small=False`` is for emulated ranks that share one model object."""
        from . import ops
        from .helpers.inference import KerasAdam
        m = self.model
        if getattr(self, "_adam_t", None) is None:
            self._adam_t = 0
            self._qs_local = m.q_sqrt.unconstrained[self.latents]              # contiguous view: updated in place
            self._qm_local = m.q_mu.unconstrained[:, self.latents].contiguous()  # master copy of the own columns
            self._mom = [torch.zeros_like(t) for t in (self._qs_local, self._qs_local, self._qm_local, self._qm_local)]
            self._small_opt = KerasAdam([p for p in self.small_grad_parameters() if p.trainable],
                                        learning_rate, beta_1, beta_2, epsilon)
        self._adam_t += 1
        kw = dict(grad_sign=-1.0, lr=learning_rate, b1=beta_1, b2=beta_2, eps=epsilon)
        if m.q_sqrt.trainable:
            ops.adam_step(self._qs_local, self.q_sqrt_grad_local.contiguous(), self._mom[0], self._mom[1],
                          self._adam_t, **kw)
        if m.q_mu.trainable:
            ops.adam_step(self._qm_local, self.q_mu_grad_local.contiguous(), self._mom[2], self._mom[3],
                          self._adam_t, **kw)
            m.q_mu.unconstrained[:, self.latents] = self._qm_local
        if update_small:
            self._small_opt.step()

    def gather_parameters(self) -> None:
        """All-gather the trained q_mu / q_sqrt blocks so that every rank holds the full model (prediction,
        ``save_model_params_dict``)."""
        m = self.model
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            return
        M = m.q_mu.unconstrained.shape[0]
        dev = m.q_sqrt.unconstrained.device
        qs = [torch.empty((len(b), M, M), dtype=torch.float64, device=dev) for b in self.latent_blocks]
        dist.all_gather(qs, m.q_sqrt.unconstrained[self.latents].contiguous())
        qm = [torch.empty((len(b), M), dtype=torch.float64, device=dev) for b in self.latent_blocks]
        dist.all_gather(qm, m.q_mu.unconstrained[:, self.latents].t().contiguous())
        m.q_sqrt.unconstrained.copy_(torch.cat(qs, dim=0))
        m.q_mu.unconstrained.copy_(torch.cat(qm, dim=0).t())

    def gather_grads(self) -> None:
        """All-gather the sharded gradients into ``model.q_mu.grad`` / ``model.q_sqrt.grad`` (compatibility /
        tests; training keeps them sharded)."""
        m = self.model
        if self.world == 1 or not (dist.is_available() and dist.is_initialized()):
            m.q_mu.grad, m.q_sqrt.grad = self.q_mu_grad_local, self.q_sqrt_grad_local
            return
        M = m.q_mu.unconstrained.shape[0]
        dev = self.q_sqrt_grad_local.device
        qs = [torch.empty((len(b), M, M), dtype=torch.float64, device=dev) for b in self.latent_blocks]
        dist.all_gather(qs, self.q_sqrt_grad_local.contiguous())
        qm = [torch.empty((len(b), M), dtype=torch.float64, device=dev) for b in self.latent_blocks]
        dist.all_gather(qm, self.q_mu_grad_local.t().contiguous())
        m.q_sqrt.grad = torch.cat(qs, dim=0)
        m.q_mu.grad = torch.cat(qm, dim=0).t().contiguous()
