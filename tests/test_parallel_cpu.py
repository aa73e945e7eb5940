"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: subject partition, ELBO all-reduce,
SubjectBatch stepping.  The per-subject compute is a stand-in (the CUDA path needs a GPU); what is
checked is that every subject is owned exactly once and that the reduced ELBO is the global sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcest_b200.parallel import (SubjectBatch, TimeShardedSVWP, alltoall_chunks, subject_partition,
                                 time_partition)


def test_partition_covers_every_subject_once():
    for n in [0, 1, 7, 8, 1000]:
        for world in [1, 2, 3, 8]:
            ids = [i for r in range(world) for i in subject_partition(n, world, r)]
            assert ids == list(range(n))
            sizes = [len(subject_partition(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    def __init__(self, sid):
        self.sid = sid
        self.steps = 0

    def elbo_and_grad(self, data, check=True):
        self.steps += 1
        return torch.tensor(-100.0 * (self.sid + 1) - data, dtype=torch.float64)


def _worker(rank, world, port, num_subjects, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = SubjectBatch(num_subjects, lambda sid: (_FakeModel(sid), float(sid)))
    total = batch.step()
    total2 = batch.step()
    out.put((rank, list(batch.subject_ids), float(total), float(total2), [m.steps for m in batch.models]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_subject_batch_allreduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    expected = sum(-100.0 * (i + 1) - float(i) for i in range(n))
    assert res[0][1] + res[1][1] == list(range(n))
    for _, ids, t1, t2, steps in res:
        assert abs(t1 - expected) < 1e-12 and abs(t2 - expected) < 1e-12
        assert steps == [2] * len(ids)


def test_time_partition_covers_every_time_step_once():
    for n in [1, 7, 1200]:
        for world in [1, 2, 3, 8]:
            cover = [i for r in range(world) for i in range(n)[time_partition(n, world, r)]]
            assert cover == list(range(n))


class _P:
    def __init__(self, shape, trainable=True):
        self.grad = None
        self.trainable = trainable
        self.shape = shape


class _FakeSparseModel:
    """Stand-in with the surface TimeShardedSVWP touches; 'ELBO' = sum(y_slice) - kl_scale * 10, gradients are
    per-slice sums so that the all-reduced values are known in closed form."""
    _sparse = True

    def __init__(self):
        ns = type("ns", (), {})
        self.q_mu, self.q_sqrt = _P((2, 3)), _P((3, 2, 2))
        self.kernel = ns(); self.kernel.variance, self.kernel.lengthscales = _P(()), _P(())
        self.likelihood = ns(); self.likelihood.A_scale_matrix, self.likelihood.additive_part = _P((2, 2)), _P((2,))
        self.likelihood._seed = 0; self.likelihood.seed = lambda s: None
        self.inducing_variable = ns(); self.inducing_variable.Z = _P((2, 1))

    def elbo_and_grad(self, data, epsilon=None, check=True, kl_scale=1.0, gk_reduce=None):
        x, y = data
        t = y.sum()
        gk = torch.full((3, 3), float(t))
        if gk_reduce is not None:
            gk_reduce(gk)
        self.q_sqrt.grad = gk.clone()
        for q in [self.q_mu, self.kernel.variance, self.kernel.lengthscales, self.likelihood.A_scale_matrix,
                  self.likelihood.additive_part, self.inducing_variable.Z]:
            q.grad = torch.full(q.shape, float(t) - kl_scale, dtype=torch.float64)
        return t - kl_scale * 10.0


def _ts_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y = torch.arange(14, dtype=torch.float64).reshape(7, 2)
    x = torch.zeros(7, 1, dtype=torch.float64)
    m = _FakeSparseModel()
    sh = TimeShardedSVWP(m)
    elbo = sh.elbo_and_grad((x, y))
    out.put((rank, float(elbo), float(m.q_sqrt.grad[0, 0]), float(m.q_mu.grad[0, 0]), float(m.inducing_variable.Z.grad[1, 0])))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_time_sharded_allreduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ts_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    total = float(sum(range(14)))
    for _, elbo, gq, gmu, gz in res:
        assert abs(elbo - (total - 10.0)) < 1e-12      # KL counted once
        assert abs(gq - total) < 1e-12                 # q_sqrt cotangent reduced inside the backward pass
        assert abs(gmu - (total - 1.0)) < 1e-12 and abs(gz - (total - 1.0)) < 1e-12


def _a2a_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the layout change of LatentShardedSVWP: rank p holds (N, 2, R_p), rank q must end up with (N_q, 2, R)
    N, R = 7, 5
    full = torch.arange(N * 2 * R, dtype=torch.float64).reshape(N, 2, R)
    lat = [subject_partition(R, world, q) for q in range(world)]
    tim = [time_partition(N, world, q) for q in range(world)]
    mine = full[:, :, lat[rank].start:lat[rank].stop]
    send = [mine[tim[q]].contiguous() for q in range(world)]
    nq = tim[rank].stop - tim[rank].start
    recv = [torch.empty((nq, 2, len(lat[p])), dtype=torch.float64) for p in range(world)]
    alltoall_chunks(send, recv)
    got = torch.cat(recv, dim=2)
    ok_fwd = bool(torch.equal(got, full[tim[rank]]))
    # and back: (N_p, 2, R) -> (N, 2, R_p)
    send = [got[:, :, lat[q].start:lat[q].stop].contiguous() for q in range(world)]
    recv = [torch.empty((tim[q].stop - tim[q].start, 2, len(lat[rank])), dtype=torch.float64) for q in range(world)]
    alltoall_chunks(send, recv)
    ok_bwd = bool(torch.equal(torch.cat(recv, dim=0), mine))
    out.put((rank, ok_fwd, ok_bwd))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_latent_time_alltoall_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
