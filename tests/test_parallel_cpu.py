"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: subject partition, ELBO all-reduce,
SubjectBatch stepping.  The per-subject compute is a stand-in (the CUDA path needs a GPU); what is
checked is that every subject is owned exactly once and that the reduced ELBO is the global sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fcest_b200.parallel import SubjectBatch, subject_partition


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
