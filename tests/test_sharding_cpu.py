"""world_size-2 gloo tests (CPU) of the date-sharded step's HOST logic: the partition of dates, the single
all-reduce over [gradient | loss], and the B_local/B_global weighting for unequal shards.  The CUDA calls
are replaced by a stub that writes known per-rank gradients (the kernels themselves are covered by -m gpu)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from factorvae_b200.batched import DateShardedStep, shard_dates


def test_shard_dates_partition():
    for B in (1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_dates(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeLayout:
    total = 37


def _worker(rank, world, port, B_locals, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from factorvae_b200 import batched, engine

    B_local = B_locals[rank]

    def fake_forward(layout, flat, x, y, date_ptr, *, loss_out=None, workspace=None, **kw):
        loss_out.fill_(10.0 + rank)                       # local mean loss of this rank's dates
        class St:                                          # minimal StepState
            pass
        st = St(); st.workspace = torch.zeros(1)
        return {"loss": loss_out}, st

    def fake_backward(layout, st, grad=None):
        grad.copy_(torch.arange(layout.total, dtype=torch.float32) * (rank + 1))   # local mean gradient
        return grad

    engine.elbo_forward, engine.elbo_backward = fake_forward, fake_backward
    flat = torch.zeros(_FakeLayout.total)
    stepper = DateShardedStep(_FakeLayout(), flat, precision="fp32")
    date_ptr = torch.arange(B_local + 1, dtype=torch.int32)
    stepper.step(torch.zeros(B_local, 1, 1), torch.zeros(B_local), date_ptr, global_dates=sum(B_locals),
                 eps=torch.zeros(B_local))
    q.put((rank, stepper.grad.clone(), float(stepper.loss)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B_locals", [(4, 4), (5, 3)])
def test_single_allreduce_gives_global_mean(B_locals):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B_locals, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Bg = sum(B_locals)
    want_grad = sum(torch.arange(37, dtype=torch.float32) * (r + 1) * B_locals[r] / Bg for r in range(2))
    want_loss = sum((10.0 + r) * B_locals[r] / Bg for r in range(2))
    for rank, grad, loss in res:
        assert torch.allclose(grad, want_grad, rtol=1e-6, atol=1e-6)
        assert abs(loss - want_loss) < 1e-5


def _worker_acc(rank, world, port, micro, q):
    """micro[rank] = the date counts of this rank's micro-batches; the stub kernels return, per call, the mean gradient
    (rank + 1) * (index of the micro-batch + 1) * arange and the mean loss 10 + rank + index."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from factorvae_b200 import engine

    calls = {"n": 0, "steps": set()}

    def fake_forward(layout, flat, x, y, date_ptr, *, loss_out=None, workspace=None, philox=None, **kw):
        loss_out.fill_(10.0 + rank + calls["n"])
        calls["steps"].add(philox[1])                      # every micro-batch of a step draws with the SAME step index
        class St:
            pass
        st = St(); st.workspace = torch.zeros(1)
        return {"loss": loss_out}, st

    def fake_backward(layout, st, grad=None):
        grad.copy_(torch.arange(layout.total, dtype=torch.float32) * (rank + 1) * (calls["n"] + 1))
        calls["n"] += 1
        return grad

    engine.elbo_forward, engine.elbo_backward = fake_forward, fake_backward
    stepper = DateShardedStep(_FakeLayout(), torch.zeros(_FakeLayout.total), precision="fp32")
    Bg = sum(sum(m) for m in micro)
    mbs, base = [], 0
    for b in micro[rank]:
        mbs.append((torch.zeros(b, 1, 1), torch.zeros(b), torch.arange(b + 1, dtype=torch.int32), base))
        base += b
    stepper.step_accumulate(mbs, global_dates=Bg)
    q.put((rank, stepper.grad.clone(), float(stepper.loss), sorted(calls["steps"]), stepper.step_index))
    dist.destroy_process_group()


@pytest.mark.parametrize("micro", [((2, 2), (2, 2)), ((3, 1, 2), (5,))])
def test_micro_batch_accumulation_weights_and_single_exchange(micro):
    """step_accumulate: local accumulation with weights B_micro / B_global, ONE all-reduce, one step index per step."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_acc, args=(r, 2, port, micro, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Bg = sum(sum(m) for m in micro)
    want_grad = sum(torch.arange(37, dtype=torch.float32) * (r + 1) * (i + 1) * b / Bg for r in range(2) for i, b in enumerate(micro[r]))
    want_loss = sum((10.0 + r + i) * b / Bg for r in range(2) for i, b in enumerate(micro[r]))
    for rank, grad, loss, steps, idx in res:
        assert torch.allclose(grad, want_grad, rtol=1e-6, atol=1e-6)
        assert abs(loss - want_loss) < 1e-5
        assert steps == [1] and idx == 1
