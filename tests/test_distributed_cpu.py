"""CPU (gloo, world_size 2) tests of the data-parallel plumbing of the hot path: start states are
sharded, weights replicated, and ONE flat-bucket all-reduce of the actor / critic gradients runs
between backward and clip_grad_norm_ (rl_sandbox_b200/utils/optimizer.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_sandbox_b200.utils.optimizer import Optimizer, allreduce_grads_


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ELU(), torch.nn.Linear(16, 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(123)
    data = torch.randn(8, 8)           # the "global batch" of start states (same on every rank)
    shard = data[rank * 4:(rank + 1) * 4]
    model = _model()
    opt = Optimizer(model, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    for _ in range(3):
        loss = model(shard).pow(2).mean()     # mean over the local shard (equal shards => mean of means)
        opt.step(loss)
    out[rank] = [p.detach().clone() for p in model.parameters()]
    dist.destroy_process_group()


def test_sharded_update_equals_single_process_update():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # reference: one process, whole batch
    torch.manual_seed(123)
    data = torch.randn(8, 8)
    model = _model()
    opt = Optimizer(model, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    for _ in range(3):
        opt.step(model(data).pow(2).mean())
    for a, b, c in zip(out[0], out[1], model.parameters()):
        assert torch.equal(a, b), "ranks diverged: the all-reduce must run before clipping"
        torch.testing.assert_close(a, c.detach(), rtol=1e-5, atol=1e-6)


def test_allreduce_is_a_no_op_without_process_group():
    m = _model()
    m(torch.ones(2, 8)).sum().backward()
    before = [p.grad.clone() for p in m.parameters()]
    allreduce_grads_(m.parameters())
    assert all(torch.equal(a, p.grad) for a, p in zip(before, m.parameters()))


def _bucket_worker(rank, world, port, out):
    from rl_sandbox_b200.utils.optimizer import GradBucket
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    actor, critic = _model(), _model()
    opt_a = Optimizer(actor, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    opt_c = Optimizer(critic, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    bucket = GradBucket(list(actor.parameters()) + list(critic.parameters()))
    assert bucket.attached() and bucket.flat.numel() == sum(p.numel() for m in (actor, critic) for p in m.parameters())
    torch.manual_seed(123)
    data = torch.randn(8, 8)
    shard = data[rank * 4:(rank + 1) * 4]
    for step in range(3):
        # "the kernel wrote the gradients": autograd.grad results copied into the bucket's views, as rlsb_ac_update does
        ga = torch.autograd.grad(actor(shard).pow(2).mean(), list(actor.parameters()))
        gc = torch.autograd.grad((critic(shard) - 1).pow(2).mean(), list(critic.parameters()))
        for p, g in zip(list(actor.parameters()) + list(critic.parameters()), ga + gc):
            p.grad.copy_(g)
        if step == 1:   # somebody dropped the views (e.g. zero_grad(set_to_none=True)): all_reduce re-attaches them
            for p in actor.parameters():
                p.grad = p.grad.clone()
            assert not bucket.attached()
        assert bucket.all_reduce() is True and bucket.attached()
        opt_a.step_with_grads(reduced=True)
        opt_c.step_with_grads(reduced=True)
    out[rank] = [p.detach().clone() for m in (actor, critic) for p in m.parameters()]
    dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_single_process_update():
    """actor + critic gradients live in ONE flat buffer (GradBucket): a single all-reduce, then both optimizers clip and
    step — identical on every rank and equal to the whole-batch single-process update."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(123)
    data = torch.randn(8, 8)
    actor, critic = _model(), _model()
    opt_a = Optimizer(actor, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    opt_c = Optimizer(critic, lr=1e-2, eps=1e-5, weight_decay=1e-6, clip=0.05)
    for _ in range(3):
        opt_a.step(actor(data).pow(2).mean())
        opt_c.step((critic(data) - 1).pow(2).mean())
    ref = [p.detach() for m in (actor, critic) for p in m.parameters()]
    for a, b, c in zip(out[0], out[1], ref):
        assert torch.equal(a, b), "ranks diverged"
        torch.testing.assert_close(a, c, rtol=1e-5, atol=1e-6)


def test_bucket_without_process_group_is_local():
    from rl_sandbox_b200.utils.optimizer import GradBucket
    m = _model()
    b = GradBucket(m.parameters())
    assert b.all_reduce() is False and b.attached()
    m(torch.ones(2, 8)).sum().backward()          # autograd accumulates into the views in place
    assert b.attached() and b.flat.abs().sum() > 0
