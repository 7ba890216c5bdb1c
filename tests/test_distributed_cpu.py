"""world_size-2 gloo tests (CPU) of the data-parallel host logic: flat gradient averaging equals the
single-process full-batch gradient for a BN-free model, unused parameters keep grad=None, metric partial sums."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import depth_b200
    from depth_b200 import distributed as D
    r, l, w = D.init_from_env(backend="gloo")
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 1, 3, padding=1))
    unused = nn.Parameter(torch.ones(3))
    params = list(net.parameters()) + [unused]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 8, 8, generator=g)
    y = torch.randn(4, 1, 8, 8, generator=g)
    red = D.GradientAllReducer(params, world=w)
    lo, hi = D.shard_range(4, r, w)
    for step in range(2):
        red.zero()
        loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        loss.backward()
        red.reduce()
    grads = [p.grad.clone() for p in net.parameters()]
    sums = D.all_reduce_metric_sums([1.0 + r, 2.0, hi - lo])
    if r == 0:
        net.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        ok = all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(grads, net.parameters()))
        q.put((ok, unused.grad is None, sums, len(red.live_parameters())))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, unused_none, sums, nlive = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "averaged shard gradients must equal the full-batch gradient"
    assert unused_none, "parameters without a gradient must stay grad=None"
    assert sums == [3.0, 4.0, 4.0] and nlive == 4


def test_shard_range_covers_everything():
    from depth_b200 import distributed as D
    for n in (0, 1, 7, 650, 8192):
        for w in (1, 2, 4, 8):
            got = []
            for r in range(w):
                lo, hi = D.shard_range(n, r, w)
                got += list(range(lo, hi))
            assert got == list(range(n))


def _worker_buckets(rank, world, port, q):
    """bucketed path: the conv weights' gradients are handed to the reducer one by one (what ops._on_side does on the
    side stream), the rest reaches it through reduce(); tiny buckets so that several collectives are issued."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import depth_b200  # noqa: F401
    from depth_b200 import distributed as D
    r, l, w = D.init_from_env(backend="gloo")
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.ReLU(),
                        nn.Conv2d(8, 1, 3, padding=1))
    params = list(net.parameters())
    weights = [m.weight for m in net if isinstance(m, nn.Conv2d)]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 8, 8, generator=g)
    y = torch.randn(4, 1, 8, 8, generator=g)
    red = D.GradientAllReducer(params, world=w, bucket_mb=0.001)
    lo, hi = D.shard_range(4, r, w)
    launched = []
    for step in range(3):
        red.zero()
        loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean()
        grads = torch.autograd.grad(loss, params)
        for p, gr in zip(params, grads):                 # biases: as autograd would have left them
            if all(p is not wt for wt in weights):
                p.grad = gr
        for p, gr in reversed(list(zip(params, grads))):  # weights: completion order = backward order
            if any(p is wt for wt in weights):
                if p is weights[1]:
                    half = gr * 0.5                       # a weight used twice in forward contributes twice per step
                    if not red.side_grad(p, half):
                        p.grad = half if p.grad is None else p.grad + half
                    if not red.side_grad(p, half):
                        p.grad = p.grad + half
                    continue
                if not red.side_grad(p, gr):
                    p.grad = gr if p.grad is None else p.grad + gr
        launched.append(sum(getattr(red, "_launched", [])))
        red.reduce()
    got = [p.grad.clone() for p in params]
    if r == 0:
        net.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        ok = all(torch.allclose(a, p.grad, atol=1e-6) for a, p in zip(got, params))
        q.put((ok, launched, len(red._buckets)))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29811 + os.getpid() % 150
    procs = [ctx.Process(target=_worker_buckets, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, launched, nb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "bucketed gradient averaging must equal the full-batch gradient"
    assert nb >= 2, "the test is meant to exercise several buckets"
    assert launched[0] == 0 and launched[2] == nb, (launched, nb)     # recording step, then every bucket from side_grad
