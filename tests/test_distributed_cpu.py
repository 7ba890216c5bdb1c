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
