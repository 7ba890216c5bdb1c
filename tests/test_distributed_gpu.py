"""Batch-global loss moments across ranks: with the batch sharded by sample over two ranks, combined_loss with
the SiLog and edge-aware terms enabled must equal the single-device loss of the whole batch (reference util.py:70, 118-121),
and each rank's gradient must be world x the matching slice of the single-device gradient (the gradient all-reduce then
averages over ranks).  Two processes share the one GPU; the two tiny collectives travel over gloo (host-staged), so no
kernel ever waits on another process."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

B_LOCAL, H, W = 2, 48, 64


def _inputs():
    g = torch.Generator().manual_seed(2024)
    t = torch.rand(2 * B_LOCAL, 1, H, W, generator=g) * 9.9 + 0.1
    t[0, :, :5, :7] = 0.0                                   # masked-out pixels (SiLog mask = target > 0), unevenly spread
    p = (t + 0.05) * torch.exp(0.2 * torch.randn(t.shape, generator=g)) * 1.2
    rgb = torch.rand(2 * B_LOCAL, 3, H, W, generator=g)
    rgb[B_LOCAL:] *= 3.0                                    # the global max of the RGB gradient lives on rank 1 only
    return p, t, rgb


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    import depth_b200
    from depth_b200 import config as cfgmod
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    p, t, rgb = _inputs()
    lo, hi = rank * B_LOCAL, (rank + 1) * B_LOCAL
    pl = p[lo:hi].cuda().requires_grad_(True)
    cfg = cfgmod.loss_config(si=1.0, silog=1.0, vf=0.85, grad=0.2, edge=0.5)
    total, parts = depth_b200.combined_loss(pl, t[lo:hi].cuda(), cfg, rgb=rgb[lo:hi].cuda())
    total.backward()
    q.put((rank, float(total), parts, pl.grad.cpu()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_single_device(pkg):
    import torch.multiprocessing as mp
    from depth_b200 import config as cfgmod
    p, t, rgb = _inputs()
    cfg = cfgmod.loss_config(si=1.0, silog=1.0, vf=0.85, grad=0.2, edge=0.5)
    pf = p.cuda().requires_grad_(True)
    total, parts = pkg.combined_loss(pf, t.cuda(), cfg, rgb=rgb.cuda())
    total.backward()
    gfull = pf.grad.cpu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, tot, pr_parts, g in res:
        # batch-global terms are identical on every rank and equal to the single-device value
        assert abs(pr_parts["silog_loss"] - parts["silog_loss"]) <= 1e-5 * abs(parts["silog_loss"]), (pr_parts, parts)
        want = 2.0 * gfull[rank * B_LOCAL:(rank + 1) * B_LOCAL]
        err = float((g - want).abs().max()) / float(want.abs().max())
        assert err <= 1e-4, (rank, err)
    # per-sample terms are local means: their rank average is the single-device value
    for k in ("si_loss", "grad_loss", "edge_loss"):
        avg = 0.5 * (res[0][2][k] + res[1][2][k])
        assert abs(avg - parts[k]) <= 1e-5 * abs(parts[k]), (k, avg, parts[k])
