"""Data-parallel plumbing for the training hot path: one process per GPU, NCCL over NVLink.

The reference has no distributed code (only commented-out nn.DataParallel, main.py:660); SURVEY section 8(e) maps
its batch-parallel intent to: shard the batch by sample, per-replica BatchNorm statistics (DataParallel
semantics, no SyncBN), one gradient all-reduce(mean) per step over the parameters that actually receive
gradients (the 8 never-used refinenet4.resConfUnit1 tensors keep grad=None so AdamW leaves them alone, as in
the reference), and one all-reduce of a handful of scalars for evaluation metrics.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


class GradientAllReducer:
    """Flat-bucket gradient averaging.  `zero()` drops the gradients (grad=None), so autograd *assigns* each fresh
    gradient instead of launching one accumulation kernel per parameter (~400 per step for the default model).
    `reduce()` gathers the gradients that exist into one flat fp32 buffer with a multi-tensor copy, all-reduces it
    once (96 MB for the default model: launch-latency bound on NVSwitch, so one bucket beats many) and re-points
    `.grad` at the buffer's views for the optimizer.  With a single rank the gradients are left where autograd put
    them.  Parameters that never receive a gradient keep grad=None (AdamW then skips them, as in the reference)."""

    def __init__(self, params, world=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.flat = None
        self.live = None
        self.views = None

    def _build(self):
        keep = getattr(self, "_live_ids", set())
        self.live = [p for p in self.params if p.grad is not None or id(p) in keep]
        self._live_ids = {id(p) for p in self.live}
        if not self.live:
            self.flat, self.views = None, []
            return
        n = sum(p.numel() for p in self.live)
        dev = self.live[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.live:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def zero(self):
        for p in self.params:
            p.grad = None

    def reduce(self):
        """call after backward(); averages gradients over ranks (no-op for world 1)."""
        if self.live is None:
            self._build()
        else:
            # a parameter that starts receiving gradients later (unfrozen layer, conditional branch) must join the
            # bucket, or the ranks diverge silently
            late = [p for p in self.params if p.grad is not None and id(p) not in self._live_ids]
            if late:
                if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("GradientAllReducer: the set of parameters with gradients changed inside a "
                                       "captured step; re-capture the graph")
                self._build()
        if self.world == 1 or not self.live:
            return
        grads, views = [], []
        for p, v in zip(self.live, self.views):
            if p.grad is None:
                v.zero_()
            else:
                grads.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                views.append(v)
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / self.world)
        for p, v in zip(self.live, self.views):
            p.grad = v

    def live_parameters(self):
        return self.live


def all_reduce_metric_sums(values, device=None):
    """sum a small list of python/torch scalars over ranks in fp64 (evaluation partials: per-rank sums of per-sample
    SI-RMSE, AbsRel, delta_k and the sample count; evaluation.py:157-176)."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def shard_range(n_items, rank, world):
    """contiguous shard [lo, hi) of n_items samples for this rank (evaluation / prediction passes)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
