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
    """Bucketed gradient averaging, overlapped with backward.

    `zero()` drops the gradients (grad=None), so autograd *assigns* each fresh gradient instead of launching one
    accumulation kernel per parameter (~400 per step for the default model).

    Weight gradients of the convolutions are produced on the side stream (ops._on_side) while the data-gradient chain
    runs on the main stream.  The first (warm-up) step records the order in which they complete; from then on the live
    gradients live in one flat fp32 buffer cut into ~`bucket_mb` buckets in that order, every weight gradient is written
    straight into its slice on the side stream, and the moment a bucket's last gradient has landed its all-reduce is
    launched asynchronously from there - so the collectives of the early buckets run under the rest of backward instead
    of after it.  `reduce()` (after backward) adds the gradients autograd produced on the main stream (norm scales,
    biases: a few hundred KB) as the last bucket, waits for the outstanding collectives and re-points `.grad` at the
    buffer's views for the optimizer.  All of it is captured in the step's CUDA graph.
    Bucket size: the convolution kernels are persistent (one CTA per SM); a collective that runs under them takes SMs away
    and the displaced CTAs run as a second wave.  Measured at N = 2 (B200, NVLink): 25 MB buckets overlapped with
    backward 960.8 images/s, one bucket launched when the last weight gradient lands 964.9 images/s (N = 1: 487.8).  The
    default is therefore ONE bucket (bucket_mb = 1024; the 96 MB all-reduce costs < 1 ms and still overlaps the tail of
    backward); pass bucket_mb = 25 (or DP_BUCKET_MB) for the finer-grained overlap.  With a single rank the gradients
    are left where autograd put them.  Parameters that never receive a gradient keep grad=None (AdamW then skips them,
    as in the reference)."""

    def __init__(self, params, world=None, bucket_mb=1024.0):
        self.params = [p for p in params if p.requires_grad]
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.bucket_bytes = int(float(os.environ.get("DP_BUCKET_MB", bucket_mb)) * (1 << 20))
        self.flat = None
        self.live = None
        self.views = None
        self._order = []          # ids of side-stream gradients in completion order (recorded until the buffer is built)
        self._calls = {}          # id -> contributions per step (a weight used twice in forward gets two)
        self._byid = {id(p): p for p in self.params}
        self._works = []

    # ---- bucket construction -----------------------------------------------------------------------------------------
    def _build(self):
        keep = getattr(self, "_live_ids", set())
        live = [p for p in self.params if p.grad is not None or id(p) in keep]
        self._live_ids = {id(p) for p in live}
        if not live:
            self.live, self.flat, self.views = [], None, []
            return
        side = [self._byid[i] for i in dict.fromkeys(self._order) if i in self._live_ids]
        side_ids = {id(p) for p in side}
        rest = [p for p in live if id(p) not in side_ids]
        self.live = side + rest                      # completion order first: buckets are contiguous slices
        n = sum(p.numel() for p in self.live)
        dev = self.live[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, self._view_of = [], {}
        off = 0
        for p in self.live:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            self._view_of[id(p)] = v
            off += p.numel()
        # buckets over the side-stream part; everything else is one tail bucket reduced in reduce()
        self._buckets, self._bucket_of = [], {}
        lo = cur = 0
        members = []
        for p in side:
            members.append(id(p))
            cur += p.numel()
            if (cur - lo) * 4 >= self.bucket_bytes:
                self._buckets.append((lo, cur, tuple(members)))
                lo, members = cur, []
        if members:
            self._buckets.append((lo, cur, tuple(members)))
        for b, (_, _, mem) in enumerate(self._buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._tail = (cur, n)
        self._reset_counters()

    def _reset_counters(self):
        if self.flat is None or not hasattr(self, "_buckets"):
            return
        self._left = {i: self._calls.get(i, 1) for i in self._bucket_of}
        self._bucket_left = [len(mem) for _, _, mem in self._buckets]
        self._launched = [False] * len(self._buckets)
        self._works = []

    def _all_reduce(self, t, async_op):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)      # gloo (CPU tests): no AVG
        t.mul_(1.0 / self.world)
        return None if not async_op else w

    # ---- per step ----------------------------------------------------------------------------------------------------
    def zero(self):
        for p in self.params:
            p.grad = None
        self._reset_counters()

    def side_grad(self, p, dw):
        """called by ops._on_side on the side stream with a freshly computed weight gradient.  Returns True when the
        gradient was placed (the caller must then not touch p.grad)."""
        i = id(p)
        if self.world == 1:
            return False
        if self.flat is None:                        # recording step
            self._order.append(i)
            self._calls[i] = self._calls.get(i, 0) + 1
            return False
        if i not in self._bucket_of:
            return False
        v = self._view_of[i]
        if p.grad is None:
            v.copy_(dw)
            p.grad = v
        else:
            v.add_(dw)
        self._left[i] -= 1
        if self._left[i] == 0:
            b = self._bucket_of[i]
            self._bucket_left[b] -= 1
            if self._bucket_left[b] == 0 and not self._launched[b]:
                lo, hi, _ = self._buckets[b]
                self._launched[b] = True
                w = self._all_reduce(self.flat[lo:hi], async_op=True)     # issued from the side stream, under backward
                if w is not None:
                    self._works.append(w)
        return True

    def reduce(self):
        """call after backward() (and ops.side_join()); averages gradients over ranks (no-op for world 1)."""
        if self.live is None:
            self._build()
            first = True
        else:
            first = False
            # a parameter that starts receiving gradients later (unfrozen layer, conditional branch) must join the
            # buffer, or the ranks diverge silently
            late = [p for p in self.params if p.grad is not None and id(p) not in self._live_ids]
            if late:
                if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("GradientAllReducer: the set of parameters with gradients changed inside a "
                                       "captured step; re-capture the graph")
                self._order = []                     # completion order is stale: everything goes through the tail once
                self._build()
                first = True
        if self.world == 1 or not self.live:
            return
        if first:
            # buffer just built: this step's gradients are still where autograd / the side stream put them
            grads, views = [], []
            for p, v in zip(self.live, self.views):
                if p.grad is None:
                    v.zero_()
                else:
                    grads.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                    views.append(v)
            if grads:
                torch._foreach_copy_(views, grads)
            self._all_reduce(self.flat, async_op=False)
        else:
            # buckets whose gradients did not all arrive on the side stream this step (or all of them, when the side
            # stream is off): fill in what is missing and reduce them here
            lo_t, hi_t = self._tail
            pending_lo = None
            grads, views = [], []
            for b, (lo, hi, mem) in enumerate(self._buckets):
                if self._launched[b]:
                    continue
                for i in mem:
                    p, v = self._byid[i], self._view_of[i]
                    if p.grad is None:
                        v.zero_()
                    elif p.grad.data_ptr() != v.data_ptr():
                        grads.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                        views.append(v)
                pending_lo = lo if pending_lo is None else min(pending_lo, lo)
            side_ids = self._bucket_of
            for p, v in zip(self.live, self.views):
                if id(p) in side_ids:
                    continue
                if p.grad is None:
                    v.zero_()
                else:
                    grads.append(p.grad if p.grad.dtype == torch.float32 else p.grad.float())
                    views.append(v)
            if grads:
                torch._foreach_copy_(views, grads)
            if pending_lo is not None:
                # un-launched buckets are contiguous with the tail only if they are the last ones; reduce each range
                for b, (lo, hi, mem) in enumerate(self._buckets):
                    if not self._launched[b]:
                        self._all_reduce(self.flat[lo:hi], async_op=False)
            if hi_t > lo_t:
                self._all_reduce(self.flat[lo_t:hi_t], async_op=False)
            for w in self._works:
                w.wait()
            self._works = []
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
