"""Whole-step CUDA graph for the training hot loop (reference main.py:125-156).

The reference's loop body - zero_grad, forward, combined_loss, backward, optimizer.step - launches a few thousand
kernels per step; on B200 the eager Python/autograd dispatch of those launches (about 130 ms) is slower than the
kernels themselves.  `GraphedTrainStep` captures one full step (our kernels, the PyTorch-run encoders, the NCCL
gradient all-reduce and the fused AdamW update) into a single CUDA graph and replays it: the host submits one
graph launch per step, copies the batch into static buffers and reads back eight floats.
"""
import os

import torch

from . import ops, util
from .network import frozen_cast
from . import distributed as D


class GraphedTrainStep:
    def __init__(self, model, optimizer, config, example_inputs, example_targets, use_rgb=True, world=None,
                 warmup=3, side_wgrad=True, overlap_h2d=True):
        assert example_inputs.is_cuda and example_targets.is_cuda
        self.model, self.opt, self.cfg, self.use_rgb = model, optimizer, config, use_rgb
        self.side_wgrad = side_wgrad
        self.x = example_inputs.clone()
        self.t = example_targets.clone()
        self.red = D.GradientAllReducer(model.parameters(), world)
        self.out = None
        self._copy = self._sx = self._st = self._consumed = None
        # Copy-stream overlap of the next batch's H2D with the current replay: host batches travel on a copy stream into
        # staging buffers while the previous replay runs, and a device copy (0.1 ms) moves them into the graph's static
        # inputs.  The copy stream and the staging buffers are created HERE: created lazily in the first call they cost a
        # one-time 30 - 130 ms (two fresh 88 / 44 MB cudaMallocs next to a 45 GB graph pool), which earlier measurements
        # of 8 - 16 steps mistook for a slow mode of the overlap itself (62 - 68 ms per step; 94.9 at N = 2).  In steady
        # state the end-to-end step equals the device step (tools/overlap_probe.py: 59.8 - 60.1 ms in 12 of 12 processes
        # against 61.9 with main-stream copies).  Needs the host to run one step ahead: read losses with loss_dict(lag=1).
        self._overlap_h2d = bool(overlap_h2d)
        if self._overlap_h2d:
            self._copy = torch.cuda.Stream()
            self._sx, self._st = torch.empty_like(self.x), torch.empty_like(self.t)
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream())
        self._step = 0
        self._hout = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._done = [torch.cuda.Event(), torch.cuda.Event()]
        # The warm-up steps below are real optimisation steps (they build the flat gradient bucket, the optimizer state
        # and the allocator's pools).  Their effect on the model and optimizer is undone before capture, so the caller
        # gets back exactly the state it handed in (a resumed checkpoint is not perturbed, BN running statistics and
        # num_batches_tracked included); the reference loop (main.py:125-144) has no such side effect.
        snap = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(warmup, 2)):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        ops.PACKS.invalidate()                       # weight re-packs must be part of the captured step (one batched launch)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        torch.cuda.synchronize()

    def _snapshot(self):
        model = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        opt = {}
        for p, st in self.opt.state.items():
            opt[id(p)] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return model, opt

    def _restore(self, snap):
        """in place: the tensors keep their addresses (the captured graph and the optimizer refer to them)"""
        model, opt = snap
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model[k])
            for p, st in self.opt.state.items():
                old = opt.get(id(p))
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()             # state created by the warm-up: back to AdamW's initial zeros
        self.red.zero()

    def _body(self):
        self.red.zero()
        ops.PACKS.repack_all()                    # every weight pack of the step in one launch (no-op in the first step)
        pred = self.model(self.x).unsqueeze(1)
        total, out = util.combined_loss_device(pred, self.t, self.cfg, rgb=self.x if self.use_rgb else None)
        ops.side_enable(self.side_wgrad)          # weight gradients overlap the data-gradient chain on a second stream
        ops._Side.reducer = self.red              # ... and are all-reduced bucket by bucket as they complete
        try:
            total.backward()
            ops.side_join()
        finally:
            ops.side_enable(False)
            ops._Side.reducer = None
        self.red.reduce()
        self.opt.step()
        self.out = out

    def __call__(self, inputs=None, targets=None):
        """one optimisation step; inputs/targets may be (pinned) host or device tensors, or None to reuse the static
        batch.  Returns the device tensor of loss scalars (index with depth_b200._lib.L_*); no host sync.

        With overlap_h2d=False host batches are copied into the graph's static inputs on the main stream;
        with overlap_h2d=True (the default) they travel on a copy stream into a staging buffer instead, so the H2D
        transfer of step i overlaps the graph replay of step i-1.  The loss scalars of every step are copied to pinned host memory asynchronously (`loss_dict`)."""
        main = torch.cuda.current_stream()
        host = [(dst, stg, src) for dst, stg, src in ((self.x, self._sx, inputs), (self.t, self._st, targets))
                if src is not None and self._overlap_h2d and not src.is_cuda]
        for dst, src in ((self.x, inputs), (self.t, targets)):
            if src is not None and not (self._overlap_h2d and not src.is_cuda):
                dst.copy_(src, non_blocking=True)      # device tensors (ordered after their producer on this stream),
                                                       # or host tensors with the overlap switched off
        if host:
            cs = self._copy
            cs.wait_event(self._consumed)              # the previous step has finished reading the staging buffers
            with torch.cuda.stream(cs):
                for _, stg, src in host:
                    stg.copy_(src, non_blocking=True)
            main.wait_stream(cs)
            for dst, stg, _ in host:
                dst.copy_(stg, non_blocking=True)
            self._consumed.record(main)
        frozen_cast.refresh_all()                  # frozen-encoder bf16 weight copies the graph reads (version check only)
        self.graph.replay()
        ops.PACKS.invalidate()                     # the replay moved the weights without bumping their _version
        slot = self._step & 1
        self._hout[slot].copy_(self.out, non_blocking=True)
        self._done[slot].record(main)
        self._step += 1
        return self.out

    def loss_dict(self, lag=0):
        """the loss scalars of the last step (lag=0) or of the one before it (lag=1: lets the host run one step ahead of
        the device, which is what overlaps the next batch's H2D copy with the current replay); waits only for that
        step's device->host copy (main.py:85-88 needs four of the scalars)."""
        from . import _lib as L
        assert lag in (0, 1) and self._step > lag
        slot = (self._step - 1 - lag) & 1
        self._done[slot].synchronize()
        h = self._hout[slot].tolist()
        return {"total": h[L.L_TOTAL], "si_loss": h[L.L_SI], "silog_loss": h[L.L_SILOG], "grad_loss": h[L.L_GRAD],
                "edge_loss": h[L.L_EDGE]}

    def close(self):
        """release the captured graph (and with it the NCCL collectives it recorded).  Call before
        torch.distributed.destroy_process_group(): tearing the process group down while a graph that captured its
        collectives is alive hung at interpreter exit (round 1, N = 2)."""
        torch.cuda.synchronize()
        if self.graph is not None:
            self.graph.reset()
            self.graph = None
        self.out = None

    def finish(self):
        """kept for round-1 callers; no longer required: every replay invalidates the eager weight-pack cache."""
        ops.PACKS.invalidate()
