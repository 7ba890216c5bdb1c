"""bf16 copies of a frozen encoder's nn.Linear parameters for its autocast forward.

torch.autocast re-casts every fp32 weight it meets on every forward: its cast cache only holds leaves that require
grad, and it is dropped when the context exits.  For the frozen DINOv2 branch of MidasNetSemantics (reference
midas_semantics.py:233-239, run under bf16 autocast by `encoder_autocast`) that is 127 cast kernels per training step
for parameters that never change.  `enable(root)` gives every frozen nn.Linear under `root` a forward that keeps its own
low-precision copies and hands them to F.linear - the same kernel on the same values, so results are bit-identical.

The copies live outside the module's parameters / buffers (state_dict unchanged).  They are refreshed IN PLACE when the
parameter's version counter moves (load_state_dict, manual edits), so a CUDA graph that captured them keeps reading
current values; GraphedTrainStep calls `refresh_all()` before every replay for exactly that case.
"""
import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

_PATCHED = weakref.WeakSet()


def _key(m, dt):
    b = m.bias
    return (m.weight._version, m.weight.data_ptr(), dt, -1 if b is None else b._version, 0 if b is None else b.data_ptr())


def _casts(m, dt):
    c = m.__dict__.get("_dp_cast")
    k = _key(m, dt)
    if c is None or c[0][1] != k[1] or c[0][2] != dt or c[0][4] != k[4]:
        with torch.no_grad():
            c = [k, m.weight.detach().to(dt), None if m.bias is None else m.bias.detach().to(dt)]
        m.__dict__["_dp_cast"] = c
    elif c[0] != k:                                   # same storage, new values: refresh in place
        with torch.no_grad():
            c[1].copy_(m.weight)
            if c[2] is not None:
                c[2].copy_(m.bias)
        c[0] = k
    return c


def _forward(self, x):
    w = self.weight
    if w.requires_grad or not w.is_cuda or not torch.is_autocast_enabled("cuda"):
        return F.linear(x, w, self.bias)
    c = _casts(self, torch.get_autocast_dtype("cuda"))
    return F.linear(x, c[1], c[2])


def enable(root):
    """idempotent; only nn.Linear modules whose parameters do not require grad are touched"""
    if root.__dict__.get("_dp_cast_enabled"):
        return
    for m in root.modules():
        if type(m) is nn.Linear and not m.weight.requires_grad and (m.bias is None or not m.bias.requires_grad):
            m.forward = _forward.__get__(m, nn.Linear)
            _PATCHED.add(m)
    root.__dict__["_dp_cast_enabled"] = True


def refresh_all():
    """bring every existing low-precision copy up to date with its parameter (cheap: version counters only)"""
    for m in list(_PATCHED):
        c = m.__dict__.get("_dp_cast")
        if c is not None and c[0] != _key(m, c[0][2]):
            _casts(m, c[0][2])
