"""ORACLE support (test infrastructure): deterministic weights / inputs shared by the golden-vector
generator (oracle/make_golden.py, which runs the real reference) and the parity tests.

Weights are filled key by key from a generator seeded with crc32(key), so they do not depend on
module construction order and can be reproduced on the GPU box without shipping checkpoints.
"""
import os
import sys
import types
import zlib
import importlib.util

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(_ROOT, "monocular-depth-estimation-cil_b200")


def load_package(alias="depth_b200"):
    """Import the product package (its directory name is not a Python identifier) under `alias`."""
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def load_standins():
    """The encoder stand-ins only (no CUDA library needed)."""
    name = "depth_b200_standins"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "standins.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Cfg(types.SimpleNamespace):
    pass


def model_cfg():
    return Cfg(use_lb=False, use_dgr=False)


def loss_config(si=1.0, silog=0.0, vf=0.85, grad=0.0, edge=0.0):
    """OmegaConf-like object with the keys main.combined_loss reads (config.yaml:34-42)."""
    return Cfg(model=Cfg(loss_function=Cfg(si_loss_alpha=si, silog_loss=Cfg(alpha=silog, variance_focus=vf),
                                           grad_loss_alpha=grad, edge_loss_alpha=edge)))


@torch.no_grad()
def fill_deterministic(module, gain=1.0):
    """Overwrite every parameter and floating-point buffer of `module` in place."""
    sd = module.state_dict()
    for key in sorted(sd.keys()):
        t = sd[key]
        if not t.is_floating_point():
            t.zero_()
            continue
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "running_var":
            t.copy_(1.0 + 0.2 * torch.rand(t.shape, generator=g))
        elif leaf == "running_mean":
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1 and leaf == "weight":          # norm scales
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1 or leaf in ("cls_token", "pos_scale"):   # biases and friends
            t.copy_(0.05 * torch.randn(t.shape, generator=g))
        else:
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            if "ConvTranspose" in key:
                fan_in = t.shape[0]
            t.copy_(gain * torch.randn(t.shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5)
    return module


def seeded(shape, seed, kind="randn", lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        return torch.randn(shape, generator=g)
    return torch.rand(shape, generator=g) * (hi - lo) + lo


def subsample(t, max_elems=20000):
    f = t.detach().reshape(-1).float()
    step = max(1, (f.numel() + max_elems - 1) // max_elems)
    return f[::step].contiguous()
