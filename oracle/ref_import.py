"""ORACLE support: import the UNMODIFIED reference from /root/reference (build container only).

The reference needs `timm` at import time (src/network/blocks.py:4 -> backbones/beit.py:1) and
`torch.hub.load` (network access) at model-construction time.  This shim inserts inert `timm` stub
modules, puts /root/reference/src on sys.path and swaps `torch.hub.load` for the package's offline
stand-ins while a model is being built.  Nothing here is available on the GPU box (no
/root/reference there); it is used only by oracle/make_golden.py to pin the oracle.
"""
import contextlib
import os
import sys
import types

import torch

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


def _stub_timm():
    if "timm" in sys.modules:
        return
    for name in ("timm", "timm.models", "timm.models.beit", "timm.models.layers"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["timm"].create_model = lambda *a, **k: None
    sys.modules["timm.models.beit"].gen_relative_position_index = lambda *a, **k: None
    sys.modules["timm.models.layers"].get_act_layer = lambda *a, **k: None
    sys.modules["timm"].models = sys.modules["timm.models"]


def import_reference():
    """returns (util, blocks, dpt_depth, midas_semantics, midas_net_custom) reference modules."""
    assert available(), "reference tree not present"
    sys.dont_write_bytecode = True
    _stub_timm()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import util as ref_util                                   # noqa
    import network.blocks as ref_blocks                       # noqa
    import network.dpt_depth as ref_dpt                       # noqa
    import network.midas_semantics as ref_sem                 # noqa
    import network.midas_net_custom as ref_small              # noqa
    import network.midas_net as ref_large                     # noqa
    return ref_util, ref_blocks, ref_dpt, ref_sem, ref_small, ref_large


@contextlib.contextmanager
def offline_hub(hub_load):
    real = torch.hub.load
    torch.hub.load = hub_load
    try:
        yield
    finally:
        torch.hub.load = real


def import_reference_main():
    """the reference's src/main.py (evaluate_model, combined_loss live there).  kornia / omegaconf are absent from this
    image and wandb must not start a session: inert stubs stand in for them (none is touched by evaluate_model)."""
    import_reference()
    for name in ("kornia", "kornia.augmentation", "kornia.geometry", "omegaconf", "wandb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].augmentation = sys.modules["kornia.augmentation"]
    sys.modules["kornia"].geometry = sys.modules["kornia.geometry"]
    if not hasattr(sys.modules["omegaconf"], "OmegaConf"):
        sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {})
    import main as ref_main                                   # noqa
    return ref_main
