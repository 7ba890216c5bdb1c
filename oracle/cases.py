"""ORACLE support: the module-level parity cases (shapes, seeds) shared by make_golden.py and tests.

A case is run as: build module -> fill_deterministic -> train() -> forward on seeded inputs ->
backward with a seeded cotangent -> collect {out, input grads, parameter grads, float buffers}.
"""
import torch

from . import fixtures as fx
from . import model as om

CASES = {
    # name: (kind, ctor kwargs, list of input shapes, extra forward kwargs)
    "rcu64": ("rcu", dict(features=64), [(2, 64, 12, 20)], {}),
    "rcu32_odd": ("rcu", dict(features=32), [(1, 32, 7, 9)], {}),
    "fusion128_expand": ("fusion", dict(features=128, expand=True), [(2, 128, 10, 14), (2, 128, 10, 14)], {}),
    "fusion64_single": ("fusion", dict(features=64, expand=False), [(2, 64, 9, 11)], {}),
    "fusion128_size": ("fusion", dict(features=128, expand=False), [(1, 128, 8, 10), (1, 128, 8, 10)],
                       dict(size=(16, 20))),
    "resblock_64_64": ("resblock", dict(cin=64, cout=64), [(2, 64, 24, 40)], {}),
    "resblock_64_32": ("resblock", dict(cin=64, cout=32), [(2, 64, 24, 40)], {}),
    "resblock_32_16": ("resblock", dict(cin=32, cout=16), [(2, 32, 16, 24)], {}),
    "dinohead": ("dinohead", dict(), [(2, 20, 384)] * 4, dict(ph=4, pw=5)),
    "xattn_1win": ("xattn", dict(dim=32), [(2, 32, 64, 96), (2, 32, 64, 96)], {}),
    "xattn_multi": ("xattn", dict(dim=32), [(1, 32, 160, 192), (1, 32, 160, 192)], {}),
    # MiDaS-large / DPT decoders: blocks.py:243-314, midas_net.py:12-76, dpt_depth.py:155-293
    "rcu_large64": ("rcu_large", dict(features=64), [(2, 64, 10, 14)], {}),
    "fusion_large64": ("fusion_large", dict(features=64), [(2, 64, 9, 12), (2, 64, 9, 12)], {}),
    "fusion_large64_single": ("fusion_large", dict(features=64), [(1, 64, 7, 10)], {}),
    "dpt_decoder64": ("dpt", dict(features=64), [(1, 256, 24, 32), (1, 512, 12, 16), (1, 768, 6, 8), (1, 768, 3, 4)], {}),
    "midas_large_decoder64": ("midas_large", dict(features=64),
                              [(1, 256, 16, 24), (1, 512, 8, 12), (1, 1024, 4, 6), (1, 2048, 2, 3)], {}),
}


def build_oracle(kind, kw):
    if kind == "rcu":
        return om.RCU(kw["features"])
    if kind == "fusion":
        return om.FusionBlock(kw["features"], expand=kw["expand"], align_corners=True)
    if kind == "resblock":
        return om.ResidualBlock(kw["cin"], kw["cout"])
    if kind == "dinohead":
        return om.Dinov2Head(1, 384, 128, use_bn=False, out_channels=[128, 256, 512, 512], use_clstoken=False)
    if kind == "xattn":
        return om.CrossAttention(kw["dim"], window_size=16)
    if kind == "rcu_large":
        return om.RCUInplace(kw["features"])
    if kind == "fusion_large":
        return om.FusionBlockLarge(kw["features"])
    if kind == "dpt":
        return om.DPTDecoder(features=kw["features"])
    if kind == "midas_large":
        return om.MidasLargeDecoder(features=kw["features"])
    raise KeyError(kind)


def case_inputs(name):
    kind, kw, shapes, fkw = CASES[name]
    base = (abs(hash(name)) if False else sum(ord(c) * (i + 1) for i, c in enumerate(name))) % 100000
    return [fx.seeded(s, base + i) for i, s in enumerate(shapes)]


def run_case(module, name, device="cpu", call=None):
    """Returns dict of tensors (on CPU, fp32)."""
    kind, kw, shapes, fkw = CASES[name]
    module = module.to(device)
    module.train()
    xs = [x.to(device).requires_grad_(True) for x in case_inputs(name)]
    if kind in ("rcu_large", "fusion_large"):
        # these blocks apply an in-place ReLU to their input (blocks.py:263,274): feed non-leaf copies
        fed = [x * 1.0 for x in xs]
        out = call(module, fed, fkw) if call is not None else module(*fed)
    elif call is not None:
        out = call(module, xs, fkw)
    elif kind == "dinohead":
        out = module(tuple(xs), fkw["ph"], fkw["pw"])
    elif kind == "fusion":
        out = module(*xs, **fkw)
    else:
        out = module(*xs)
    cot = fx.seeded(tuple(out.shape), 777, "randn").to(device)
    (out.float() * cot).sum().backward()
    res = {"out": out.detach().float().cpu()}
    for i, x in enumerate(xs):
        res[f"gin{i}"] = x.grad.detach().float().cpu()
    for k, p in module.named_parameters():
        if p.grad is not None:
            res[f"gp.{k}"] = p.grad.detach().float().cpu()
    for k, b in module.named_buffers():
        res[f"buf.{k}"] = b.detach().float().cpu()
    return res


# ---- full default model ------------------------------------------------------------------------
FULL_INPUT = (2, 3, 64, 96)


def build_oracle_semantics(standins):
    m = om.MidasNetSemantics(None, features=64, backbone="efficientnet_lite3", exportable=True, non_negative=True,
                             cfg=fx.model_cfg(), blocks={"expand": True}, dinov2_type="dinov2_vits14",
                             hub_load=standins.hub_load_standin)
    return m


def build_oracle_small(standins):
    return om.MidasNet_small(None, features=64, backbone="efficientnet_lite3", exportable=True, non_negative=True,
                             cfg=fx.model_cfg(), blocks={"expand": True}, hub_load=standins.hub_load_standin)


def prepare_full(model):
    """deterministic weights; lift the last bias so the ReLU'd depth stays clear of 0 (the SI loss gradient
    ~ 1/(p+1e-6) makes near-zero predictions dominate and turns gradient comparisons chaotic)."""
    fx.fill_deterministic(model)
    with torch.no_grad():
        if hasattr(model, "depth_head"):
            model.depth_head[1].bias.add_(4.0)
        else:
            model.scratch.output_conv[4].bias.add_(8.0)
    return model


def full_batch():
    x = fx.seeded(FULL_INPUT, 4242)
    t = fx.seeded((FULL_INPUT[0], 1) + FULL_INPUT[2:], 4243, "rand", 0.1, 10.0)
    return x, t
