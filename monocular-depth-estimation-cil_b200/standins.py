"""Deterministic stand-ins for the two third-party encoders the reference pulls from torch.hub.

The reference builds its encoders with ``torch.hub.load`` (reference ``src/network/blocks.py:166-173``
for ``tf_efficientnet_lite3`` and ``src/network/midas_semantics.py:168`` for DINOv2).  Neither the
hub source nor the weights are available offline, and both are outside the hot-path scope
(rows 8 and 10 are "out of scope - third-party arithmetic").  What the in-scope decoder needs
is only the *feature-map contract*: channels 32/48/136/384 at strides 4/8/16/32 for the
EfficientNet-Lite3 trunk, and ``get_intermediate_layers(x, n, return_class_token=False)`` giving
n tensors of shape (B, (H/14)*(W/14), 384) plus ``.blocks[0].attn.qkv.in_features`` for the ViT.

These stand-ins run through plain PyTorch (cuDNN/cuBLAS on the GPU); the same module instance
feeds the oracle and the CUDA decoder in every parity test so that encoder arithmetic cancels.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# EfficientNet-Lite3-shaped trunk
# ----------------------------------------------------------------------------------------------
class _DSBlock(nn.Module):
    """depthwise-separable block (first stage of the trunk)."""

    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin, eps=1e-3)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout, eps=1e-3)
        self.has_residual = stride == 1 and cin == cout

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_dw(x)))
        y = self.bn2(self.conv_pw(y))
        return x + y if self.has_residual else y


class _IRBlock(nn.Module):
    """inverted-residual block, expansion 6, no squeeze-excite (the Lite variants drop SE)."""

    def __init__(self, cin, cout, k, stride, expand=6):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid, eps=1e-3)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid, eps=1e-3)
        self.act2 = nn.ReLU6(inplace=True)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout, eps=1e-3)
        self.has_residual = stride == 1 and cin == cout

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return x + y if self.has_residual else y


class EfficientNetLite3StandIn(nn.Module):
    """Exposes conv_stem, bn1, act1, blocks[0..6] exactly as consumed at reference blocks.py:176-186."""

    # (kind, kernel, stride, out_channels, repeats)
    _STAGES = [
        ("ds", 3, 1, 24, 1),
        ("ir", 3, 2, 32, 3),
        ("ir", 5, 2, 48, 3),
        ("ir", 3, 2, 96, 5),
        ("ir", 5, 1, 136, 5),
        ("ir", 5, 2, 232, 6),
        ("ir", 3, 1, 384, 1),
    ]

    def __init__(self):
        super().__init__()
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32, eps=1e-3)
        self.act1 = nn.ReLU6(inplace=True)
        blocks = []
        cin = 32
        for kind, k, s, cout, reps in self._STAGES:
            stage = []
            for i in range(reps):
                stride = s if i == 0 else 1
                stage.append(_DSBlock(cin, cout, k, stride) if kind == "ds" else _IRBlock(cin, cout, k, stride))
                cin = cout
            blocks.append(nn.Sequential(*stage))
        self.blocks = nn.Sequential(*blocks)


# ----------------------------------------------------------------------------------------------
# DINOv2 ViT-S/14-shaped frozen branch
# ----------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        y = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(y.transpose(1, 2).reshape(B, N, C))


class _ViTBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class DinoV2StandIn(nn.Module):
    """ViT-S/14 (dim 384, 12 blocks, 6 heads) or ViT-B/14 (768, 12, 12) shaped module."""

    def __init__(self, dim=384, depth=12, heads=6, patch=14):
        super().__init__()
        self.patch_size = patch
        self.patch_embed = nn.Conv2d(3, dim, patch, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_scale = nn.Parameter(torch.ones(1, 1, dim) * 0.02)
        self.blocks = nn.ModuleList([_ViTBlock(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.dim = dim

    def _pos(self, h, w, device, dtype):
        # fixed sin/cos table (resolution-free, so no interpolation is needed)
        ys = torch.arange(h, device=device, dtype=torch.float32)
        xs = torch.arange(w, device=device, dtype=torch.float32)
        d = self.dim // 4
        freq = torch.exp(-math.log(100.0) * torch.arange(d, device=device, dtype=torch.float32) / d)
        py = ys[:, None] * freq[None]
        px = xs[:, None] * freq[None]
        pe = torch.cat([
            py.sin()[:, None, :].expand(h, w, d), py.cos()[:, None, :].expand(h, w, d),
            px.sin()[None, :, :].expand(h, w, d), px.cos()[None, :, :].expand(h, w, d)], dim=-1)
        return pe.reshape(1, h * w, self.dim).to(dtype)

    def get_intermediate_layers(self, x, n=1, return_class_token=False):
        B = x.shape[0]
        t = self.patch_embed(x)
        h, w = t.shape[-2:]
        t = t.flatten(2).transpose(1, 2)
        t = t + self._pos(h, w, t.device, t.dtype) * self.pos_scale.to(t.dtype)
        t = torch.cat([self.cls_token.to(t.dtype).expand(B, -1, -1), t], dim=1)
        outs = []
        first = len(self.blocks) - n
        for i, blk in enumerate(self.blocks):
            t = blk(t)
            if i >= first:
                outs.append(self.norm(t))
        if return_class_token:
            return tuple((o[:, 1:], o[:, 0]) for o in outs)
        return tuple(o[:, 1:] for o in outs)


def hub_load_standin(repo, name, *args, **kwargs):
    """Drop-in for ``torch.hub.load`` covering the two hub models of the default path."""
    if "efficientnet" in name:
        return EfficientNetLite3StandIn()
    if name == "dinov2_vits14":
        return DinoV2StandIn(384, 12, 6)
    if name == "dinov2_vitb14":
        return DinoV2StandIn(768, 12, 12)
    raise ValueError(f"no offline stand-in for torch.hub model {repo}:{name}")
