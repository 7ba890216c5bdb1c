"""Drop-in for the reference's ``src/network/midas_semantics.py`` (the default model, midas_semantics.py:14-267)."""
import torch
import torch.nn as nn

from .. import ops
from . import blocks as _blocks
from . import frozen_cast
from .blocks import enter, leave
from .dpt_depth import Dinov2Head
from .midas_net_custom import MidasNet_small


class CrossAttention(nn.Module):
    """reference midas_semantics.py:14-127.  The window loop is evaluated in its exact last-writer closed form
    (csrc/attention.cu); the stride-2 / transposed convs run on tcgen05 (parity planes / output phases); BatchNorm is applied per call,
    so the shared spatial_reduction BN is updated twice per forward exactly as in the reference."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, window_size=16):
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.window_size = window_size
        self.norm_q = nn.LayerNorm(dim)
        self.norm_k = nn.LayerNorm(dim)
        self.norm_v = nn.LayerNorm(dim)
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.norm_out = nn.LayerNorm(dim)
        self.proj = nn.Linear(dim, dim)

        def down():
            return [nn.Conv2d(dim, dim, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(dim), nn.ReLU(inplace=True)]

        def up():
            return [nn.ConvTranspose2d(dim, dim, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(dim), nn.ReLU(inplace=True)]

        self.spatial_reduction = nn.Sequential(*down(), *down(), *down())
        self.spatial_upsample = nn.Sequential(*up(), *up(), *up())

    def _reduce(self, t):
        sr = self.spatial_reduction
        for i in (0, 3, 6):
            c = ops.conv_strided(t, sr[i].weight, sr[i].bias, 2, 1)
            t = ops.bn_act(sr[i + 1], c, relu=True)
        return t

    def fused(self, x, context):
        B, H, W, C = x.shape
        if C != 32 or self.num_heads != 8:
            raise NotImplementedError("the B200 attention kernels cover dim=32 / 8 heads (features=64, the reference default)")
        xr = self._reduce(x)
        cr = self._reduce(context)
        hr, wr = H // 8, W // 8
        assert xr.shape[1] == hr and xr.shape[2] == wr, "input size must be a multiple of 8"
        N = hr * wr
        q = ops.ln_linear(xr, self.norm_q, self.q).reshape(B, N, C)
        k = ops.ln_linear(cr, self.norm_k, self.k).reshape(B, N, C)
        v = ops.ln_linear(cr, self.norm_v, self.v).reshape(B, N, C)
        o = ops.attention(q, k, v, hr, wr, self.window_size, self.scale)
        p = ops.ln_linear(o, self.norm_out, self.proj, out_bf16=True).reshape(B, hr, wr, C)
        su = self.spatial_upsample
        for i in (0, 3, 6):
            c = ops.conv_transposed(p, su[i].weight, su[i].bias, 2, 1)
            p = ops.bn_act(su[i + 1], c, relu=True)
        return ops.add(p, x)

    def forward(self, x, context):
        t, pub = enter(x)
        c, _ = enter(context)
        return leave(self.fused(t, c), pub)


class ResidualBlock(nn.Module):
    """reference midas_semantics.py:129-151.  Train-mode BN statistics come out of the conv epilogues."""

    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        assert stride == 1, "the reference only instantiates stride-1 blocks (midas_semantics.py:185-202)"
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(out_channels))

    def fused(self, x):
        """one autograd node (ops.res_block): bn1 + ReLU live in conv2's operand path and in the epilogue of its data
        gradient, the shortcut's gradient rides in conv1's data-gradient epilogue"""
        if self.conv1.weight.shape[0] % 8 or self.conv1.weight.shape[1] % 8:
            raise NotImplementedError("ResidualBlock on the tensor-core path needs channel counts that are multiples of 8")
        sc = None if len(self.shortcut) == 0 else (self.shortcut[0], self.shortcut[1])
        return ops.res_block(x, self.conv1, self.bn1, self.conv2, self.bn2, sc)

    def forward(self, x):
        t, pub = enter(x)
        return leave(self.fused(t), pub)


class MidasNetSemantics(MidasNet_small):
    """reference midas_semantics.py:153-267."""

    def __init__(self, path=None, features=32, backbone="efficientnet_lite3", non_negative=True, exportable=True,
                 channels_last=False, align_corners=True, cfg=None, blocks={'expand': True},
                 dinov2_type='dinov2_vits14'):
        super().__init__(path, features, backbone, non_negative, exportable, channels_last, align_corners, cfg, blocks)
        self.scratch.output_conv = self.scratch.output_conv[0:4] + self.scratch.output_conv[6:]
        self.dinov2 = _blocks._hub('facebookresearch/dinov2', dinov2_type)
        for param in self.dinov2.parameters():
            param.requires_grad = False
        dim = self.dinov2.blocks[0].attn.qkv.in_features
        self.dinov2_head = Dinov2Head(1, dim, 128, use_bn=False, out_channels=[128, 256, 512, 512], use_clstoken=False)
        self.DINOv2_IMAGE_SIZE = (224, 280)
        self.cross_attention = CrossAttention(features // 2, window_size=16)
        self.fusion_blocks = nn.Sequential(ResidualBlock(features, features))
        self.fusion_head = nn.Sequential(
            ResidualBlock(features, features // 2),
            nn.Conv2d(features // 2, features // 2, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(features // 2),
            nn.ReLU(True),
        )
        self.depth_head = nn.Sequential(
            ResidualBlock(features // 2, features // 4),
            nn.Conv2d(features // 4, 1, kernel_size=3, stride=1, padding=1),
            nn.ReLU(True) if non_negative else nn.Identity(),
        )

    def dino_tokens(self, x):
        x_dinov2 = ops.resize_planes_f32(x, self.DINOv2_IMAGE_SIZE, True)
        if self.encoder_autocast:
            frozen_cast.enable(self.dinov2)       # the frozen branch keeps bf16 copies of its Linear weights
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.dinov2.get_intermediate_layers(x_dinov2, 4, return_class_token=False)
        return self.dinov2.get_intermediate_layers(x_dinov2, 4, return_class_token=False)

    def forward(self, x):
        midas = self.head_features(self.decoder_trunk(self.encoder_features(x)))            # (B,H,W,32)
        ph, pw = self.DINOv2_IMAGE_SIZE[0] // 14, self.DINOv2_IMAGE_SIZE[1] // 14
        dino = self.dinov2_head.fused(self.dino_tokens(x), ph, pw)                          # (B,224,280,32)
        dino = ops.resize(dino, midas.shape[1:3], True)
        attended = self.cross_attention.fused(midas, dino)
        y = ops.concat_channels(attended, midas)
        for blk in self.fusion_blocks:
            y = blk.fused(y)
        fh = self.fusion_head
        y = fh[0].fused(y)
        tr = fh[2].training
        r = ops.conv_tc(y, fh[1].weight, fh[1].bias, stats=tr)
        c, st = r if tr else (r, None)
        y = ops.bn_act(fh[2], c, st, relu=True)
        dh = self.depth_head
        y = dh[0].fused(y)
        if self.use_lb:
            raise NotImplementedError
        return ops.head_conv(y, dh[1].weight, dh[1].bias, isinstance(dh[2], nn.ReLU))       # (B,H,W) fp32
