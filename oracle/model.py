"""ORACLE (test infrastructure, never shipped, never on the product path).

Plain-PyTorch fp32 restatement of the reference's in-repo encoder-decoder modules with the same
constructor arguments, attribute names and state_dict keys, so a reference checkpoint loads into it
and vice versa.  oracle/make_golden.py checks these against the real reference modules (imported from
/root/reference in the build container) with copied state_dicts; outputs and gradients must agree
to fp32 round-off.  Third-party encoders are the stand-ins from the package (identical instance on
both sides of every parity test).

Citations are to /root/reference/src/network/*.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _bilinear(x, align_corners, size=None, scale_factor=None):
    return F.interpolate(x, size=size, scale_factor=scale_factor, mode="bilinear", align_corners=align_corners)


class Upsample2x(nn.Module):
    """blocks.py:208-240 `Interpolate` (default align_corners=False)."""

    def __init__(self, scale_factor, mode, align_corners=False):
        super().__init__()
        self.scale_factor, self.mode, self.align_corners = scale_factor, mode, align_corners

    def forward(self, x):
        return F.interpolate(x, scale_factor=self.scale_factor, mode=self.mode, align_corners=self.align_corners)


class RCU(nn.Module):
    """blocks.py:319-376 `ResidualConvUnit_custom` with bn=False: x + conv2(relu(conv1(relu(x))))."""

    def __init__(self, features):
        super().__init__()
        self.conv1 = nn.Conv2d(features, features, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(features, features, 3, 1, 1, bias=True)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class RCUInplace(nn.Module):
    """blocks.py:243-279 `ResidualConvUnit`: the first ReLU is in place, so the skip adds relu(x)
    and the caller's tensor is overwritten with relu(x)."""

    def __init__(self, features):
        super().__init__()
        self.conv1 = nn.Conv2d(features, features, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(features, features, 3, 1, 1, bias=True)

    def forward(self, x):
        r = F.relu(x)
        return self.conv2(F.relu(self.conv1(r))) + r


class FusionBlock(nn.Module):
    """blocks.py:379-438 `FeatureFusionBlock_custom`."""

    def __init__(self, features, expand=False, align_corners=True, size=None):
        super().__init__()
        self.align_corners = align_corners
        self.size = size
        self.out_conv = nn.Conv2d(features, features // 2 if expand else features, 1, bias=True)
        self.resConfUnit1 = RCU(features)
        self.resConfUnit2 = RCU(features)

    def forward(self, *xs, size=None):
        y = xs[0]
        if len(xs) == 2:
            y = y + self.resConfUnit1(xs[1])
        y = self.resConfUnit2(y)
        if size is None and self.size is None:
            y = _bilinear(y, self.align_corners, scale_factor=2)
        else:
            y = _bilinear(y, self.align_corners, size=self.size if size is None else size)
        return self.out_conv(y)


class FusionBlockLarge(nn.Module):
    """blocks.py:282-314 `FeatureFusionBlock` (MiDaS v2.1 large)."""

    def __init__(self, features):
        super().__init__()
        self.resConfUnit1 = RCUInplace(features)
        self.resConfUnit2 = RCUInplace(features)

    def forward(self, *xs):
        y = xs[0]
        if len(xs) == 2:
            y = y + self.resConfUnit1(xs[1])
        y = self.resConfUnit2(y)
        return _bilinear(y, True, scale_factor=2)


def make_scratch(in_shape, out_shape, expand=False):
    """blocks.py:133-163."""
    s = nn.Module()
    mult = [1, 2, 4, 8] if expand else [1, 1, 1, 1]
    for i, cin in enumerate(in_shape):
        setattr(s, f"layer{i + 1}_rn", nn.Conv2d(cin, out_shape * mult[i], 3, 1, 1, bias=False))
    return s


def make_efficientnet_backbone(effnet):
    """blocks.py:176-186."""
    p = nn.Module()
    p.layer1 = nn.Sequential(effnet.conv_stem, effnet.bn1, effnet.act1, *effnet.blocks[0:2])
    p.layer2 = nn.Sequential(*effnet.blocks[2:3])
    p.layer3 = nn.Sequential(*effnet.blocks[3:5])
    p.layer4 = nn.Sequential(*effnet.blocks[5:9])
    return p


class MidasNet_small(nn.Module):
    """midas_net_custom.py:45-185 (use_lb / use_dgr off)."""

    def __init__(self, path=None, features=64, backbone="efficientnet_lite3", non_negative=True, exportable=True,
                 channels_last=False, align_corners=True, cfg=None, blocks={"expand": True}, hub_load=None):
        super().__init__()
        assert backbone == "efficientnet_lite3"
        assert not (cfg.use_lb or cfg.use_dgr), "oracle covers the default path only"
        self.use_lb = False
        expand = bool(blocks.get("expand", False))
        f = [features, features * 2, features * 4, features * 8] if expand else [features] * 4
        self.pretrained = make_efficientnet_backbone(hub_load("rwightman/gen-efficientnet-pytorch", "tf_efficientnet_lite3"))
        self.scratch = make_scratch([32, 48, 136, 384], features, expand=expand)
        self.scratch.activation = nn.ReLU(False)
        self.scratch.refinenet4 = FusionBlock(f[3], expand=expand, align_corners=align_corners)
        self.scratch.refinenet3 = FusionBlock(f[2], expand=expand, align_corners=align_corners)
        self.scratch.refinenet2 = FusionBlock(f[1], expand=expand, align_corners=align_corners)
        self.scratch.refinenet1 = FusionBlock(f[0], expand=False, align_corners=align_corners)
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, features // 2, 3, 1, 1),
            Upsample2x(2, "bilinear"),
            nn.Conv2d(features // 2, 32, 3, 1, 1),
            self.scratch.activation,
            nn.Conv2d(32, 1, 1),
            nn.ReLU(True) if non_negative else nn.Identity(),
            nn.Identity(),
        )

    def decoder(self, x):
        l1 = self.pretrained.layer1(x)
        l2 = self.pretrained.layer2(l1)
        l3 = self.pretrained.layer3(l2)
        l4 = self.pretrained.layer4(l3)
        s = self.scratch
        p4 = s.refinenet4(s.layer4_rn(l4))
        p3 = s.refinenet3(p4, s.layer3_rn(l3))
        p2 = s.refinenet2(p3, s.layer2_rn(l2))
        p1 = s.refinenet1(p2, s.layer1_rn(l1))
        return s.output_conv(p1)

    def forward(self, x):
        return torch.squeeze(self.decoder(x), dim=1)


class Dinov2Head(nn.Module):
    """dpt_depth.py:32-153 (nclass=1, use_clstoken=False)."""

    def __init__(self, nclass, in_channels, features=256, use_bn=False, out_channels=(256, 512, 1024, 1024),
                 use_clstoken=False):
        super().__init__()
        assert nclass == 1 and not use_bn and not use_clstoken
        oc = list(out_channels)
        self.projects = nn.ModuleList([nn.Conv2d(in_channels, c, 1) for c in oc])
        self.resize_layers = nn.ModuleList([
            nn.ConvTranspose2d(oc[0], oc[0], 4, 4, 0),
            nn.ConvTranspose2d(oc[1], oc[1], 2, 2, 0),
            nn.Identity(),
            nn.Conv2d(oc[3], oc[3], 3, 2, 1),
        ])
        self.scratch = make_scratch(oc, features, expand=False)
        for i in (1, 2, 3, 4):
            setattr(self.scratch, f"refinenet{i}", FusionBlock(features, expand=False, align_corners=True))
        self.scratch.output_conv1 = nn.Conv2d(features, features // 2, 3, 1, 1)
        self.scratch.output_conv2 = nn.Sequential(nn.Conv2d(features // 2, 32, 3, 1, 1), nn.ReLU(True), nn.Identity())

    def forward(self, feats, ph, pw):
        maps = []
        for i, t in enumerate(feats):
            m = t.permute(0, 2, 1).reshape(t.shape[0], t.shape[-1], ph, pw)
            maps.append(self.resize_layers[i](self.projects[i](m)))
        s = self.scratch
        r = [s.layer1_rn(maps[0]), s.layer2_rn(maps[1]), s.layer3_rn(maps[2]), s.layer4_rn(maps[3])]
        p4 = s.refinenet4(r[3], size=r[2].shape[2:])
        p3 = s.refinenet3(p4, r[2], size=r[1].shape[2:])
        p2 = s.refinenet2(p3, r[1], size=r[0].shape[2:])
        p1 = s.refinenet1(p2, r[0])
        y = s.output_conv1(p1)
        y = _bilinear(y, True, size=(int(ph * 14), int(pw * 14)))
        return s.output_conv2(y)


class CrossAttention(nn.Module):
    """midas_semantics.py:14-127, restated with the reference's own window loop (later ranges overwrite
    earlier ones).  `last_writer=True` evaluates the equivalent closed form used by the CUDA kernel."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, window_size=16):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads, self.head_dim, self.window_size = num_heads, dim // num_heads, window_size
        self.scale = self.head_dim ** -0.5
        self.norm_q, self.norm_k, self.norm_v = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.norm_out = nn.LayerNorm(dim)
        self.proj = nn.Linear(dim, dim)

        def down():
            return [nn.Conv2d(dim, dim, 3, 2, 1), nn.BatchNorm2d(dim), nn.ReLU(inplace=True)]

        def up():
            return [nn.ConvTranspose2d(dim, dim, 4, 2, 1), nn.BatchNorm2d(dim), nn.ReLU(inplace=True)]

        self.spatial_reduction = nn.Sequential(*down(), *down(), *down())
        self.spatial_upsample = nn.Sequential(*up(), *up(), *up())

    @staticmethod
    def window_ranges(hr, wr, ws):
        """token ranges [lo, hi) in the reference's iteration order (midas_semantics.py:93-104)."""
        out = []
        for h in range((hr + ws - 1) // ws):
            for w in range((wr + ws - 1) // ws):
                lo = h * ws * wr + w * ws
                hi = min(h * ws + ws, hr) * wr + min(w * ws + ws, wr)
                out.append((lo, min(hi, hr * wr)))
        return out

    def forward(self, x, context):
        B, C, H, W = x.shape
        xr = self.spatial_reduction(x)
        cr = self.spatial_reduction(context)  # same conv weights and BN as for x (second BN update)
        xf = xr.flatten(2).transpose(1, 2)
        cf = cr.flatten(2).transpose(1, 2)
        nh, hd = self.num_heads, self.head_dim
        q = self.q(self.norm_q(xf)).reshape(B, -1, nh, hd).permute(0, 2, 1, 3)
        k = self.k(self.norm_k(cf)).reshape(B, -1, nh, hd).permute(0, 2, 1, 3)
        v = self.v(self.norm_v(cf)).reshape(B, -1, nh, hd).permute(0, 2, 1, 3)
        hr, wr = H // 8, W // 8
        out = torch.zeros_like(xf)
        for lo, hi in self.window_ranges(hr, wr, self.window_size):
            a = ((q[:, :, lo:hi] @ k[:, :, lo:hi].transpose(-2, -1)) * self.scale).softmax(dim=-1)
            out[:, lo:hi, :] = (a @ v[:, :, lo:hi]).transpose(1, 2).reshape(B, -1, C)
        out = self.proj(self.norm_out(out))
        out = out.transpose(1, 2).reshape(B, C, hr, wr)
        return self.spatial_upsample(out) + x


class ResidualBlock(nn.Module):
    """midas_semantics.py:129-151."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y)) + self.shortcut(x)
        return F.relu(y)


class MidasNetSemantics(MidasNet_small):
    """midas_semantics.py:153-267."""

    def __init__(self, path=None, features=32, backbone="efficientnet_lite3", non_negative=True, exportable=True,
                 channels_last=False, align_corners=True, cfg=None, blocks={"expand": True},
                 dinov2_type="dinov2_vits14", hub_load=None):
        super().__init__(path, features, backbone, non_negative, exportable, channels_last, align_corners, cfg, blocks,
                         hub_load=hub_load)
        oc = self.scratch.output_conv
        self.scratch.output_conv = oc[0:4] + oc[6:]
        self.dinov2 = hub_load("facebookresearch/dinov2", dinov2_type)
        for p in self.dinov2.parameters():
            p.requires_grad = False
        dim = self.dinov2.blocks[0].attn.qkv.in_features
        self.dinov2_head = Dinov2Head(1, dim, 128, use_bn=False, out_channels=[128, 256, 512, 512], use_clstoken=False)
        self.DINOv2_IMAGE_SIZE = (224, 280)
        self.cross_attention = CrossAttention(features // 2, window_size=16)
        self.fusion_blocks = nn.Sequential(ResidualBlock(features, features))
        self.fusion_head = nn.Sequential(
            ResidualBlock(features, features // 2),
            nn.Conv2d(features // 2, features // 2, 3, 1, 1),
            nn.BatchNorm2d(features // 2),
            nn.ReLU(True),
        )
        self.depth_head = nn.Sequential(
            ResidualBlock(features // 2, features // 4),
            nn.Conv2d(features // 4, 1, 3, 1, 1),
            nn.ReLU(True) if non_negative else nn.Identity(),
        )

    def forward(self, x):
        midas = self.decoder(x)
        xd = _bilinear(x, True, size=self.DINOv2_IMAGE_SIZE)
        ph, pw = self.DINOv2_IMAGE_SIZE[0] // 14, self.DINOv2_IMAGE_SIZE[1] // 14
        toks = self.dinov2.get_intermediate_layers(xd, 4, return_class_token=False)
        dino = self.dinov2_head(toks, ph, pw)
        dino = _bilinear(dino, True, size=midas.shape[2:])
        att = self.cross_attention(midas, dino)
        y = self.fusion_blocks(torch.cat([att, midas], dim=1))
        y = self.fusion_head(y)
        return torch.squeeze(self.depth_head(y), dim=1)


class DPTDecoder(nn.Module):
    """Decoder + head of dpt_depth.py:155-293 (`DPT`/`DPTDepthModel`) fed with the four reassembled
    feature maps (the timm backbones are third-party and absent): layerN_rn (no expand) ->
    FusionBlock x4 with size= of the next level -> head conv3x3, x2 (align_corners=True), conv3x3, ReLU,
    conv1x1, ReLU."""

    def __init__(self, in_shape=(256, 512, 768, 768), features=256, head_features_2=32, non_negative=True):
        super().__init__()
        self.scratch = make_scratch(list(in_shape), features, expand=False)
        for i in (1, 2, 3, 4):
            setattr(self.scratch, f"refinenet{i}", FusionBlock(features, expand=False, align_corners=True))
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, features // 2, 3, 1, 1),
            Upsample2x(2, "bilinear", align_corners=True),
            nn.Conv2d(features // 2, head_features_2, 3, 1, 1),
            nn.ReLU(True),
            nn.Conv2d(head_features_2, 1, 1),
            nn.ReLU(True) if non_negative else nn.Identity(),
            nn.Identity(),
        )

    def forward(self, l1, l2, l3, l4):
        s = self.scratch
        r = [s.layer1_rn(l1), s.layer2_rn(l2), s.layer3_rn(l3), s.layer4_rn(l4)]
        p4 = s.refinenet4(r[3], size=r[2].shape[2:])
        p3 = s.refinenet3(p4, r[2], size=r[1].shape[2:])
        p2 = s.refinenet2(p3, r[1], size=r[0].shape[2:])
        p1 = s.refinenet1(p2, r[0])
        return s.output_conv(p1).squeeze(dim=1)


class MidasLargeDecoder(nn.Module):
    """Decoder + head of midas_net.py:12-76 (`MidasNet`, ResNeXt101-WSL encoder absent) fed with four
    feature maps [256,512,1024,2048]."""

    def __init__(self, features=256, non_negative=True):
        super().__init__()
        self.scratch = make_scratch([256, 512, 1024, 2048], features, expand=False)
        for i in (4, 3, 2, 1):        # registration order of midas_net.py:38-41 (state_dict key order)
            setattr(self.scratch, f"refinenet{i}", FusionBlockLarge(features))
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, 128, 3, 1, 1),
            Upsample2x(2, "bilinear"),
            nn.Conv2d(128, 32, 3, 1, 1),
            nn.ReLU(True),
            nn.Conv2d(32, 1, 1),
            nn.ReLU(True) if non_negative else nn.Identity(),
        )

    def forward(self, l1, l2, l3, l4):
        s = self.scratch
        p4 = s.refinenet4(s.layer4_rn(l4))
        p3 = s.refinenet3(p4, s.layer3_rn(l3))
        p2 = s.refinenet2(p3, s.layer2_rn(l2))
        p1 = s.refinenet1(p2, s.layer1_rn(l1))
        return s.output_conv(p1).squeeze(dim=1)
