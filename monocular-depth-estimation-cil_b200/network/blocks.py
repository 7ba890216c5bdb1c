"""Drop-in for the reference's ``src/network/blocks.py`` (decoder building blocks).

Same class names, constructor arguments, attribute names and state_dict keys; nn.Conv2d / nn.BatchNorm2d
objects are kept purely as parameter containers (fp32, OIHW) and their arithmetic is executed by the
sm_100a kernels through ``ops``.  Modules accept either the reference's NCHW fp32 tensors (converted at the
boundary, results converted back) or the internal NHWC bf16 tensors used between fused kernels.
"""
import torch
import torch.nn as nn

from .. import ops

# torch.hub.load replacement hook: tests / bench install the offline stand-ins here (no network on the box).
hub_load = None


def _hub(repo, name, **kw):
    fn = hub_load if hub_load is not None else torch.hub.load
    return fn(repo, name, **kw)


def is_internal(x):
    """internal NHWC bf16 tensors carry an explicit tag set by `ops` (never inferred from the dtype)"""
    return ops.is_internal(x)


def enter(x):
    """-> (nhwc bf16 tensor, was_public).  Public tensors are the reference's NCHW maps (any float dtype)."""
    if is_internal(x):
        return x, False
    return ops.to_nhwc(x), True


def from_nchw(f):
    """encoder feature map (B,C,H,W), fp32 or (autocast) bf16 -> NHWC bf16.  A channels_last bf16 map is already
    NHWC in memory: its permuted view is used as is (autograd routes the gradient back through the view)."""
    if f.dtype == torch.bfloat16:
        t = f.permute(0, 2, 3, 1)
        return ops.mark_internal(t if t.is_contiguous() else t.contiguous())
    return ops.to_nhwc(f)


def leave(y, was_public):
    return ops.to_nchw(y) if was_public else y


def _make_scratch(in_shape, out_shape, groups=1, expand=False):
    """reference blocks.py:133-163"""
    assert groups == 1
    scratch = nn.Module()
    mult = [1, 2, 4, 8] if expand else [1, 1, 1, 1]
    for i, cin in enumerate(in_shape):
        setattr(scratch, f"layer{i + 1}_rn",
                nn.Conv2d(cin, out_shape * mult[i], kernel_size=3, stride=1, padding=1, bias=False, groups=groups))
    return scratch


def _make_efficientnet_backbone(effnet):
    """reference blocks.py:176-186"""
    pretrained = nn.Module()
    pretrained.layer1 = nn.Sequential(effnet.conv_stem, effnet.bn1, effnet.act1, *effnet.blocks[0:2])
    pretrained.layer2 = nn.Sequential(*effnet.blocks[2:3])
    pretrained.layer3 = nn.Sequential(*effnet.blocks[3:5])
    pretrained.layer4 = nn.Sequential(*effnet.blocks[5:9])
    return pretrained


def _make_pretrained_efficientnet_lite3(use_pretrained, exportable=False):
    """reference blocks.py:166-173"""
    effnet = _hub("rwightman/gen-efficientnet-pytorch", "tf_efficientnet_lite3", pretrained=use_pretrained,
                  exportable=exportable)
    return _make_efficientnet_backbone(effnet)


def _make_resnet_backbone(resnet):
    """reference blocks.py:189-199"""
    pretrained = nn.Module()
    pretrained.layer1 = nn.Sequential(resnet.conv1, resnet.bn1, resnet.relu, resnet.maxpool, resnet.layer1)
    pretrained.layer2 = resnet.layer2
    pretrained.layer3 = resnet.layer3
    pretrained.layer4 = resnet.layer4
    return pretrained


def _make_encoder(backbone, features, use_pretrained, groups=1, expand=False, exportable=True, hooks=None,
                  use_vit_only=False, use_readout="ignore", in_features=[96, 256, 512, 1024]):
    """reference blocks.py:32-130.  The timm transformer backbones are third-party and absent offline; their
    reassembled feature-map channel counts are kept so the decoder (`scratch`) is built identically."""
    table = {
        "beitl16_512": [256, 512, 1024, 1024], "beitl16_384": [256, 512, 1024, 1024], "beitb16_384": [96, 192, 384, 768],
        "swin2l24_384": [192, 384, 768, 1536], "swin2b24_384": [128, 256, 512, 1024], "swin2t16_256": [96, 192, 384, 768],
        "swinl12_384": [192, 384, 768, 1536], "next_vit_large_6m": list(in_features), "levit_384": [384, 512, 768],
        "vitl16_384": [256, 512, 1024, 1024], "vitb_rn50_384": [256, 512, 768, 768], "vitb16_384": [96, 192, 384, 768],
    }
    if backbone == "efficientnet_lite3":
        pretrained = _make_pretrained_efficientnet_lite3(use_pretrained, exportable=exportable)
        scratch = _make_scratch([32, 48, 136, 384], features, groups=groups, expand=expand)
    elif backbone == "resnext101_wsl":
        pretrained = _make_resnet_backbone(_hub("facebookresearch/WSL-Images", "resnext101_32x8d_wsl"))
        scratch = _make_scratch([256, 512, 1024, 2048], features, groups=groups, expand=expand)
    elif backbone in table:
        pretrained = _hub("timm", backbone, hooks=hooks, use_readout=use_readout)
        scratch = _make_scratch(table[backbone], features, groups=groups, expand=expand)
    else:
        print(f"Backbone '{backbone}' not implemented")
        assert False
    return pretrained, scratch


class Interpolate(nn.Module):
    """reference blocks.py:208-240 (align_corners defaults to False)."""

    def __init__(self, scale_factor, mode, align_corners=False):
        super().__init__()
        assert mode == "bilinear"
        self.scale_factor, self.mode, self.align_corners = scale_factor, mode, align_corners

    def forward(self, x):
        t, pub = enter(x)
        B, H, W, C = t.shape
        y = ops.resize(t, (int(H * self.scale_factor), int(W * self.scale_factor)), self.align_corners)
        return leave(y, pub)


class ResidualConvUnit_custom(nn.Module):
    """reference blocks.py:319-376: x + conv2(act(conv1(act(x))))  (bn=False in every use of the reference)."""

    def __init__(self, features, activation, bn):
        super().__init__()
        assert not bn, "the reference never enables bn in its fusion blocks (midas_net_custom.py:88-91)"
        self.bn = bn
        self.groups = 1
        self.conv1 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True, groups=1)
        self.conv2 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True, groups=1)
        self.activation = activation

    def fused(self, x_skip, x_act, res2=None, dual=False):
        """x_skip: tensor added back; x_act: relu(x) feeding conv1; res2: extra addend fused into conv2's epilogue;
        dual: also return relu(result)."""
        a = ops.conv_tc(x_act, self.conv1.weight, self.conv1.bias, relu=True)
        return ops.conv_tc(a, self.conv2.weight, self.conv2.bias, res=x_skip, res2=res2, dual=dual)

    def forward(self, x):
        t, pub = enter(x)
        return leave(self.fused(t, ops.relu(t)), pub)


class ResidualConvUnit(ResidualConvUnit_custom):
    """reference blocks.py:243-279: the first ReLU is in place, so the skip adds relu(x)."""

    def __init__(self, features):
        super().__init__(features, nn.ReLU(False), False)
        self.relu = nn.ReLU(inplace=True)
        del self.activation

    def forward(self, x):
        t, pub = enter(x)
        r = ops.relu(t)
        return leave(self.fused(r, r), pub)


class FeatureFusionBlock_custom(nn.Module):
    """reference blocks.py:379-438."""

    def __init__(self, features, activation, deconv=False, bn=False, expand=False, align_corners=True, size=None):
        super().__init__()
        self.deconv = deconv
        self.align_corners = align_corners
        self.groups = 1
        self.expand = expand
        out_features = features // 2 if expand else features
        self.out_conv = nn.Conv2d(features, out_features, kernel_size=1, stride=1, padding=0, bias=True, groups=1)
        self.resConfUnit1 = ResidualConvUnit_custom(features, activation, bn)
        self.resConfUnit2 = ResidualConvUnit_custom(features, activation, bn)
        self.size = size

    def fused(self, x0, x1_pair=None, size=None):
        """x0: NHWC tensor (two-input form) or (raw, relu) pair (single-input form); x1_pair = (raw, relu) of xs[1]."""
        if x1_pair is not None:
            y, y_act = self.resConfUnit1.fused(x1_pair[0], x1_pair[1], res2=x0, dual=True)   # xs[0] + RCU1(xs[1])
        else:
            y, y_act = x0
        z = self.resConfUnit2.fused(y, y_act)
        B, H, W, C = z.shape
        if size is None and self.size is None:
            target = (2 * H, 2 * W)
        else:
            target = self.size if size is None else size
            if isinstance(target, int):
                target = (target, target)
        # the 1x1 out_conv commutes with bilinear interpolation (both linear; interpolation weights sum to 1 so the
        # bias survives): run it at the low resolution (4x fewer MACs), then resize.
        p = ops.conv_tc(z, self.out_conv.weight, self.out_conv.bias)
        return ops.resize(p, target, self.align_corners)

    def forward(self, *xs, size=None):
        t0, pub = enter(xs[0])
        if len(xs) == 2:
            t1, _ = enter(xs[1])
            out = self.fused(t0, (t1, ops.relu(t1)), size=size)
        else:
            out = self.fused((t0, ops.relu(t0)), None, size=size)
        return leave(out, pub)


class FeatureFusionBlock(nn.Module):
    """reference blocks.py:282-314 (MiDaS v2.1 large): in-place-ReLU residual units, x2 upsample, no out_conv."""

    def __init__(self, features):
        super().__init__()
        self.resConfUnit1 = ResidualConvUnit(features)
        self.resConfUnit2 = ResidualConvUnit(features)

    def fused(self, x0, x1_act=None):
        """x0: NHWC tensor (two-input) or relu'd tensor (single input); x1_act = relu(xs[1])."""
        if x1_act is not None:
            _, y_act = self.resConfUnit1.fused(x1_act, x1_act, res2=x0, dual=True)
        else:
            y_act = x0
        z = self.resConfUnit2.fused(y_act, y_act)
        B, H, W, C = z.shape
        return ops.resize(z, (2 * H, 2 * W), True)

    def forward(self, *xs):
        t0, pub = enter(xs[0])
        if len(xs) == 2:
            t1, _ = enter(xs[1])
            out = self.fused(t0, ops.relu(t1))
        else:
            out = self.fused(ops.relu(t0), None)
        return leave(out, pub)
