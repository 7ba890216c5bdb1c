"""EfficientNet-Lite3 trunk on the sm_100a kernels.

The reference takes the trunk from ``torch.hub`` (``rwightman/gen-efficientnet-pytorch: tf_efficientnet_lite3``,
reference src/network/blocks.py:166-173) and slices it into ``pretrained.layer1..4`` (blocks.py:176-186).  Nothing is
re-parameterised here: the hub model's own submodules stay the parameter containers (state_dict keys unchanged) and this
module only *executes* them - stem conv, depthwise-separable and inverted-residual blocks, BatchNorm (train / eval),
ReLU6 - as NHWC bf16 kernels:

  conv_stem 3x3/s2      -> tcgen05 stride-2 implicit GEMM (3 channels zero-padded to 8), BN partials from the epilogue
  conv_pw / conv_pwl    -> tcgen05 1x1 implicit GEMM, BN partials from the epilogue
  conv_dw k3/k5 s1/s2   -> csrc/depthwise.cu (register-window depthwise kernel, BN partials fused)
  bnN + ReLU6 (+skip)   -> dp_bn_finalize + dp_bn_apply (one pass), backward = reduce + apply

Blocks are recognised by the attribute names gen-efficientnet uses (conv_pw, bn1, act1, conv_dw, bn2, act2, conv_pwl, bn3,
has_residual / conv_dw, bn1, act1, conv_pw, bn2).  Anything else makes `supported()` return False and the caller keeps
running the trunk through PyTorch.
"""
import torch.nn as nn

from .. import ops


def _is_relu6(m):
    return isinstance(m, nn.ReLU6) or type(m).__name__ in ("ReLU6",)


def _conv_ok(c, depthwise=False):
    if not isinstance(c, nn.Conv2d) or c.bias is not None or c.dilation != (1, 1):
        return False
    k, s = c.kernel_size, c.stride
    if k[0] != k[1] or s[0] != s[1]:
        return False
    if depthwise:
        return c.groups == c.in_channels == c.out_channels and k[0] in (3, 5) and s[0] in (1, 2) and c.in_channels % 8 == 0
    return c.groups == 1 and k[0] == 1 and s[0] == 1 and c.in_channels % 8 == 0 and c.out_channels % 8 == 0


def _block_kind(b):
    names = set(dict(b.named_children()).keys())
    if {"conv_pw", "bn1", "act1", "conv_dw", "bn2", "act2", "conv_pwl", "bn3"} <= names:
        ok = (_conv_ok(b.conv_pw) and _conv_ok(b.conv_dw, True) and _conv_ok(b.conv_pwl) and _is_relu6(b.act1)
              and _is_relu6(b.act2) and not any(n.startswith("se") and not isinstance(getattr(b, n), nn.Identity) for n in names))
        return "ir" if ok else None
    if {"conv_dw", "bn1", "act1", "conv_pw", "bn2"} <= names:
        ok = _conv_ok(b.conv_dw, True) and _conv_ok(b.conv_pw) and _is_relu6(b.act1)
        return "ds" if ok else None
    return None


def _walk(seq):
    """flatten nn.Sequential nesting into a list of leaf blocks / stem modules"""
    out = []
    for m in seq:
        if isinstance(m, nn.Sequential):
            out.extend(_walk(m))
        else:
            out.append(m)
    return out


def supported(pretrained):
    try:
        l1 = _walk(pretrained.layer1)
    except Exception:
        return False
    if len(l1) < 3:
        return False
    stem, bn, act = l1[0], l1[1], l1[2]
    if not (isinstance(stem, nn.Conv2d) and stem.kernel_size == (3, 3) and stem.stride == (2, 2) and stem.padding == (1, 1)
            and stem.bias is None and stem.in_channels <= 8 and stem.out_channels % 8 == 0
            and isinstance(bn, nn.BatchNorm2d) and _is_relu6(act)):
        return False
    rest = l1[3:] + _walk(pretrained.layer2) + _walk(pretrained.layer3) + _walk(pretrained.layer4)
    return all(_block_kind(b) is not None for b in rest)


def _dw_geom(conv, Hi, Wi):
    """(stride, pad_top, pad_left, Ho, Wo); TF-'SAME' wrappers (gen-efficientnet Conv2dSame) pad dynamically."""
    k, s = conv.kernel_size[0], conv.stride[0]
    if "Same" in type(conv).__name__:
        Ho, Wo = -(-Hi // s), -(-Wi // s)
        ph = max((Ho - 1) * s + k - Hi, 0)
        pw = max((Wo - 1) * s + k - Wi, 0)
        return s, ph // 2, pw // 2, Ho, Wo
    p = conv.padding[0]
    return s, p, p, (Hi + 2 * p - k) // s + 1, (Wi + 2 * p - k) // s + 1


def _pw(x, conv, bn, relu, res=None):
    tr = bn.training
    r = ops.conv_tc(x, conv.weight, None, stats=tr)
    c, st = r if tr else (r, None)
    return ops.bn_act(bn, c, st, relu=relu, res=res)


def _dw(x, conv, bn):
    tr = bn.training
    _, Hi, Wi, _ = x.shape
    s, pt, pl, Ho, Wo = _dw_geom(conv, Hi, Wi)
    r = ops.dwconv(x, conv.weight, s, pt, pl, Ho, Wo, stats=tr)
    c, st = r if tr else (r, None)
    return ops.bn_act(bn, c, st, relu=2)


def run_block(b, x):
    kind = _block_kind(b)
    skip = x if getattr(b, "has_residual", False) else None
    if kind == "ir":
        y = _pw(x, b.conv_pw, b.bn1, 2)
        y = _dw(y, b.conv_dw, b.bn2)
        return _pw(y, b.conv_pwl, b.bn3, 0, res=skip)
    y = _dw(x, b.conv_dw, b.bn1)
    return _pw(y, b.conv_pw, b.bn2, 0, res=skip)


def forward(pretrained, x):
    """(B,3,H,W) fp32 image -> four NHWC bf16 feature maps (strides 4, 8, 16, 32)."""
    l1 = _walk(pretrained.layer1)
    stem, bn = l1[0], l1[1]
    tr = bn.training
    r = ops.stem_conv(x, stem.weight, stats=tr)
    c, st = r if tr else (r, None)
    y = ops.bn_act(bn, c, st, relu=2)
    feats = []
    for b in l1[3:]:
        y = run_block(b, y)
    feats.append(y)
    for layer in (pretrained.layer2, pretrained.layer3, pretrained.layer4):
        for b in _walk(layer):
            y = run_block(b, y)
        feats.append(y)
    return feats
