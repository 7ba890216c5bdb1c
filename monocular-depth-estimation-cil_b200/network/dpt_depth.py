"""Drop-in for the in-scope parts of the reference's ``src/network/dpt_depth.py``: Dinov2Head (dpt_depth.py:32-153)
and the DPT / DPTDepthModel decoder + head (dpt_depth.py:155-293).  The timm transformer backbones are third-party
and unavailable offline; DPT takes the four reassembled feature maps through ``forward_features``."""
import torch
import torch.nn as nn

from .. import ops
from . import blocks as _blocks
from .base_model import BaseModel
from .blocks import FeatureFusionBlock_custom, Interpolate, _make_scratch, enter, leave, from_nchw


def _make_fusion_block(features, use_bn, size=None):
    return FeatureFusionBlock_custom(features, nn.ReLU(False), deconv=False, bn=use_bn, expand=False,
                                     align_corners=True, size=size)


class Dinov2Head(nn.Module):
    def __init__(self, nclass, in_channels, features=256, use_bn=False, out_channels=[256, 512, 1024, 1024],
                 use_clstoken=False):
        super(Dinov2Head, self).__init__()
        assert nclass == 1 and not use_clstoken, "the reference builds Dinov2Head(1, ..., use_clstoken=False) (midas_semantics.py:176)"
        self.nclass = nclass
        self.use_clstoken = use_clstoken
        self.projects = nn.ModuleList([nn.Conv2d(in_channels, oc, kernel_size=1, stride=1, padding=0) for oc in out_channels])
        self.resize_layers = nn.ModuleList([
            nn.ConvTranspose2d(out_channels[0], out_channels[0], kernel_size=4, stride=4, padding=0),
            nn.ConvTranspose2d(out_channels[1], out_channels[1], kernel_size=2, stride=2, padding=0),
            nn.Identity(),
            nn.Conv2d(out_channels[3], out_channels[3], kernel_size=3, stride=2, padding=1)])
        self.scratch = _make_scratch(out_channels, features, groups=1, expand=False)
        self.scratch.stem_transpose = None
        self.scratch.refinenet1 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet2 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet3 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet4 = _make_fusion_block(features, use_bn)
        head_features_1 = features
        head_features_2 = 32
        self.scratch.output_conv1 = nn.Conv2d(head_features_1, head_features_1 // 2, kernel_size=3, stride=1, padding=1)
        self.scratch.output_conv2 = nn.Sequential(
            nn.Conv2d(head_features_1 // 2, head_features_2, kernel_size=3, stride=1, padding=1),
            nn.ReLU(True),
            nn.Identity(),
        )

    def fused(self, out_features, patch_h, patch_w):
        """four (B, N, C) fp32 token tensors -> (B, 14*patch_h, 14*patch_w, 32) NHWC bf16."""
        maps = []
        for i, tok in enumerate(out_features):
            m = ops.tokens_to_nhwc(tok, patch_h, patch_w)
            m = ops.conv_tc(m, self.projects[i].weight, self.projects[i].bias)
            rl = self.resize_layers[i]
            if isinstance(rl, nn.ConvTranspose2d):
                m = ops.conv_transposed(m, rl.weight, rl.bias, rl.stride[0], rl.padding[0])
            elif isinstance(rl, nn.Conv2d):
                m = ops.conv_strided(m, rl.weight, rl.bias, rl.stride[0], rl.padding[0])
            maps.append(m)
        s = self.scratch
        rn = [ops.conv_tc(maps[i], getattr(s, f"layer{i + 1}_rn").weight, None, dual=True) for i in range(4)]
        p4 = s.refinenet4.fused(rn[3], None, size=rn[2][0].shape[1:3])
        p3 = s.refinenet3.fused(p4, rn[2], size=rn[1][0].shape[1:3])
        p2 = s.refinenet2.fused(p3, rn[1], size=rn[0][0].shape[1:3])
        p1 = s.refinenet1.fused(p2, rn[0])
        out = ops.conv_tc(p1, s.output_conv1.weight, s.output_conv1.bias)
        out = ops.resize(out, (int(patch_h * 14), int(patch_w * 14)), True)
        oc2 = s.output_conv2[0]
        return ops.conv_tc(out, oc2.weight, oc2.bias, relu=True)

    def forward(self, out_features, patch_h, patch_w):
        return ops.to_nchw(self.fused(out_features, patch_h, patch_w))


class DPT(BaseModel):
    """reference dpt_depth.py:155-266: decoder over four reassembled transformer feature maps."""

    _IN_SHAPES = {
        "beitl16_512": [256, 512, 1024, 1024], "beitl16_384": [256, 512, 1024, 1024], "beitb16_384": [96, 192, 384, 768],
        "swin2l24_384": [192, 384, 768, 1536], "swin2b24_384": [128, 256, 512, 1024], "swin2t16_256": [96, 192, 384, 768],
        "swinl12_384": [192, 384, 768, 1536], "vitl16_384": [256, 512, 1024, 1024], "vitb_rn50_384": [256, 512, 768, 768],
        "vitb16_384": [96, 192, 384, 768],
    }

    def __init__(self, head, features=256, backbone="vitb_rn50_384", readout="project", channels_last=False,
                 use_bn=False, **kwargs):
        super(DPT, self).__init__()
        assert not use_bn
        self.channels_last = channels_last
        if backbone not in self._IN_SHAPES:
            raise NotImplementedError(f"backbone {backbone}: only the 4-level DPT decoders are on the B200 path")
        # the timm backbone itself is third-party (absent offline): installed through blocks.hub_load when available
        self.pretrained = None
        if _blocks.hub_load is not None:
            try:
                self.pretrained = _blocks._hub("timm", backbone, readout=readout)
            except Exception:
                self.pretrained = None
        self.scratch = _make_scratch(self._IN_SHAPES[backbone], features, groups=1, expand=False)
        self.number_layers = 4
        self.scratch.stem_transpose = None
        self.scratch.refinenet1 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet2 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet3 = _make_fusion_block(features, use_bn)
        self.scratch.refinenet4 = _make_fusion_block(features, use_bn)
        self.scratch.output_conv = head

    def forward_features(self, layer_1, layer_2, layer_3, layer_4):
        """the decoder + head from the four reassembled maps (NCHW fp32 in, (B,H,W) fp32 out)."""
        s = self.scratch
        feats = [layer_1, layer_2, layer_3, layer_4]
        rn = [ops.conv_tc(from_nchw(f), getattr(s, f"layer{i + 1}_rn").weight, None, dual=True) for i, f in enumerate(feats)]
        p4 = s.refinenet4.fused(rn[3], None, size=rn[2][0].shape[1:3])
        p3 = s.refinenet3.fused(p4, rn[2], size=rn[1][0].shape[1:3])
        p2 = s.refinenet2.fused(p3, rn[1], size=rn[0][0].shape[1:3])
        p1 = s.refinenet1.fused(p2, rn[0])
        oc = s.output_conv
        a = ops.conv_tc(p1, oc[0].weight, oc[0].bias)
        B, H, W, _ = a.shape
        b = ops.resize(a, (int(H * oc[1].scale_factor), int(W * oc[1].scale_factor)), oc[1].align_corners)
        c = ops.conv_tc(b, oc[2].weight, oc[2].bias, relu=True)
        return ops.head_conv(c, oc[4].weight, oc[4].bias, isinstance(oc[5], nn.ReLU))

    def forward(self, x):
        if self.pretrained is None:
            raise RuntimeError("DPT: the timm transformer backbone is third-party and not available offline; "
                               "feed the four reassembled feature maps to forward_features()")
        layers = self.pretrained(x)
        return self.forward_features(*layers).unsqueeze(1)


class DPTDepthModel(DPT):
    """reference dpt_depth.py:269-293."""

    def __init__(self, path=None, non_negative=True, **kwargs):
        features = kwargs["features"] if "features" in kwargs else 256
        head_features_1 = kwargs["head_features_1"] if "head_features_1" in kwargs else features
        head_features_2 = kwargs["head_features_2"] if "head_features_2" in kwargs else 32
        kwargs.pop("head_features_1", None)
        kwargs.pop("head_features_2", None)
        head = nn.Sequential(
            nn.Conv2d(head_features_1, head_features_1 // 2, kernel_size=3, stride=1, padding=1),
            Interpolate(scale_factor=2, mode="bilinear", align_corners=True),
            nn.Conv2d(head_features_1 // 2, head_features_2, kernel_size=3, stride=1, padding=1),
            nn.ReLU(True),
            nn.Conv2d(head_features_2, 1, kernel_size=1, stride=1, padding=0),
            nn.ReLU(True) if non_negative else nn.Identity(),
            nn.Identity(),
        )
        super().__init__(head, **kwargs)
        if path is not None:
            self.load(path)

    def forward(self, x):
        return super().forward(x).squeeze(dim=1)
