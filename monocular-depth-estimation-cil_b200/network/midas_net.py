"""Drop-in for the reference's ``src/network/midas_net.py`` (MiDaS v2.1 large decoder, midas_net.py:12-76).
The ResNeXt101-WSL encoder is a torch.hub model (third-party); the decoder + head run on the sm_100a kernels."""
import torch
import torch.nn as nn

from .. import ops
from . import blocks as _blocks
from .base_model import BaseModel
from .blocks import FeatureFusionBlock, Interpolate, _make_scratch, _make_resnet_backbone, enter, from_nchw


class MidasNet(BaseModel):
    def __init__(self, path=None, features=256, non_negative=True):
        print("Loading weights: ", path)
        super(MidasNet, self).__init__()
        self.pretrained = None
        try:
            self.pretrained = _make_resnet_backbone(_blocks._hub("facebookresearch/WSL-Images", "resnext101_32x8d_wsl"))
        except Exception:
            self.pretrained = None      # offline: decoder-only use through forward_features()
        self.scratch = _make_scratch([256, 512, 1024, 2048], features, groups=1, expand=False)
        self.scratch.refinenet4 = FeatureFusionBlock(features)
        self.scratch.refinenet3 = FeatureFusionBlock(features)
        self.scratch.refinenet2 = FeatureFusionBlock(features)
        self.scratch.refinenet1 = FeatureFusionBlock(features)
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, 128, kernel_size=3, stride=1, padding=1),
            Interpolate(scale_factor=2, mode="bilinear"),
            nn.Conv2d(128, 32, kernel_size=3, stride=1, padding=1),
            nn.ReLU(True),
            nn.Conv2d(32, 1, kernel_size=1, stride=1, padding=0),
            nn.ReLU(True) if non_negative else nn.Identity(),
        )
        if path:
            self.load(path)

    def forward_features(self, layer_1, layer_2, layer_3, layer_4):
        s = self.scratch
        feats = [layer_1, layer_2, layer_3, layer_4]
        rn = [ops.conv_tc(from_nchw(f), getattr(s, f"layer{i + 1}_rn").weight, None, dual=True) for i, f in enumerate(feats)]
        p4 = s.refinenet4.fused(rn[3][1], None)                 # single input: relu(x) feeds both conv1 and the skip
        p3 = s.refinenet3.fused(p4, rn[2][1])
        p2 = s.refinenet2.fused(p3, rn[1][1])
        p1 = s.refinenet1.fused(p2, rn[0][1])
        oc = s.output_conv
        a = ops.conv_tc(p1, oc[0].weight, oc[0].bias)
        B, H, W, _ = a.shape
        b = ops.resize(a, (int(H * oc[1].scale_factor), int(W * oc[1].scale_factor)), oc[1].align_corners)
        c = ops.conv_tc(b, oc[2].weight, oc[2].bias, relu=True)
        return ops.head_conv(c, oc[4].weight, oc[4].bias, isinstance(oc[5], nn.ReLU))

    def forward(self, x):
        if self.pretrained is None:
            raise RuntimeError("MidasNet: the ResNeXt101-WSL encoder is a torch.hub model that is not available offline; "
                               "feed four feature maps to forward_features()")
        l1 = self.pretrained.layer1(x)
        l2 = self.pretrained.layer2(l1)
        l3 = self.pretrained.layer3(l2)
        l4 = self.pretrained.layer4(l3)
        return self.forward_features(l1, l2, l3, l4)
