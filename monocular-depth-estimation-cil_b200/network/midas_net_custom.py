"""Drop-in for the reference's ``src/network/midas_net_custom.py`` (MidasNet_small, midas_net_custom.py:45-185).

EfficientNet-Lite3 encoder (third-party, run by PyTorch) -> layerN_rn reassemble convs -> four fusion blocks ->
output_conv head, all on the sm_100a kernels.  LocalBins / DGR (off in the reference config, config.yaml:30-31)
are not part of the hot path and raise if requested.
"""
import torch
import torch.nn as nn

from .. import ops
from .base_model import BaseModel
from . import encoder_fused
from .blocks import FeatureFusionBlock_custom, Interpolate, _make_encoder, enter, from_nchw


class MidasNet_small(BaseModel):
    def __init__(self, path=None, features=64, backbone="efficientnet_lite3", non_negative=True, exportable=True,
                 channels_last=False, align_corners=True, cfg=None, blocks={'expand': True}):
        print("Loading weights: ", path)
        super(MidasNet_small, self).__init__()
        use_pretrained = False if path else True
        self.use_lb = cfg.use_lb
        self.use_dgr = cfg.use_dgr
        if self.use_lb or self.use_dgr:
            raise NotImplementedError("LocalBins / DGR heads are outside the B200 hot path (config.yaml:30-31 keeps them off)")
        self.channels_last = channels_last
        self.blocks = blocks
        self.backbone = backbone
        self.groups = 1
        self.non_negative = non_negative
        self.expand = bool("expand" in self.blocks and self.blocks['expand'] == True)  # noqa: E712
        f = [features, features * 2, features * 4, features * 8] if self.expand else [features] * 4

        self.pretrained, self.scratch = _make_encoder(self.backbone, features, use_pretrained, groups=self.groups,
                                                      expand=self.expand, exportable=exportable)
        self.scratch.activation = nn.ReLU(False)
        act = self.scratch.activation
        self.scratch.refinenet4 = FeatureFusionBlock_custom(f[3], act, deconv=False, bn=False, expand=self.expand, align_corners=align_corners)
        self.scratch.refinenet3 = FeatureFusionBlock_custom(f[2], act, deconv=False, bn=False, expand=self.expand, align_corners=align_corners)
        self.scratch.refinenet2 = FeatureFusionBlock_custom(f[1], act, deconv=False, bn=False, expand=self.expand, align_corners=align_corners)
        self.scratch.refinenet1 = FeatureFusionBlock_custom(f[0], act, deconv=False, bn=False, align_corners=align_corners)
        self.scratch.output_conv = nn.Sequential(
            nn.Conv2d(features, features // 2, kernel_size=3, stride=1, padding=1, groups=self.groups),
            Interpolate(scale_factor=2, mode="bilinear"),
            nn.Conv2d(features // 2, 32, kernel_size=3, stride=1, padding=1),
            self.scratch.activation,
            nn.Conv2d(32, 1, kernel_size=1, stride=1, padding=0),
            nn.ReLU(True) if non_negative else nn.Identity(),
            nn.Identity(),
        )
        # run the third-party encoder under bf16 autocast + channels_last (bench); parity tests keep fp32
        self.encoder_autocast = False
        # run the EfficientNet-Lite3 trunk on the sm_100a kernels (network/encoder_fused.py).  This is the default on a
        # CUDA device; a trunk whose structure is not gen-efficientnet's raises instead of silently running on PyTorch.
        # Set it to False explicitly to keep a third-party trunk on PyTorch (the parity tests compare both).
        self.fused_encoder = True
        if path:
            self.load(path)

    # -- pieces shared with MidasNetSemantics ----------------------------------------------------------
    def encoder_features(self, x):
        self._feats_nhwc = False
        if self.fused_encoder and x.is_cuda:
            if not encoder_fused.supported(self.pretrained):
                raise NotImplementedError(
                    "fused_encoder=True but `pretrained` is not a gen-efficientnet EfficientNet-Lite3 trunk "
                    "(network/encoder_fused.py); set model.fused_encoder = False to run this trunk through PyTorch")
            self._feats_nhwc = True
            return tuple(encoder_fused.forward(self.pretrained, x))      # NHWC bf16 (internal layout)
        if self.encoder_autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                xc = x.contiguous(memory_format=torch.channels_last)
                l1 = self.pretrained.layer1(xc)
                l2 = self.pretrained.layer2(l1)
                l3 = self.pretrained.layer3(l2)
                l4 = self.pretrained.layer4(l3)
            return l1, l2, l3, l4
        l1 = self.pretrained.layer1(x)
        l2 = self.pretrained.layer2(l1)
        l3 = self.pretrained.layer3(l2)
        l4 = self.pretrained.layer4(l3)
        return l1, l2, l3, l4

    def decoder_trunk(self, feats):
        """four encoder maps (NCHW) -> path_1 (NHWC bf16, stride-2 resolution)."""
        s = self.scratch
        rn = []
        for i, f in enumerate(feats):
            conv = getattr(s, f"layer{i + 1}_rn")
            t = f if getattr(self, "_feats_nhwc", False) else from_nchw(f)   # the fused trunk hands over NHWC bf16
            rn.append(ops.conv_tc(t, conv.weight, None, dual=True))      # (raw, relu) pairs
        p4 = s.refinenet4.fused(rn[3], None)
        p3 = s.refinenet3.fused(p4, rn[2])
        p2 = s.refinenet2.fused(p3, rn[1])
        return s.refinenet1.fused(p2, rn[0])

    def head_features(self, path_1):
        """output_conv[0:4]: conv3x3 -> x2 bilinear (align_corners=False) -> conv3x3 -> ReLU; NHWC bf16, 32 ch."""
        oc = self.scratch.output_conv
        a = ops.conv_tc(path_1, oc[0].weight, oc[0].bias)
        B, H, W, _ = a.shape
        sf = oc[1].scale_factor
        b = ops.resize(a, (int(H * sf), int(W * sf)), oc[1].align_corners)
        return ops.conv_tc(b, oc[2].weight, oc[2].bias, relu=True)

    def forward(self, x):
        feats = self.encoder_features(x)
        c = self.head_features(self.decoder_trunk(feats))
        oc = self.scratch.output_conv
        return ops.head_conv(c, oc[4].weight, oc[4].bias, isinstance(oc[5], nn.ReLU))      # (B,H,W) fp32
