"""diagnostic: where does the 448x576 forward drift come from?"""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import depth_b200 as pkg
from depth_b200 import ops
from oracle import fixtures as fx
from tests.test_benched_config_gpu import _pair, _batch, rel_l2, rel_max

ora, prod = _pair(pkg)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x, t = _batch(B)
with torch.no_grad():
    ref = ora(x)
    for fused in (False, True):
        prod.fused_encoder = fused
        out = prod(x)
        print(f"fused_encoder={fused}: max {rel_max(out, ref):.4f} l2 {rel_l2(out, ref):.4f}")
    auto = copy.deepcopy(ora)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        oa = auto(x).float()
    print(f"stock autocast oracle: max {rel_max(oa, ref):.4f} l2 {rel_l2(oa, ref):.4f}")
    # trunk features
    prod.fused_encoder = True
    ff = prod.encoder_features(x)
    prod.fused_encoder = False
    ft = prod.encoder_features(x)
    for i, (a, b) in enumerate(zip(ff, ft)):
        a = a.permute(0, 3, 1, 2).float()
        print(f"  trunk map {i}: fused vs torch fp32 max {rel_max(a, b):.4f} l2 {rel_l2(a, b):.4f}")
    # decoder stages on identical (torch fp32) features: product vs oracle
    fo = ora.encoder_features(x) if hasattr(ora, "encoder_features") else None
    print("oracle has encoder_features:", fo is not None)
    print("ref stats: mean %.3f std %.3f max %.3f" % (float(ref.mean()), float(ref.std()), float(ref.max())))
