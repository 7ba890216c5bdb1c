"""Sweep of the bilinear resize forward / backward kernels over small and odd shapes, both corner conventions, dense and
channel-slice inputs, against F.interpolate.  `sweep()` returns the cases whose max error exceeds 2 % of the reference
maximum (bf16 rounding is ~0.4 %); tests/test_ops_gpu.py runs it, `python tools/resize_sweep.py` prints them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sweep(channels=(8, 32, 64, 256, 512)):
    import torch
    import torch.nn.functional as F
    from depth_b200 import ops
    torch.manual_seed(0)
    bad = []
    for al in (True, False):
        for (Hi, Wi) in [(2, 3), (4, 6), (8, 12), (16, 24), (32, 48), (3, 5), (7, 9), (16, 20)]:
            for C in channels:
                for strided in (False, True):
                    B = 2
                    Ho, Wo = 2 * Hi, 2 * Wi
                    x = torch.randn(B, C, Hi, Wi).to(torch.bfloat16).float()
                    xr = x.clone().requires_grad_(True)
                    ref = F.interpolate(xr, size=(Ho, Wo), mode="bilinear", align_corners=al)
                    cot = torch.randn(B, C, Ho, Wo).to(torch.bfloat16).float()
                    ref.backward(cot)
                    if strided:      # the operand is a channel slice of a wider NHWC buffer (pixel stride 2C)
                        wide = torch.zeros(B, Hi, Wi, 2 * C, dtype=torch.bfloat16, device="cuda")
                        wide[..., C:] = x.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
                        wide.requires_grad_(True)
                        xp = wide[..., C:]
                    else:
                        wide = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda().requires_grad_(True)
                        xp = wide
                    out = ops.resize(xp, (Ho, Wo), al)
                    out.backward(cot.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda())
                    o = out.detach().float().permute(0, 3, 1, 2).cpu()
                    g = (wide.grad[..., C:] if strided else wide.grad).float().permute(0, 3, 1, 2).cpu()
                    ef = float((o - ref.detach()).abs().max() / ref.detach().abs().max())
                    eb = float((g - xr.grad).abs().max() / xr.grad.abs().max())
                    if ef > 0.02 or eb > 0.02:
                        bad.append((al, Hi, Wi, C, strided, round(ef, 4), round(eb, 4)))
    return bad


if __name__ == "__main__":
    import depth_b200  # noqa: F401  (fails loudly when the library is not built)
    cases = sweep()
    for c in cases:
        print("BAD align=%s %dx%d C=%d strided=%s fwd %.4f bwd %.4f" % c)
    print("bad", len(cases))
