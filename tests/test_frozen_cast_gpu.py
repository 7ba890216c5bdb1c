"""bf16 weight copies of a frozen encoder's nn.Linear layers under autocast (network/frozen_cast.py): the forward is
bit-identical to stock autocast, the copies follow in-place parameter updates, and the state_dict does not change."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _encoder():
    torch.manual_seed(3)
    m = nn.Sequential(nn.Linear(64, 128), nn.LayerNorm(128), nn.GELU(), nn.Linear(128, 48, bias=False)).cuda().eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def test_cached_casts_match_stock_autocast_and_follow_updates(pkg):
    from depth_b200.network import frozen_cast
    ref = _encoder()
    m = copy.deepcopy(ref)
    keys = list(m.state_dict().keys())
    frozen_cast.enable(m)
    frozen_cast.enable(m)                                   # idempotent
    x = torch.randn(7, 64, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = ref(x), m(x)
    assert b.dtype == torch.bfloat16 and torch.equal(a, b)
    assert list(m.state_dict().keys()) == keys
    assert torch.equal(m(x), ref(x))                        # outside autocast: the plain fp32 path
    # in-place parameter update (load_state_dict does exactly this): the copy is refreshed in its own storage
    w16 = m[0].__dict__["_dp_cast"][1]
    ptr = w16.data_ptr()
    with torch.no_grad():
        for mod in (ref, m):
            mod[0].weight.mul_(1.5)
    frozen_cast.refresh_all()
    assert m[0].__dict__["_dp_cast"][1].data_ptr() == ptr
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(ref(x), m(x))
    # a trainable layer is left alone
    t = nn.Linear(8, 8).cuda()
    frozen_cast.enable(t)
    assert "forward" not in t.__dict__
