import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_b200
from oracle import cases, fixtures as fx
from tests.test_modules_gpu import build_product, round_weights_bf16_, run, rel_err, rel_l2
names = sys.argv[1:] or ["resblock_64_64", "fusion128_expand", "xattn_1win", "dinohead"]
for name in names:
    kind, kw, shapes, fkw = cases.CASES[name]
    ora = fx.fill_deterministic(cases.build_oracle(kind, kw))
    prod = fx.fill_deterministic(build_product(depth_b200, kind, kw))
    round_weights_bf16_(ora)
    prod.load_state_dict(ora.state_dict(), strict=True)
    r_o = run(ora, name, "cpu"); r_p = run(prod, name, "cuda")
    print("==", name)
    for k, v in r_o.items():
        if k in r_p:
            print(f"   {k:50s} max {rel_err(r_p[k], v):.4f}  l2 {rel_l2(r_p[k], v):.4f}")
        else:
            print(f"   {k:50s} MISSING")
