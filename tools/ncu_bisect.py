import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench, depth_b200
from depth_b200 import config as fx
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
x, t = bench.synthetic_batch(2, 1234)
x, t = x.to(dev), t.to(dev)
out = model(x)
torch.cuda.synchronize()
print("fwd ok")
loss, _ = depth_b200.combined_loss(out.unsqueeze(1), t, fx.loss_config(), rgb=x)
loss.backward()
torch.cuda.synchronize()
print("bwd ok")
