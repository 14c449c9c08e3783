"""One FocalSplatLoss forward+gradient at the bench shape (32x80x128x128, 128 objects per image) -- for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_b200 import losses as L

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
B, C, H, W, M = 32, 80, 128, 128, 128
pred = torch.clamp(torch.sigmoid(torch.randn(B, C, H, W, device=dev, generator=g) - 2.19), 1e-4, 1 - 1e-4).requires_grad_(True)
cls = torch.randint(0, C, (B, M), device=dev, generator=g)
cx = torch.randint(0, W, (B, M), device=dev, generator=g)
cy = torch.randint(0, H, (B, M), device=dev, generator=g)
rad = torch.randint(0, 12, (B, M), device=dev, generator=g)
val = torch.ones(B, M, dtype=torch.uint8, device=dev)
crit = L.FocalSplatLoss()
for _ in range(3):
    crit(pred, cls, cx, cy, rad, val)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    crit(pred, cls, cx, cy, rad, val)
e1.record()
torch.cuda.synchronize()
print("FocalSplatLoss fwd+grad: %.4f ms (CNB_FOCAL_PARTS=%s)" % (e0.elapsed_time(e1) / 20, os.environ.get("CNB_FOCAL_PARTS", "default")))
