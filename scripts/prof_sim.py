"""Tiny driver for profiling the similarity tile kernel under ncu: one materialised 10 500 × 10 500 × d=100 pass."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openea_b200 import finding as F
g = torch.Generator().manual_seed(0)
e2 = torch.randn(10500, 100, generator=g).cuda()
e1 = (e2 + 0.5 * torch.randn(10500, 100, generator=g).cuda() * 0 + 0.5 * torch.randn(10500, 100, device="cuda"))
d1, d = F.to_device_rows(e1, False); d2, _ = F.to_device_rows(e2, False)
for _ in range(3):
    s = F.sim_matrix(d1, d2, d, "inner")
torch.cuda.synchronize()
print("ok", float(s[0, 0]))
