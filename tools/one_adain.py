"""Run the AdaIN C-ABI call a few times on one level (for ncu).  python tools/one_adain.py batch ch blend(0/1) iters"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
batch, ch, blend, iters = (int(a) for a in (sys.argv[1:5] + ["8", "256", "1", "4"][len(sys.argv) - 1:]))
H = W = 512
dev = torch.device("cuda")
c = torch.relu(torch.randn(batch, ch, H, W, device=dev) + 0.5)
s = torch.relu(torch.randn(batch, ch, H, W, device=dev) * 2 + 1)
p = torch.randn(batch, ch, H, W, device=dev) if blend else None
for _ in range(iters):
    out = rpst.adain_blend(p, c, s) if blend else rpst.adaptive_instance_normalization(c, s)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
