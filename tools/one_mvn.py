"""One mean_variance_norm call on 8x256x512x512 (profiler target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
x = torch.relu(torch.randn(8, 256, 512, 512, device="cuda") + 0.5)
for _ in range(3):
    y = rpst.mean_variance_norm(x)
torch.cuda.synchronize()
print("ok", float(y[0, 0, 0, 0]))
