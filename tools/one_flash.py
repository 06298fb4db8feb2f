"""One SANet attention core call per precision at config #4's relu4_1 size (C=512, L=16384) — profiler target."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(4)
f = torch.randn(b, 512, side, side, device="cuda", generator=g) * 0.3
k = torch.randn(b, 512, side, side, device="cuda", generator=g) * 0.3
v = torch.randn(b, 512, side, side, device="cuda", generator=g)
for _ in range(2):
    for prec in ("fp32", "bf16"):
        out = rpst.attention_core(f, k, v, precision=prec)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
