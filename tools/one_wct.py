import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c, s = R.synth_features((n, 256, 512, 512), cfg=3, device="cuda")
for _ in range(2):
    out = rpst.wct_fuse(c, s)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
