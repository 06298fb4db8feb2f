import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
c, s = R.synth_features((1, 512, 64, 64), cfg=6, device="cuda")
for _ in range(3):
    out = rpst.mrf_match(c, s, 5, want_loss=True)
torch.cuda.synchronize()
print("ok")
