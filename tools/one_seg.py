import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
n, ch, h, w = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 1024, 2048
if len(sys.argv) > 2:
    rpst.set_tuning("seg_lag_bytes", int(sys.argv[2]) << 20)
c, s = R.synth_features((n, ch, h, w), cfg=5, device="cuda")
cl = R.synth_labels(n, h, w, seed=4000, device="cuda"); sl = R.synth_labels(n, h, w, seed=5000, device="cuda")
for _ in range(3):
    out = rpst.seg_adain_batch(c, s, cl, sl)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
