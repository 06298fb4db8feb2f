import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
shape = tuple(int(v) for v in sys.argv[1:5])
method = sys.argv[5] if len(sys.argv) > 5 else "closed-form"
c, s = R.synth_features(shape, cfg=3, device="cuda")
for _ in range(2):
    out = rpst.wct_fuse(c, s, method)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0, 0]))
