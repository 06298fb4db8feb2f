import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ch, L = 512, side * side
torch.manual_seed(0)
m = rpst.AdaptiveSANet(ch, L, ada_module="aea").cuda()
m.keep_claims = False
c, s = R.synth_features((1, ch, side, side), cfg=4, device="cuda")
with torch.no_grad():
    for _ in range(2):
        out = m(c, s)
torch.cuda.synchronize()
print("ok")
