"""Time stamps (%globaltimer, ns) inside cov_tma_kernel for CTA 0 and the last CTA: start, prologue done, main loop done,
epilogue done (barrier entry), barrier passed, S done, finalize done.  GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
c, s = R.synth_features((1, 256, 512, 512), cfg=3, device="cuda")
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
rpst.wct_fuse(c, s)
rpst.set_tuning("wct_cov_prof", buf.data_ptr())
rpst.wct_fuse(c, s)
torch.cuda.synchronize()
rpst.set_tuning("wct_cov_prof", 0)
b = buf.cpu().tolist()
for name, off in (("cta0", 0), ("last", 8)):
    t = b[off:off + 7]
    print(name, [round((x - t[0]) / 1e3, 2) for x in t])
