import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
n, ch, h, w = 1, 256, 1024, 2048
c, s = R.synth_features((n, ch, h, w), cfg=5, device="cuda")
cl = R.synth_labels(n, h, w, seed=4000, device="cuda"); sl = R.synth_labels(n, h, w, seed=5000, device="cuda")
E = c.numel() * 4
for lag in (24, 32, 48, 64, 96, 128, 256):
    rpst.set_tuning("seg_lag_bytes", lag << 20)
    for _ in range(2): rpst.seg_adain_batch(c, s, cl, sl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): rpst.seg_adain_batch(c, s, cl, sl)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5
    print(json.dumps({"seg_lag_mb": lag, "ms": t, "GBs": 3 * E / t / 1e6}), flush=True)
