"""seg-AdaIN config #5 (1x256x1024x2048, 19 labels): statistics->apply lag sweep.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
n, ch, h, w = (int(a) for a in os.environ.get("SEG_SHAPE", "1,256,1024,2048").split(","))
c, s = R.synth_features((n, ch, h, w), cfg=5, device="cuda")
cl = R.synth_labels(n, h, w, seed=4000, device="cuda"); sl = R.synth_labels(n, h, w, seed=5000, device="cuda")
E = c.numel() * 4
alg = 3 * E + 2 * h * w
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for flush in (4,):
  for lag_mib in [int(a) for a in sys.argv[1:]] or [24, 32, 48, 64, 96, 128, 256]:
    rpst.set_tuning("seg_lag_bytes", lag_mib << 20)
    for _ in range(6):   # the first configuration of a fresh process needs a long warm-up (clock ramp, lazy kernel load)
        rpst.seg_adain_batch(c, s, cl, sl)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = rpst.seg_adain_batch(c, s, cl, sl)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({"groups": flush, "seg_lag_MiB": lag_mib, "ms": ms, "GBs": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / peak}), flush=True)
