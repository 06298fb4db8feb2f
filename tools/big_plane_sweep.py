"""Plain AdaIN forward on 8 MiB planes (1x256x1024x2048): statistics->apply lag x merge mode.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
c, s = R.synth_features((1, 256, 1024, 2048), cfg=5, device="cuda")
E = c.numel() * 4
for min_spp in (512, 1 << 30):
    rpst.set_tuning("adain_group_merge_min_spp", min_spp)
    for lag in (24, 32, 48, 64, 96, 128, 256):
        rpst.set_tuning("adain_big_lag_bytes", lag << 20)
        for _ in range(2):
            rpst.adaptive_instance_normalization(c, s)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            rpst.adaptive_instance_normalization(c, s)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(json.dumps({"group_merge": min_spp == 512, "lag_MiB": lag, "ms": ms, "GBs": 3 * E / ms / 1e6}), flush=True)
