import os, sys, torch, json
sys.path.insert(0, "/root/repo")
import rpst
from oracle import restate as R
def t(fn, it=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for shape in ((3, 16, 8, 8), (1, 32, 16, 16), (2, 64, 100, 36)):
    c, s = R.synth_features(shape, cfg=3, device="cuda")
    row = {"shape": shape}
    for knob in ("wct_cov_tma", "pw_x_tma", "wct_fused_cov"):
        for v in (0, 1):
            rpst.set_tuning(knob, v)
            row[f"{knob}{v}"] = round(t(lambda: rpst.wct_fuse(c, s)), 4)
        rpst.set_tuning(knob, 1)
    print(json.dumps(row), flush=True)
