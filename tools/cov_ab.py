"""WCT covariance kernels A/B: register-staged vs TMA-staged (knob wct_cov_tma); outputs must agree.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for shape in ((4, 256, 512, 512), (2, 128, 256, 256), (2, 64, 100, 36), (1, 200, 75, 76)):
    c, s = R.synth_features(shape, cfg=3, device="cuda")
    row = {"shape": shape}
    outs = {}
    for prec in ("fp32", "bf16"):
        for knob in (0, 1):
            rpst.set_tuning("wct_cov_tma", knob)
            outs[knob] = rpst.wct_fuse(c, s, precision=prec)
            row[f"ms_per_sample_{prec}_tma{knob}"] = round(t(lambda: rpst.wct_fuse(c, s, precision=prec)) / shape[0], 4)
        row[f"rel_l2_{prec}"] = float((outs[0] - outs[1]).norm() / outs[0].norm())
    print(json.dumps(row), flush=True)
rpst.set_tuning("wct_cov_tma", 1)
