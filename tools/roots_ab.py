"""WCT matrix roots A/B: Jacobi only (knob wct_roots_ns 0) vs Newton-Schulz (1) vs Newton-Schulz with every matrix
flagged, i.e. the predicated Jacobi path (2).  Transforms must agree.  GPU box only."""
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
shapes = [(16, 256, 512, 512), (2, 128, 256, 256), (2, 64, 100, 36), (1, 200, 31, 5), (2, 512, 32, 32), (3, 16, 8, 8)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for shape in shapes:
    c, s = R.synth_features(shape, cfg=3, device="cuda")
    for method in ("closed-form", "original"):
        row = {"shape": shape, "method": method}
        tr = {}
        for knob in (0, 1, 2):
            rpst.set_tuning("wct_roots_ns", knob)
            out, tr[knob] = rpst.wct_fuse(c, s, method, return_transform=True)
            if knob < 2:
                row[f"ms_per_sample_ns{knob}"] = round(t(lambda: rpst.wct_fuse(c, s, method)) / shape[0], 4)
        row["T_rel_ns_vs_jacobi"] = float((tr[1] - tr[0]).norm() / tr[0].norm())
        row["T_rel_flagged_vs_jacobi"] = float((tr[2] - tr[0]).norm() / tr[0].norm())
        print(json.dumps(row), flush=True)
rpst.set_tuning("wct_roots_ns", 1)
