"""mean_variance_norm (rpst_adain_fwd with style == NULL) on 8x256x512x512: lag / merge-lead / twin-item sweep.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
x = torch.relu(torch.randn(8, 256, 512, 512, device="cuda") + 0.5)
E = x.numel() * 4
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
shape = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (8, 256, 512, 512)
x = torch.relu(torch.randn(*shape, device="cuda") + 0.5)
E = x.numel() * 4
for twin in (1,):
    rpst.set_tuning("adain_twin_apply", twin)
    for lag in (36, 40, 44, 48, 52, 56):
        for lead in (0, 12, 16, 32):
            rpst.set_tuning("adain_mvn_lag_bytes", lag << 20)
            rpst.set_tuning("adain_merge_lead", lead)      # 0 = lag / 2
            tm = t(lambda: rpst.mean_variance_norm(x))
            print(json.dumps({"twin": twin, "lag_MiB": lag, "lead": lead, "mvn_ms": round(tm, 4), "mvn_GBs": round(2 * E / tm / 1e6)}), flush=True)
