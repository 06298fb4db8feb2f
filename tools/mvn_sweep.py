"""adain_tma_kernel on 512x512 planes: statistics->apply lag sweep (mvn / plain / blend).  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
x = torch.relu(torch.randn(8, 256, 512, 512, device="cuda") + 0.5)
s = torch.relu(torch.randn(8, 256, 512, 512, device="cuda") * 2 + 1)
E = x.numel() * 4
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for lag, lead in [(32, 0), (32, 8), (28, 0), (24, 0), (24, 8), (20, 0), (16, 0), (40, 0)]:
    rpst.set_tuning("adain_lag_bytes", lag << 20)
    rpst.set_tuning("adain_merge_lead", lead)      # 0 = lag / 2
    tm = t(lambda: rpst.mean_variance_norm(x))
    ta = t(lambda: rpst.adaptive_instance_normalization(x, s))
    tb = t(lambda: rpst.adain_blend(s, x, s))
    print(json.dumps({"lag_MiB": lag, "lead": lead, "mvn_GBs": round(2 * E / tm / 1e6), "adain_GBs": round(3 * E / ta / 1e6),
                      "blend_GBs": round(4 * E / tb / 1e6)}), flush=True)
