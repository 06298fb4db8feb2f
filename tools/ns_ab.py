"""fp64 GEMM of the Newton-Schulz iteration: DFMA kernel (knob ns_dmma 0) vs fp64 tensor-core kernel (1).  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for (b, n, lmin) in ((16, 256, 1e-4), (16, 256, 1.0), (4, 512, 1e-4), (8, 128, 1e-4), (3, 200, 1e-4)):
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(b, n, 3 * n, device="cuda", dtype=torch.float64, generator=g)
    a = a @ a.transpose(1, 2) / (3 * n) + (lmin if lmin >= 1 else 0.0) * torch.eye(n, device="cuda", dtype=torch.float64)
    row = {"batch": b, "n": n, "lmin": lmin}
    res = {}
    for knob in (0, 1):
        rpst.set_tuning("ns_dmma", knob)
        res[knob] = rpst.spd_roots(a, lmin=lmin)
        row[f"ms_dmma{knob}"] = round(t(lambda: rpst.spd_roots(a, lmin=lmin)), 4)
    row["root_rel_diff"] = float((res[0][0] - res[1][0]).norm() / res[0][0].norm())
    row["iroot_rel_diff"] = float((res[0][1] - res[1][1]).norm() / res[0][1].norm())
    row["flags"] = [int(res[0][2].sum()), int(res[1][2].sum())]
    eye = torch.eye(n, device="cuda", dtype=torch.float64)
    row["residual_dmma1"] = float((res[1][0] @ res[1][0] - (a + 1e-4 * eye)).norm() / a.norm())
    print(json.dumps(row), flush=True)
rpst.set_tuning("ns_dmma", 1)
