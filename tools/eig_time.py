"""Batched 256x256 matrix square roots: 16-column blocks (8-CTA clusters) vs 32-column blocks (4-CTA clusters), with the
cluster occupancy the driver reports.  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
n = 256
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for batch in (1, 8, 12, 16, 32):
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(batch, n, 2 * n, device="cuda", dtype=torch.float64, generator=g)
    a = a @ a.transpose(1, 2) / (2 * n)
    row = {"batch": batch}
    for wide, name in ((0, "bc16"), (2, "bc16_one_cta_per_sm"), (1, "bc32")):
        rpst.set_tuning("eig_wide", wide)
        row["ms_" + name] = round(t(lambda: rpst.matrix_sqrt(a), 2 if wide == 2 else 5), 3)
    print(json.dumps(row), flush=True)
rpst.set_tuning("eig_wide", -2)
a = torch.randn(16, n, 2 * n, device="cuda", dtype=torch.float64)
a = a @ a.transpose(1, 2) / (2 * n)
rpst.matrix_sqrt(a)
torch.cuda.synchronize()
rpst.set_tuning("eig_wide", -1)
