"""WCT timing breakdown: batched eigensolver alone and the whole fuse at config #3 (GPU box only)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
dev = torch.device("cuda")
def timeit(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for n, batch in ((256, 1), (256, 16), (64, 16), (128, 16), (512, 4)):
    g = torch.Generator(device=dev).manual_seed(n)
    a = torch.randn(batch, n, 2 * n, device=dev, dtype=torch.float64, generator=g)
    a = a @ a.transpose(1, 2) / (2 * n)
    t = timeit(lambda: rpst.matrix_sqrt(a))
    r = rpst.matrix_sqrt(a)
    err = float(((r @ r) - (a + 1e-4 * torch.eye(n, device=dev, dtype=torch.float64))).norm() / a.norm())
    print(json.dumps({"op": "matrix_sqrt", "n": n, "batch": batch, "ms": t, "residual": err}), flush=True)
c, s = R.synth_features((16, 256, 512, 512), cfg=3, device=dev)
t = timeit(lambda: rpst.wct_fuse(c, s))
print(json.dumps({"op": "wct_fuse 16x256x512x512 fp32-grade", "ms_per_sample": t / 16}), flush=True)
t = timeit(lambda: rpst.wct_fuse(c, s, precision="bf16"))
print(json.dumps({"op": "wct_fuse 16x256x512x512 bf16", "ms_per_sample": t / 16}), flush=True)
