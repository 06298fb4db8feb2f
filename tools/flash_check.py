"""Flash attention bring-up: smallest case first, then timing at config #4 sizes (GPU box only)."""
import sys, os, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R

dev = torch.device("cuda")


def check(b, lc_hw, ls_hw, prec, scale=0.3, seed=0):
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(b, 512, *lc_hw, generator=g) * scale
    k = torch.randn(b, 512, *ls_hw, generator=g) * scale
    v = torch.randn(b, 512, *ls_hw, generator=g)
    want = R.attention_core(f.reshape(b, 512, -1).double(), k.reshape(b, 512, -1).double(), v.reshape(b, 512, -1).double())
    got = rpst.attention_core(f.cuda(), k.cuda(), v.cuda(), precision=prec)
    torch.cuda.synchronize()
    e = R.rel_l2(got.reshape(b, 512, -1), want)
    print(json.dumps({"check": [b, lc_hw, ls_hw, prec], "rel_l2": e}), flush=True)
    return e


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


if "quick" in sys.argv or len(sys.argv) == 1:
    for prec in ("bf16", "fp32"):
        check(1, (8, 16), (16, 16), prec)
        check(1, (8, 16), (32, 32), prec)
        check(2, (16, 16), (16, 32), prec)
if "time" in sys.argv or len(sys.argv) == 1:
    for (b, side) in ((4, 64), (4, 128), (16, 128)):
        g = torch.Generator(device=dev).manual_seed(4)
        f = torch.randn(b, 512, side, side, device=dev, generator=g) * 0.3
        k = torch.randn(b, 512, side, side, device=dev, generator=g) * 0.3
        v = torch.randn(b, 512, side, side, device=dev, generator=g)
        L = side * side
        res = {"b": b, "L": L}
        for flash in (1, 0):
            rpst.set_tuning("attn_flash", flash)
            if flash == 0 and b * L > 4 * 16384:
                continue
            for prec in ("bf16", "fp32"):
                t = timeit(lambda: rpst.attention_core(f, k, v, precision=prec), 3, 1)
                res[f"{'flash' if flash else 'gemm3'}_{prec}_ms_per_sample"] = t / b
                res[f"{'flash' if flash else 'gemm3'}_{prec}_TFLOPs"] = 4 * L * L * 512 * b / t / 1e9
        rpst.set_tuning("attn_flash", 1)
        print(json.dumps(res), flush=True)
