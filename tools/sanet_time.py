"""SANet(512) module forward (inference): fused projections + flash attention vs cuDNN convolutions + flash, vs eager torch."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
dev = torch.device("cuda")
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
torch.manual_seed(0)
m = rpst.SANet(512).to(dev)
for b, side in ((4, 64), (4, 128)):
    c, s = R.synth_features((b, 512, side, side), cfg=4, device=dev)
    res = {"op": f"SANet(512) module forward b={b} L={side * side}"}
    with torch.no_grad():
        for prec in ("fp32", "bf16"):
            m.precision = prec
            m.fused = True
            res[f"fused_{prec}_ms_per_sample"] = timeit(lambda: m(c, s)) / b
            m.fused = False
            res[f"cudnn_convs_{prec}_ms_per_sample"] = timeit(lambda: m(c, s)) / b
        m.fused = True

        def eager():
            mvn = lambda x: (x - x.mean((2, 3), keepdim=True)) / (x.var((2, 3), keepdim=True) + 1e-5).sqrt()
            F = m.f(mvn(c)).flatten(2).transpose(1, 2)
            G = m.g(mvn(s)).flatten(2)
            H = m.h(s).flatten(2)
            S = torch.softmax(torch.bmm(F, G), -1)
            return m.out_conv(torch.bmm(H, S.transpose(1, 2)).view_as(c)) + c
        if b * side * side <= 4 * 16384:
            res["eager_gpu_ms_per_sample"] = timeit(eager, 2, 1) / b
    print(json.dumps(res), flush=True)
# conv1x1 alone
for (cin, cout, side) in ((512, 512, 128), (256, 256, 512)):
    x = torch.randn(2, cin, side, side, device=dev)
    w = torch.randn(cout, cin, 1, 1, device=dev) / cin ** 0.5
    bias = torch.randn(cout, device=dev)
    conv = torch.nn.Conv2d(cin, cout, 1).to(dev)
    with torch.no_grad():
        t = timeit(lambda: rpst.conv1x1(x, w, bias)) / 2
        ts = timeit(lambda: rpst.conv1x1(x, w, bias, act="lrelu", want_stats=True)) / 2
        tc = timeit(lambda: conv(x)) / 2
    E = x[0].numel() * 4
    print(json.dumps({"op": f"conv1x1 {cin}->{cout} @ {side}x{side}", "rpst_ms": t, "rpst_lrelu_stats_ms": ts, "cudnn_ms": tc,
                      "GBs": (E + E * cout / cin) / t / 1e6}), flush=True)
