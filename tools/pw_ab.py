"""Pointwise-convolution kernel A/B: register-staged converters (knob pw_x_tma 0) vs TMA-staged x tiles (1) on the WCT
colouring, the SANet module and plain conv1x1 shapes; outputs must be bit-identical (same arithmetic).  GPU box only."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst
from oracle import restate as R
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
c, s = R.synth_features((8, 256, 512, 512), cfg=3, device="cuda")
row = {"op": "wct_fuse 8x256x512x512"}
outs = {}
for knob in (0, 1):
    rpst.set_tuning("pw_x_tma", knob)
    outs[knob] = rpst.wct_fuse(c, s)
    row[f"ms_per_sample_tma{knob}"] = round(t(lambda: rpst.wct_fuse(c, s), 3) / 8, 4)
row["identical"] = bool(torch.equal(outs[0], outs[1]))
print(json.dumps(row), flush=True)
del c, s, outs
for (b, cin, cout, side) in ((1, 256, 256, 512), (2, 512, 512, 128), (1, 64, 128, 256)):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(b, cin, side, side, device="cuda", generator=g)
    w = torch.randn(cout, cin, device="cuda", generator=g) / cin ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    row = {"op": f"conv1x1 {cin}->{cout} @ {side}x{side} b={b}"}
    outs = {}
    for knob in (0, 1):
        rpst.set_tuning("pw_x_tma", knob)
        outs[knob] = rpst.conv1x1(x, w, bias)
        row[f"ms_tma{knob}"] = round(t(lambda: rpst.conv1x1(x, w, bias)), 4)
    row["identical"] = bool(torch.equal(outs[0], outs[1]))
    row["rel_l2_vs_torch_fp64"] = float((outs[1].double() - (torch.einsum("oc,bchw->bohw", w.double(), x.double()) + bias.double().view(1, -1, 1, 1))).norm() / outs[1].double().norm())
    print(json.dumps(row), flush=True)
rpst.set_tuning("pw_x_tma", 1)
