"""Segment AdaIN: agreement with the oracle on a few shapes, then timing at BASELINE configs[4]
(1x256x1024x2048, 19 classes) and on mid-size planes.  Usage: python tools/seg_time.py [--quick]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst  # noqa: E402
from oracle import restate as R  # noqa: E402


def dev_ms(fn, iters=10, warm=3):
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


def check(shape_c, shape_s, classes, block, with_prev):
    n = shape_c[0]
    c = torch.relu(torch.randn(shape_c, generator=torch.Generator().manual_seed(1)) + 0.5)
    s = torch.relu(torch.randn(shape_s, generator=torch.Generator().manual_seed(2)) * 2 + 1)
    cl = R.synth_labels(n, shape_c[2], shape_c[3], classes=classes, block=block, seed=4000)
    sl = R.synth_labels(n, shape_s[2], shape_s[3], classes=classes, block=block, seed=5000)
    cl[:, :2, :3] = 255
    prev = torch.randn(shape_c) if with_prev else None
    want = R.seg_adain_batch(c, s, cl, sl, dtype=torch.float64)
    if with_prev:
        want = want + prev.double()
    got = rpst.seg_adain_batch(c.cuda(), s.cuda(), cl.cuda(), sl.cuda(), prev=prev.cuda() if with_prev else None)
    torch.cuda.synchronize()
    err = R.rel_l2(got, want)
    print(json.dumps({"shape_c": shape_c, "shape_s": shape_s, "classes": classes, "block": block, "prev": with_prev, "err": err}),
          flush=True)
    assert err < 5e-6, err


def main():
    quick = "--quick" in sys.argv
    for args in (((1, 3, 256, 512), (1, 3, 256, 512), 6, 16, False),
                 ((2, 2, 256, 512), (2, 2, 256, 512), 6, 32, True),
                 ((2, 3, 304, 400), (2, 3, 208, 336), 19, 8, False),
                 ((1, 2, 512, 1024), (1, 2, 256, 512), 40, 4, False),
                 ((1, 2, 512, 512), (1, 2, 512, 512), 200, 4, True),
                 ((3, 1, 1024, 2048), (3, 1, 1024, 2048), 19, 32, False)):
        check(*args)
    if quick:
        return
    n, ch, h, w = 1, 256, 1024, 2048
    c, s = R.synth_features((n, ch, h, w), cfg=5, device="cuda")
    cl = R.synth_labels(n, h, w, seed=4000, device="cuda")
    sl = R.synth_labels(n, h, w, seed=5000, device="cuda")
    E = c.numel() * 4
    alg = 3 * E + 2 * h * w
    for _ in range(2):
        with torch.no_grad():
            ms = dev_ms(lambda: rpst.seg_adain_batch(c, s, cl, sl))
        print(json.dumps({"ms": ms, "GBs_3E": alg / ms / 1e6}), flush=True)
    prev = torch.randn_like(c)
    with torch.no_grad():
        ms = dev_ms(lambda: rpst.seg_adain_batch(c, s, cl, sl, prev=prev))
    print(json.dumps({"prev": True, "ms": ms, "GBs_4E": (alg + E) / ms / 1e6}), flush=True)
    del prev
    for (hh, ww, cc) in ((256, 512, 512), (512, 1024, 256)):
        c, s = R.synth_features((1, cc, hh, ww), cfg=5, device="cuda")
        cl = R.synth_labels(1, hh, ww, seed=4000, device="cuda")
        sl = R.synth_labels(1, hh, ww, seed=5000, device="cuda")
        with torch.no_grad():
            ms = dev_ms(lambda: rpst.seg_adain_batch(c, s, cl, sl))
        print(json.dumps({"plane": [hh, ww], "channels": cc, "ms": ms, "GBs_3E": 3 * c.numel() * 4 / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
