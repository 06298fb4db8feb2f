"""Per-op measurements for the other BASELINE configs (one JSON line each): rpst kernels vs the
reference's eager op sequence on the same GPU.  GPU box only.
    python tools/bench_ops.py [adain1 seg wct sanet mrf bwd]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rpst  # noqa: E402
from oracle import restate as R  # noqa: E402  (bench tool: eager-GPU restatements for comparison)

dev = torch.device("cuda")
PEAK_GBS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
PEAK_TF = 1668.9


def timeit(fn, iters=10, warm=3):
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


def eager_stats(x):
    n, c = x.shape[:2]
    f = x.reshape(n, c, -1)
    return f.mean(2).view(n, c, 1, 1), (f.var(2) + 1e-5).sqrt().view(n, c, 1, 1)


def eager_adain(c, s):
    ms, ss = eager_stats(s)
    mc, sc = eager_stats(c)
    return (c - mc) / sc * ss + ms


def emit(**kw):
    print(json.dumps(kw), flush=True)


def bench_adain1():
    c, s = R.synth_features((1, 512, 64, 64), cfg=1, device=dev)
    t = timeit(lambda: rpst.adaptive_instance_normalization(c, s), 50)
    te = timeit(lambda: eager_adain(c, s), 50)
    emit(op="adain config#1 1x512x64x64", rpst_us=t * 1e3, eager_gpu_us=te * 1e3, speedup=te / t,
         note="launch-latency bound: 25 MB algorithmic = 3.9 us at the HBM peak")


def bench_bwd():
    shape = (8, 256, 512, 512)
    c, s = R.synth_features(shape, cfg=2, device=dev)
    c.requires_grad_(); s.requires_grad_()
    g = torch.randn(shape, device=dev)
    out = rpst.adaptive_instance_normalization(c, s)
    t = timeit(lambda: torch.autograd.grad(out, (c, s), g, retain_graph=True), 5)
    E = c.numel() * 4
    emit(op="adain backward 8x256x512x512", ms=t, GBs=5 * E / t / 1e6, frac_of_peak=5 * E / t / 1e6 / PEAK_GBS)


def bench_stats():
    """calc_mean_std / mean_variance_norm alone (E and 2E algorithmic bytes)."""
    for shape in ((8, 256, 512, 512), (1, 256, 1024, 2048)):
        x, _ = R.synth_features(shape, cfg=2, device=dev)
        E = x.numel() * 4
        t = timeit(lambda: rpst.calc_mean_std(x), 10)
        tm = timeit(lambda: rpst.mean_variance_norm(x), 10)
        te = timeit(lambda: eager_stats(x), 5)
        emit(op=f"calc_mean_std {shape}", ms=t, GBs=E / t / 1e6, frac_of_peak=E / t / 1e6 / PEAK_GBS, eager_gpu_ms=te,
             mvn_ms=tm, mvn_GBs=2 * E / tm / 1e6)


def bench_train5():
    """config #5 training transform: AdaIN forward + backward on a Cityscapes-sized level (1x256x1024x2048)."""
    shape = (1, 256, 1024, 2048)
    c, s = R.synth_features(shape, cfg=5, device=dev)
    c.requires_grad_(); s.requires_grad_()
    g = torch.randn(shape, device=dev)

    def step():
        out = rpst.adaptive_instance_normalization(c, s)
        return torch.autograd.grad(out, (c, s), g)
    t = timeit(step, 5)
    E = c.numel() * 4
    out = rpst.adaptive_instance_normalization(c, s)
    tb = timeit(lambda: torch.autograd.grad(out, (c, s), g, retain_graph=True), 5)
    with torch.no_grad():
        tf = timeit(lambda: rpst.adaptive_instance_normalization(c, s), 5)
    emit(op="adain fwd+bwd config#5 level 1x256x1024x2048", ms=t, GBs=8 * E / t / 1e6, frac_of_peak=8 * E / t / 1e6 / PEAK_GBS,
         fwd_ms=tf, fwd_GBs=3 * E / tf / 1e6, bwd_ms=tb, bwd_GBs=5 * E / tb / 1e6,
         note="3E forward + 5E backward algorithmic bytes")


def bench_seg():
    n, ch, h, w = 1, 256, 1024, 2048
    c, s = R.synth_features((n, ch, h, w), cfg=5, device=dev)
    cl = R.synth_labels(n, h, w, seed=4000, device=dev)
    sl = R.synth_labels(n, h, w, seed=5000, device=dev)
    t = timeit(lambda: rpst.seg_adain_batch(c, s, cl, sl), 5)
    E = c.numel() * 4
    emit(op="seg-adain config#5 1x256x1024x2048, 19 labels", ms=t, GBs=(3 * E + 2 * h * w) / t / 1e6,
         frac_of_peak=(3 * E + 2 * h * w) / t / 1e6 / PEAK_GBS)


def bench_wct():
    n, ch, h, w = 16, 256, 512, 512      # BASELINE configs[2] at its full batch
    c, s = R.synth_features((n, ch, h, w), cfg=3, device=dev)
    t = timeit(lambda: rpst.wct_fuse(c, s), 3, 1)
    t16 = timeit(lambda: rpst.wct_fuse(c, s, precision="bf16"), 3, 1)

    def msqrt(a, p):
        a = a.clone(); a.diagonal().add_(1e-4)
        _, e, vh = torch.linalg.svd(a)
        v = vh.t()
        return (v * e.pow(p)) @ v.t()

    def eager_one():   # the reference's fp64 eager op sequence for ONE sample (network/wct_rp.py:82-114,157-166)
        cf = c[0].reshape(ch, -1).double(); sf = s[0].reshape(ch, -1).double()
        cm = cf.mean(1, keepdim=True); xc = cf - cm
        cc = xc @ xc.t() / (xc.shape[1] - 1) + torch.eye(ch, dtype=torch.float64, device=dev)
        sm = sf.mean(1, keepdim=True); xs = sf - sm
        cs = xs @ xs.t() / (xs.shape[1] - 1)
        r, ir = msqrt(cc, 0.5), msqrt(cc, -0.5)
        t_ = ir @ msqrt(r @ cs @ r, 0.5) @ ir
        return (t_ @ xc + sm).float()
    te = timeit(eager_one, 1, 1)
    flops = 3 * 2 * ch * ch * h * w * n
    emit(op="wct config#3 batch 16 x 256x512x512", ms_per_sample=t / n, bf16_ms_per_sample=t16 / n,
         eager_gpu_fp64_ms_per_sample=te, speedup_vs_eager=te / (t / n), algorithmic_TFLOPs=flops / t / 1e9,
         algorithmic_GBs=4 * c.numel() * 4 / t / 1e6)


def bench_sanet():
    for (b, l_side) in ((2, 64), (1, 128)):
        ch = 512
        g = torch.Generator(device=dev).manual_seed(4)
        f = torch.randn(b, ch, l_side, l_side, device=dev, generator=g) * 0.3
        k = torch.randn(b, ch, l_side, l_side, device=dev, generator=g) * 0.3
        v = torch.randn(b, ch, l_side, l_side, device=dev, generator=g)
        L = l_side * l_side
        t3 = timeit(lambda: rpst.attention_core(f, k, v), 3, 1)
        t1 = timeit(lambda: rpst.attention_core(f, k, v, precision="bf16"), 3, 1)

        def eager():
            F = f.reshape(b, ch, -1); G = k.reshape(b, ch, -1); H = v.reshape(b, ch, -1)
            S = torch.softmax(torch.bmm(F.transpose(1, 2), G), -1)
            return torch.bmm(H, S.transpose(1, 2))
        te = timeit(eager, 3, 1)
        flops = 4 * L * L * ch * b
        emit(op=f"sanet attention core config#4 b={b} L={L} C=512", fp32grade_ms=t3, bf16_ms=t1, eager_gpu_fp32_ms=te,
             algorithmic_TFLOPs_fp32grade=flops / t3 / 1e9, algorithmic_TFLOPs_bf16=flops / t1 / 1e9,
             speedup_fp32grade=te / t3, speedup_bf16=te / t1)


def bench_sanet_bwd():
    """attention core forward+backward at the SAModel training sizes (network/sanet.py:253-276 at 512^2
    crops: relu4_1 L=4096, relu5_1 L=1024) vs eager autograd on the same GPU."""
    for (b, l_side) in ((4, 64), (4, 32)):
        ch = 512
        g = torch.Generator(device=dev).manual_seed(4)
        f, k, v = (torch.randn(b, ch, l_side, l_side, device=dev, generator=g) * sc for sc in (0.3, 0.3, 1.0))
        w = torch.randn(b, ch, l_side, l_side, device=dev, generator=g)
        L = l_side * l_side
        for t in (f, k, v):
            t.requires_grad_()

        def ours(prec):
            out = rpst.attention_core(f, k, v, precision=prec)
            return torch.autograd.grad(out, (f, k, v), w)

        def eager():
            F = f.reshape(b, ch, -1); G = k.reshape(b, ch, -1); H = v.reshape(b, ch, -1)
            S = torch.softmax(torch.bmm(F.transpose(1, 2), G), -1)
            return torch.autograd.grad(torch.bmm(H, S.transpose(1, 2)).reshape(w.shape), (f, k, v), w)
        t3, t1, te = timeit(lambda: ours("fp32"), 3, 1), timeit(lambda: ours("bf16"), 3, 1), timeit(eager, 3, 1)
        flops = (2 + 6) * 2 * L * L * ch * b     # forward 2 GEMMs, backward recompute + 5 GEMMs (one shared): 8 in total
        emit(op=f"sanet attention fwd+bwd b={b} L={L} C=512", fp32grade_ms=t3, bf16_ms=t1, eager_gpu_fp32_ms=te,
             TFLOPs_fp32grade=flops / t3 / 1e9, TFLOPs_bf16=flops / t1 / 1e9, speedup_fp32grade=te / t3, speedup_bf16=te / t1)


def bench_adaptive():
    """AdaptiveSANet attention (config #4, 'aea' clamp) forward at relu5_1 / relu4_1 sizes of a 1024^2 image:
    whole module (convs in cuDNN) vs the reference's op sequence in eager torch on the same GPU."""
    for l_side in (64, 128):
        ch, L = 512, l_side * l_side
        torch.manual_seed(0)
        m = rpst.AdaptiveSANet(ch, L, ada_module="aea").to(dev)
        m.keep_claims = False
        c, s = R.synth_features((1, ch, l_side, l_side), cfg=4, device=dev)

        def eager():
            al = m.attention_layer
            mvn = lambda x: (x - x.mean((2, 3), keepdim=True)) / (x.var((2, 3), keepdim=True) + 1e-5).sqrt()
            F = m.f(mvn(c)).view(1, ch, L).permute(0, 2, 1)
            G = m.g(mvn(s)).view(1, ch, L)
            H = m.h(s).view(1, ch, L)
            aff = torch.bmm(torch.nn.functional.normalize(c.view(1, ch, L), dim=1).permute(0, 2, 1),
                            torch.nn.functional.normalize(s.view(1, ch, L), dim=1))
            S = torch.softmax(torch.bmm(F, G), -1)
            clamp = (al.f_psi(aff.view(L, L)) * al.value_interval + al.from_value).view(1, L, 1)
            S = torch.sigmoid(al.scale_value * (S - clamp))
            return m.out_conv(torch.bmm(H, S.permute(0, 2, 1)).view(1, ch, l_side, l_side)) + c
        with torch.no_grad():
            t = timeit(lambda: m(c, s), 3, 1)
            te = timeit(eager, 3, 1)
            err = float((m(c, s) - eager()).norm() / eager().norm())
        flops = (2 * L * L * ch * 3 + 2 * L * L * (L // 16))
        emit(op=f"adaptive sanet ('aea') module forward config#4 L={L} C=512", rpst_ms=t, eager_gpu_ms=te, speedup=te / t,
             algorithmic_TFLOPs=flops / t / 1e9, rel_l2_vs_eager_tf32_default=err)


def bench_mrf():
    ch, side, k = 512, 64, 5
    c, s = R.synth_features((1, ch, side, side), cfg=6, device=dev)
    t = timeit(lambda: rpst.mrf_match(c, s, k, want_loss=True), 5)

    def eager():
        a = c.view(ch, -1); b = s.view(ch, -1)
        an = torch.nn.functional.normalize(a, dim=0); bn = torch.nn.functional.normalize(b, dim=0)
        m = an.t() @ bn
        aff = torch.zeros_like(m)
        aff.scatter_(0, torch.topk(m, k, 0)[1], 1.0); aff.scatter_(1, torch.topk(m, k, 1)[1], 1.0)
        d = (a * a).sum(0)[:, None] + (b * b).sum(0)[None] - 2 * a.t() @ b
        return (aff * d).sum() / (side * side * k)
    te = timeit(eager, 5)
    emit(op="mrf match+loss C=512 L=4096 k=5", rpst_ms=t, eager_gpu_ms=te, speedup=te / t)


def bench_losses():
    """calc_style_loss / calc_content_loss(norm) on the training step's VGG shapes (batch 8 @ 512^2 crops
    -> relu1_1 64x512^2) and on an RP-level tensor; eager = the reference's op sequence on the same GPU."""
    for shape in ((8, 64, 512, 512), (8, 512, 64, 64)):
        x, y = R.synth_features(shape, cfg=8, device=dev)
        E = x.numel() * 4
        t = timeit(lambda: rpst.calc_style_loss(x, y), 10)

        def eager_style():
            mi, si = eager_stats(x); mt, st = eager_stats(y)
            return torch.nn.functional.mse_loss(mi, mt) + torch.nn.functional.mse_loss(si, st)

        def eager_content():
            mi, si = eager_stats(x); mt, st = eager_stats(y)
            return torch.nn.functional.mse_loss((x - mi) / si, (y - mt) / st)
        te, tc = timeit(eager_style, 5), timeit(eager_content, 5)
        xg = x.clone().requires_grad_()
        loss = rpst.calc_content_loss(xg, y, norm=True)
        tb = timeit(lambda: torch.autograd.grad(loss, xg, retain_graph=True), 5)
        emit(op=f"loss statistics {shape}", one_pass_ms=t, GBs=2 * E / t / 1e6, frac_of_peak=2 * E / t / 1e6 / PEAK_GBS,
             eager_style_ms=te, eager_content_norm_ms=tc, speedup_style=te / t, speedup_content_norm=tc / t,
             content_norm_bwd_ms=tb, bwd_GBs=3 * E / tb / 1e6)


ALL = {"losses": bench_losses, "sanet_bwd": bench_sanet_bwd, "adain1": bench_adain1, "bwd": bench_bwd, "stats": bench_stats, "train5": bench_train5, "seg": bench_seg, "wct": bench_wct, "sanet": bench_sanet, "adaptive": bench_adaptive, "mrf": bench_mrf}
for name in (sys.argv[1:] or list(ALL)):
    try:
        ALL[name]()
    except Exception as e:  # keep going: one op failing must not hide the others
        emit(op=name, error=repr(e)[:300])
