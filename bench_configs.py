"""Per-config measurements for bench.py (BASELINE.json configs #1, #3, #4, #5; configs[1] is the headline in
bench.py itself).  Every entry: device time of the rpst call (CUDA events on the launch stream), the algorithmic
bytes / flops of SURVEY.md §8(d), the roofline fraction against MEASURED_PEAKS.json, the tensor-pipe figure from the
tracked ncu summary (profiles/r02_tensor_pipe.json) and a CPU baseline of the same op (oracle port on the host cores,
on a bounded slice scaled linearly — samples / rows / channels are independent on every path).

Also: the PCIe roofline of the `e2e` leg, NUMA-local pinning, and the `e2e_images` leg (images in host memory ->
stub RP encoder (cuDNN) -> rpst decode loop -> image back), which is SURVEY §8(d)'s end-to-end `*.test()` shape
(network/adain_rp.py:251-269)."""
from __future__ import annotations

import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def tensor_pipe_summary():
    p = os.path.join(ROOT, "profiles", "r02_tensor_pipe.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


def dev_time_ms(fn, iters, warm=3):
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


def cpu_time_s(fn, repeats=2):
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


# ------------------------------------------------------------------------------------------------- config #1
def config1(dev, peaks, cpu=True):
    """AdaIN-RP inference, one 512x512 pair at VGG relu4_1: 1x512x64x64 (25.2 MB algorithmic = 3.9 us of HBM time:
    launch-latency bound, reported in us).  Inputs rotate over 16 buffer sets (400 MB > L2)."""
    import rpst
    from oracle import restate as R
    sets = [R.synth_features((1, 512, 64, 64), cfg=100 + i, device=dev) for i in range(16)]
    it = [0]

    def call():
        c, s = sets[it[0] & 15]
        it[0] += 1
        return rpst.adaptive_instance_normalization(c, s)
    with torch.no_grad():
        ms = dev_time_ms(call, 320, 32)
        # the same call captured in a CUDA graph (what a serving loop would replay): no host launch path
        c0, s0 = sets[0]
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            rpst.adaptive_instance_normalization(c0, s0)
            with torch.cuda.graph(g, stream=side):
                outs = [rpst.adaptive_instance_normalization(*sets[i]) for i in range(16)]
        torch.cuda.current_stream().wait_stream(side)
        graph_ms = dev_time_ms(g.replay, 20, 3) / 16
    alg = 3 * 512 * 64 * 64 * 4
    out = {"workload": "configs[0]: AdaIN @VGG relu4_1, 1x512x64x64, single 512x512 pair", "us_per_call": ms * 1e3,
           "us_per_call_graph": graph_ms * 1e3, "algorithmic_bytes": alg,
           "roofline": {"bound": "hbm (launch-latency bound at this size)", "achieved": alg / (graph_ms / 1e3) / 1e9,
                        "achieved_eager_launch": alg / (ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": alg / (graph_ms / 1e3) / 1e9 / peaks["hbm_gbs"]},
           "l2": "16 rotating input sets (400 MB) so no call re-reads L2-resident data"}
    if cpu:
        c, s = R.synth_features((1, 512, 64, 64), cfg=100)
        t = cpu_time_s(lambda: R.adain_aten(c, s), 5)
        out["cpu_baseline"] = {"value": t * 1e6, "unit": "us/call", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "the full op (1x512x64x64), best of 5, reference aten op sequence"}
    return out


# ------------------------------------------------------------------------------------------------- config #3
def config3(dev, peaks, cpu=True, batch=16):
    """deeper RP-WCT, batch 16 @512^2: WCTRPNet.fuse on 16x256x512x512 (network/wct_rp.py:157-166)."""
    import rpst
    from oracle import restate as R
    n, ch, h, w = batch, 256, 512, 512
    c, s = R.synth_features((n, ch, h, w), cfg=3, device=dev)
    with torch.no_grad():
        ms32 = dev_time_ms(lambda: rpst.wct_fuse(c, s), 3, 1)
        ms16 = dev_time_ms(lambda: rpst.wct_fuse(c, s, precision="bf16"), 3, 1)
    flops = 3 * 2 * ch * ch * h * w * n            # two covariances + apply counted as full GEMMs (103.1 GFLOP / sample)
    byts = 4 * ch * h * w * 4 * n                  # read c twice, s once, write once
    tp = tensor_pipe_summary().get("wct", {})
    out = {"workload": "configs[2]: WCT whitening/colouring 16x256x512x512 (closed form, fp64 Newton-Schulz matrix roots on device)",
           "ms_per_sample_fp32grade": ms32 / n, "ms_per_sample_bf16": ms16 / n, "flops_per_sample": flops // n,
           "bytes_per_sample": byts // n,
           "roofline": {"bound": "mixed: tensor + hbm (covariance / apply) + fp64 pipe (matrix roots)",
                        "achieved": flops / (ms32 / 1e3) / 1e12, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": flops / (ms32 / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
                        "achieved_bf16": flops / (ms16 / 1e3) / 1e12,
                        "hbm_achieved_GBs": byts / (ms32 / 1e3) / 1e9, "hbm_frac": byts / (ms32 / 1e3) / 1e9 / peaks["hbm_gbs"],
                        "tensor_pipe": tp or None,
                        "note": "flops = algorithmic fp32 GEMM count; the fp32-grade mode issues 3 bf16 MMA passes per product"},
           "l2": "inputs 4 GiB per tensor, far beyond L2"}
    del c, s
    torch.cuda.empty_cache()
    if cpu:
        c1, s1 = R.synth_features((1, ch, h, w), cfg=3)
        t = cpu_time_s(lambda: R.wct_fuse(c1, s1), 1)
        out["cpu_baseline"] = {"value": t * 1e3, "unit": "ms/sample", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1 of 16 samples (samples are independent), fp64 like network/wct_rp.py:161"}
    return out


# ------------------------------------------------------------------------------------------------- config #4
def config4(dev, peaks, cpu=True, batch=16):
    """SANet attention, batch 16 @1024^2: relu4_1 16x512x128x128 (L=16384) and relu5_1 16x512x64x64 (L=4096)."""
    import rpst
    from oracle import restate as R
    out = {"workload": "configs[3]: SANet attention core softmax(F^T G) H, C=512, batch 16: L=16384 (relu4_1 @1024^2) and L=4096 (relu5_1)"}
    tp = tensor_pipe_summary().get("attention", {})
    total_flops, total_ms = 0.0, {"fp32": 0.0, "bf16": 0.0}
    for side in (128, 64):
        L = side * side
        g = torch.Generator(device=dev).manual_seed(4)
        f = torch.randn(batch, 512, side, side, device=dev, generator=g) * 0.3
        k = torch.randn(batch, 512, side, side, device=dev, generator=g) * 0.3
        v = torch.randn(batch, 512, side, side, device=dev, generator=g)
        flops = 4.0 * L * L * 512 * batch
        total_flops += flops
        for prec in ("fp32", "bf16"):
            with torch.no_grad():
                ms = dev_time_ms(lambda: rpst.attention_core(f, k, v, precision=prec), 3, 1)
            total_ms[prec] += ms
            out[f"L{L}_{prec}"] = {"ms_per_sample": ms / batch, "TFLOPs": flops / (ms / 1e3) / 1e12,
                                   "frac_of_bf16_sustained": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"]}
        del f, k, v
    out["roofline"] = {"bound": "tensor", "achieved": total_flops / (total_ms["fp32"] / 1e3) / 1e12,
                       "achieved_bf16": total_flops / (total_ms["bf16"] / 1e3) / 1e12,
                       "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                       "frac": total_flops / (total_ms["fp32"] / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
                       "frac_bf16": total_flops / (total_ms["bf16"] / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
                       "tensor_pipe": tp or None,
                       "note": "algorithmic flops 4 L^2 C per sample; fp32-grade issues 3 bf16 passes on the logits + 1 half pass on P.V"}
    # the whole SANet(512) module at relu4_1 (instance normalisation + four 1x1 projections + attention + residual):
    # projections on the tcgen05 block with the normalisation folded in (rpst.conv) vs cuDNN convolutions around the core
    try:
        torch.manual_seed(0)
        m = rpst.SANet(512).to(dev)
        c, s = R.synth_features((4, 512, 128, 128), cfg=4, device=dev)
        mod = {}
        with torch.no_grad():
            for prec in ("fp32", "bf16"):
                m.precision = prec
                m.fused = True
                mod[f"fused_{prec}_ms_per_sample"] = dev_time_ms(lambda: m(c, s), 3, 1) / 4
                m.fused = False
                mod[f"cudnn_projections_{prec}_ms_per_sample"] = dev_time_ms(lambda: m(c, s), 3, 1) / 4
        out["module_L16384"] = mod
        del m, c, s
    except Exception as e:
        out["module_L16384"] = {"error": repr(e)[:200]}
    # the adaptive variant of the same config (network/sanet.py:100-138, 'aea' clamp), one sample at relu4_1, and the MRF
    # match + loss of network/mrf_rp.py at relu4_1 of a 512^2 image (top-k selected in the NCC GEMM epilogue)
    try:
        torch.manual_seed(0)
        ma = rpst.AdaptiveSANet(512, 16384, ada_module="aea").to(dev)
        ma.keep_claims = False
        c, s = R.synth_features((1, 512, 128, 128), cfg=4, device=dev)
        with torch.no_grad():
            ms_ada = dev_time_ms(lambda: ma(c, s), 3, 1)
        L = 16384
        out["adaptive_module_L16384"] = {"ms_per_sample": ms_ada, "algorithmic_TFLOPs": (2.0 * L * L * 512 * 3 + 2.0 * L * L * (L // 16)) / ms_ada / 1e9,
                                         "note": "affinity + f_psi Linear(L -> L/16) + logits + P.V, all fp32-grade (bf16x3)"}
        del ma, c, s
        torch.cuda.empty_cache()
        c, s = R.synth_features((1, 512, 64, 64), cfg=6, device=dev)
        with torch.no_grad():
            ms_mrf = dev_time_ms(lambda: rpst.mrf_match(c, s, 5, want_loss=True), 5, 2)
        out["mrf_match_loss_L4096_k5"] = {"ms": ms_mrf, "note": "C=512; NCC and NCC^T products never stored, indices bit-exact"}
        del c, s
    except Exception as e:
        out["adaptive_module_L16384"] = {"error": repr(e)[:200]}
    out["l2"] = "operands of one call: 3 x 512 MiB (L=16384) / 3 x 128 MiB (L=4096), beyond L2"
    torch.cuda.empty_cache()
    if cpu:
        # 1024 of the 16384 query rows of one sample (rows are independent), scaled x16; both levels
        def one(L, rows):
            g = torch.Generator().manual_seed(4)
            f = torch.randn(1, 512, rows, generator=g) * 0.3
            k = torch.randn(1, 512, L, generator=g) * 0.3
            v = torch.randn(1, 512, L, generator=g)
            return cpu_time_s(lambda: R.attention_core(f, k, v), 2) * (L / rows)
        t = one(16384, 1024) + one(4096, 1024)
        out["cpu_baseline"] = {"value": t * 1e3, "unit": "ms/sample (both levels)", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1024 query rows of one sample per level, scaled by L/1024 (rows and samples independent), fp32 bmm + softmax"}
    return out


# ------------------------------------------------------------------------------------------------- config #5
def config5(dev, peaks, cpu=True):
    """seg-AdaIN-RP on Cityscapes-shaped inputs: 1x256x1024x2048 + uint8 label maps, 19 classes (forward), and the
    training transform of the same level (AdaIN forward + backward)."""
    import rpst
    from oracle import restate as R
    n, ch, h, w = 1, 256, 1024, 2048
    c, s = R.synth_features((n, ch, h, w), cfg=5, device=dev)
    cl = R.synth_labels(n, h, w, seed=4000, device=dev)
    sl = R.synth_labels(n, h, w, seed=5000, device=dev)
    E = c.numel() * 4
    with torch.no_grad():
        ms = dev_time_ms(lambda: rpst.seg_adain_batch(c, s, cl, sl), 10, 3)
    alg = 3 * E + 2 * h * w
    cg, sg = c.clone().requires_grad_(), s.clone().requires_grad_()
    gout = torch.randn_like(c)

    def train_step():
        o = rpst.adaptive_instance_normalization(cg, sg)
        return torch.autograd.grad(o, (cg, sg), gout)
    ms_train = dev_time_ms(train_step, 5, 2)
    out = {"workload": "configs[4]: seg-AdaIN 1x256x1024x2048 + uint8 labels (19 classes), forward; plus AdaIN fwd+bwd of that level",
           "ms": ms, "algorithmic_bytes": alg,
           "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": alg / (ms / 1e3) / 1e9 / peaks["hbm_gbs"]},
           "train_fwd_bwd": {"ms": ms_train, "algorithmic_bytes": 8 * E, "GBs": 8 * E / (ms_train / 1e3) / 1e9,
                             "frac": 8 * E / (ms_train / 1e3) / 1e9 / peaks["hbm_gbs"]},
           "l2": "2 GiB per tensor, far beyond L2"}
    del c, s, cg, sg, gout
    torch.cuda.empty_cache()
    if cpu:
        sub = 16
        c1, s1 = R.synth_features((1, sub, h, w), cfg=5)
        cl1, sl1 = R.synth_labels(1, h, w, seed=4000), R.synth_labels(1, h, w, seed=5000)
        t = cpu_time_s(lambda: R.seg_adain_batch(c1, s1, cl1, sl1), 1) * (ch / sub)
        out["cpu_baseline"] = {"value": t * 1e3, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{sub} of 256 channels (channels are independent), scaled x{ch // sub}"}
    return out


def all_configs(dev, cpu=True):
    peaks = load_peaks()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {}
    for key, fn in (("0", config1), ("2", config3), ("3", config4), ("4", config5)):
        try:
            res[key] = fn(dev, peaks, cpu)
        except Exception as e:      # one config failing must not take the headline line down
            res[key] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------- PCIe / NUMA
def bind_to_gpu_numa_node(local: int):
    """Best effort: run this process on the CPUs of the NUMA node the GPU hangs off, BEFORE pinned buffers are
    allocated (first-touch), so host<->device copies do not cross the socket interconnect."""
    try:
        props = torch.cuda.get_device_properties(local)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
        return {"numa_node": node, "cpus": len(use), "bound": bool(use), "pci": bdf}
    except Exception as e:
        return {"bound": False, "why": repr(e)[:120]}


def pcie_bandwidth(dev, mib=256, iters=4):
    """Pinned H2D and D2H copies running CONCURRENTLY (two streams), GB/s each — the denominator of the e2e leg."""
    n = mib << 20
    hin = torch.empty(n, dtype=torch.uint8).pin_memory()
    hout = torch.empty(n, dtype=torch.uint8).pin_memory()
    din = torch.empty(n, dtype=torch.uint8, device=dev)
    dout = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(2):
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s1):
        e[0].record(s1)
        for _ in range(iters):
            din.copy_(hin, non_blocking=True)
        e[1].record(s1)
    with torch.cuda.stream(s2):
        e[2].record(s2)
        for _ in range(iters):
            hout.copy_(dout, non_blocking=True)
        e[3].record(s2)
    torch.cuda.synchronize()
    return {"h2d_GBs": n * iters / (e[0].elapsed_time(e[1]) / 1e3) / 1e9, "d2h_GBs": n * iters / (e[2].elapsed_time(e[3]) / 1e3) / 1e9,
            "how": f"{mib} MiB pinned copies x{iters}, H2D and D2H concurrently on two streams"}


# ------------------------------------------------------------------------------------------------- e2e_images
class StubRPNet(torch.nn.Module):
    """Shapes of config/v100/train_deeper_multiscale_rp_adain.yaml:27-34 (hidden_dim 16, rp_blocks 5, deeper,
    inception_num 3): encoder 3->16->32->64->128->256, stride 1, reflect-pad 3x3 + three 1x1 convs + LeakyReLU(0.2);
    decoder 256->128->64->32->16->3.  Random weights (no checkpoints ship with the reference); the convolutions are
    cuDNN (out of the transform's scope); `decode` is rpst's fused loop."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        dims = [3, 16, 32, 64, 128, 256]

        def block(i, o, act=True, inception=3):
            layers = [nn.ReflectionPad2d(1), nn.Conv2d(i, o, 3, 1)]
            layers += [nn.Conv2d(o, o, 1, 1) for _ in range(inception)]
            if act:
                layers.append(nn.LeakyReLU(0.2, inplace=True))
            return nn.Sequential(*layers)
        self.rp_shared_encoder = nn.ModuleList([block(dims[i], dims[i + 1]) for i in range(5)])
        self.rp_decoder = nn.ModuleList([block(256, 128, inception=0), block(128, 64, inception=0), block(64, 32, inception=0),
                                         block(32, 16, inception=0), block(16, 3, act=False, inception=0)])
        self._sort = False

    def encode_rp_intermediate(self, x):
        feats = []
        for enc in self.rp_shared_encoder:
            x = enc(x)
            feats.append(x)
        return feats


def stylize_images(net, content, style, oracle=False):
    """`MultiScaleAdaINRPNet.test` without the file handling: encode both images, multiscale decode."""
    cf, sf = net.encode_rp_intermediate(content), net.encode_rp_intermediate(style)
    if oracle:   # CPU arm: the reference's op sequence in eager torch
        from oracle import restate as R
        st = net.rp_decoder[0](R.adain_aten(cf[-1], sf[-1]))
        for i, l in enumerate(range(len(cf) - 2, -1, -1)):
            st = net.rp_decoder[i + 1](st + R.adain_aten(cf[l], sf[l]))
        return st
    from rpst.decode import decode_multiscale
    return decode_multiscale(net, cf, sf)


def e2e_images(dev, world, images_per_step=2, steps=6, cpu=True):
    """512^2 content/style IMAGES in pinned host memory -> H2D -> stub RP encoder -> rpst decode -> image -> D2H, every step."""
    torch.manual_seed(0)
    net = StubRPNet().to(dev).eval()
    g = torch.Generator().manual_seed(1001)
    hc = torch.rand(images_per_step, 3, 512, 512, generator=g).pin_memory()
    hs = torch.rand(images_per_step, 3, 512, 512, generator=g).pin_memory()
    ho = torch.empty(images_per_step, 3, 512, 512).pin_memory()

    def step():
        c = hc.to(dev, non_blocking=True)
        s = hs.to(dev, non_blocking=True)
        out = stylize_images(net, c, s)
        ho.copy_(out, non_blocking=True)
    with torch.no_grad():
        ms = dev_time_ms(step, steps, 2)
    res = {"value": world * images_per_step / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "images_per_step": images_per_step,
           "h2d_bytes_per_step": 2 * hc.numel() * 4, "d2h_bytes_per_step": ho.numel() * 4,
           "note": "stub RP encoder/decoder convolutions are cuDNN fp32 (random weights) and dominate; the transform is rpst.decode.decode_multiscale"}
    if cpu:
        cnet = StubRPNet().eval()
        cnet.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
        c1, s1 = hc[:1].clone(), hs[:1].clone()
        with torch.no_grad():
            t = cpu_time_s(lambda: stylize_images(cnet, c1, s1, oracle=True), 1)
            # parity of the whole image pipeline (TF32-free cuDNN fp32 vs CPU fp32): loose, conv order differs
            got = stylize_images(net, c1.to(dev), s1.to(dev)).cpu()
            want = stylize_images(cnet, c1, s1, oracle=True)
        res["cpu_baseline"] = {"value": 1.0 / t, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1 image pair through the same stub network, eager torch CPU, reference AdaIN op sequence"}
        res["rel_l2_vs_cpu"] = float((got - want).norm() / want.norm())
    return res
