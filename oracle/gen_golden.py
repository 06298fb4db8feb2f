"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions (container only).

    python -m oracle.gen_golden            # rewrites tests/golden/

The reference ships no tests/golden vectors, so these fixtures — inputs and the
outputs the reference itself produced on CPU (torch 2.11, fp32 unless the reference casts to fp64)
— are the pin for both the oracle (`oracle/restate.py`) and the CUDA path.  Everything is seeded;
inputs are stored next to outputs so the fixtures stay valid if a torch release changes `randn`.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402
from oracle.restate import synth_features, synth_labels  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (_np(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def gen_adain(net):
    base = sys.modules["network.base"]
    cases = {}
    for tag, shape, signed in [("odd", (2, 5, 7, 9), False), ("sq", (1, 8, 16, 16), False),
                               ("signed", (3, 4, 12, 10), True), ("hw2", (2, 3, 1, 2), True)]:
        c, s = synth_features(shape, cfg=len(cases), signed=signed)
        g = torch.Generator().manual_seed(77 + len(cases))
        prev = torch.randn(shape, generator=g)
        mean, std = base.calc_mean_std(c)
        out = base.adaptive_instance_normalization(c, s)
        cases[tag] = True
        _save(f"adain_{tag}", content=c, style=s, prev=prev, mean=mean, std=std, out=out,
              blend=prev + out, mvn=sys.modules["network.sanet"].mean_variance_norm(c))


def gen_seg(net):
    base = sys.modules["network.base"]
    c, s = synth_features((1, 6, 24, 20), cfg=11)
    s = torch.relu(torch.randn((1, 6, 18, 30), generator=torch.Generator().manual_seed(3011)) * 2 + 1)
    cl = synth_labels(1, 24, 20, classes=5, block=6, seed=4000)[0]
    sl = synth_labels(1, 18, 30, classes=5, block=6, seed=5000)[0]
    # adversarial labels (network/base.py:435): label 7 absent in style, label 8 has <=10 content
    # pixels, label 9 has a count ratio >= 100 (handled via a tiny style region)
    cl[0:3, 0:4] = 7
    cl[10:12, 0:4] = 8           # 8 pixels -> invalid
    sl[0:2, 0:6] = 8
    cl[12:24, 10:20] = 9         # 120 px content
    sl[17, 29] = 9               # 1 px style -> cnt_s<=10 -> invalid
    saved = base.get_segment_and_info
    base.get_segment_and_info = lambda cp, sp, csh, ssh: (
        cl.numpy(), sl.numpy(), *base.compute_label_info(cl.numpy(), sl.numpy()))
    try:
        out = base.adaptive_instance_normalization_with_segment(c, s, None, None)
    finally:
        base.get_segment_and_info = saved
    _save("seg_adain", content=c, style=s, c_labels=cl, s_labels=sl, out=out)


def gen_wct(net):
    wct = sys.modules["network.wct_rp"]
    g = torch.Generator().manual_seed(21)
    a = torch.randn(16, 40, generator=g, dtype=torch.float64)
    spd = a @ a.t() / 39
    _save("wct_matfn", a=spd, sqrt=wct.matrix_sqrt(spd), inv_sqrt=wct.matrix_inv_sqrt(spd))
    c, s = synth_features((2, 16, 20, 20), cfg=3)
    dummy = object.__new__(wct.WCTRPNet)  # whiten_and_color / fuse use no module state
    cf, sf = c[0].reshape(16, -1).double(), s[0].reshape(16, -1).double()
    _save("wct", content=c, style=s,
          closed_form=wct.WCTRPNet.whiten_and_color(dummy, cf, sf),
          original=wct.WCTRPNet.whiten_and_color(dummy, cf, sf, method="original"),
          fuse=wct.WCTRPNet.fuse(dummy, c, s))


def gen_sanet(net):
    sanet = sys.modules["network.sanet"]
    torch.manual_seed(0)
    planes = 16
    c4, s4 = synth_features((2, planes, 8, 8), cfg=4)
    c5, s5 = synth_features((2, planes, 4, 4), cfg=5)
    m = sanet.SANet(planes)
    arrays = {"c4": c4, "s4": s4, "c5": c5, "s5": s5, "sanet_out": m(c4, s4)}
    arrays.update({"sanet." + k: v for k, v in m.state_dict().items()})
    tr = sanet.Transform(planes)
    arrays["transform_out"] = tr(c4, s4, c5, s5)
    arrays.update({"transform." + k: v for k, v in tr.state_dict().items()})
    for mode in ("aea", "relu"):
        am = sanet.AdaptiveSANet(planes, 64, ada_module=mode)
        arrays[f"ada_{mode}_out"] = am(c4, s4)
        arrays[f"ada_{mode}_before"] = am.claim_before
        arrays[f"ada_{mode}_after"] = am.claim_after
        arrays[f"ada_{mode}_clamp"] = am.claim_value
        arrays.update({f"ada_{mode}." + k: v for k, v in am.state_dict().items()})
        at = sanet.AdaptiveTransform(planes, 64, 16, ada_module=mode)
        arrays[f"adatr_{mode}_out"] = at(c4, s4, c5, s5)
        arrays.update({f"adatr_{mode}." + k: v for k, v in at.state_dict().items()})
    arrays["affinity"] = sanet.cal_affinity_matrix(c4, s4)
    with torch.no_grad():
        _save("sanet", **arrays)


def gen_mrf(net):
    base = sys.modules["network.base"]
    mrf = sys.modules["network.mrf_rp"]
    c, s = synth_features((1, 16, 8, 8), cfg=6)
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # reference hard-codes .cuda() (base.py:325-343)
    try:
        k = 3
        aff = base.cal_affinity_map(c, s, k)
        loss = mrf.MRFLoss(k)(c, s)
        loss_all = mrf.MRFLoss(k, mean="all")(c, s)
    finally:
        torch.Tensor.cuda = saved
    cn = torch.nn.functional.normalize(c.squeeze(), dim=0).view(16, -1)
    sn = torch.nn.functional.normalize(s.squeeze(), dim=0).view(16, -1)
    ncc = cn.t() @ sn
    _save("mrf", content=c, style=s, k=k, affinity=aff, loss=loss, loss_all=loss_all,
          dist=base.cal_dist(c.view(16, -1), s.view(16, -1)),
          idx0=torch.topk(ncc, k, 0)[1], idx1=torch.topk(ncc, k, 1)[1])


def gen_se(net):
    att = sys.modules["network.attention"]
    torch.manual_seed(1)
    m = att.SELayer(32, reduction=16)
    x, _ = synth_features((2, 32, 6, 10), cfg=7)
    with torch.no_grad():
        _save("se", x=x, out=m(x), gate=m.attention_map,
              w1=m.fc[0].weight, w2=m.fc[2].weight)


def gen_losses(net):
    """calc_style_loss / calc_content_loss(norm) are methods that only touch `self.mse_loss`."""
    import types
    stub = types.SimpleNamespace(mse_loss=torch.nn.MSELoss())
    adain_rp, sanet = sys.modules["network.adain_rp"], sys.modules["network.sanet"]
    x, y = synth_features((2, 6, 9, 11), cfg=8)
    g = torch.Generator().manual_seed(88)
    near = x + 0.05 * torch.randn(x.shape, generator=g)      # stylized ~ content: small normalised loss
    _save("losses", x=x, y=y, near=near,
          style=adain_rp.AdaINRPNet.calc_style_loss(stub, x, y),
          style_sanet=sanet.SAModel.calc_style_loss(stub, x, y),
          content_norm=sanet.SAModel.calc_content_loss(stub, x, y, norm=True),
          content_norm_near=sanet.SAModel.calc_content_loss(stub, near, x, norm=True),
          content_plain=sanet.SAModel.calc_content_loss(stub, x, y))


def gen_channel_maps(net):
    """shuffle / sort_by_weights are methods that read only config-derived attributes and the encoders'
    stored attention maps: call them unbound on a stub."""
    import types
    adain_rp = sys.modules["network.adain_rp"]
    base = sys.modules["network.base"]
    c, s = synth_features((2, 8, 6, 5), cfg=9)
    g = torch.Generator().manual_seed(99)
    att = torch.rand(2, 8, 1, 1, generator=g)          # tie-free by construction (checked below)
    assert att.reshape(2, 8).sort(dim=1).values.diff(dim=1).abs().min() > 1e-4
    stub = types.SimpleNamespace(_shuffle_layers=10, rp_shared_encoder=[types.SimpleNamespace(attention_map=att)])
    shuffled = adain_rp.MultiScaleAdaINRPNet.shuffle(stub, c, 0)
    sorted_c = adain_rp.MultiScaleAdaINRPNet.sort_by_weights(stub, [c])[0]
    sorted_s = adain_rp.MultiScaleAdaINRPNet.sort_by_weights(stub, [s])[0]
    _save("channel_maps", content=c, style=s, attention=att, shuffled=shuffled, sorted_content=sorted_c,
          adain_sorted=base.adaptive_instance_normalization(sorted_c, sorted_s),
          adain_shuffled=base.adaptive_instance_normalization(shuffled, adain_rp.MultiScaleAdaINRPNet.shuffle(stub, s, 0)))


def decode_stub(kind, seed=0):
    """A module with exactly the attributes the reference's decode()/test() methods touch; shared by the
    generator (reference methods bound to it) and tests/test_decode_gpu.py (rpst mirrors bound to it)."""
    import types
    torch.manual_seed(seed)
    m = torch.nn.Module()
    if kind == "multiscale":          # features shallow -> deep: 4, 8, 16 channels (network/adain_rp.py:286-302)
        m.rp_decoder = torch.nn.ModuleList([torch.nn.Conv2d(16, 8, 3, padding=1), torch.nn.Conv2d(8, 4, 3, padding=1),
                                            torch.nn.Conv2d(4, 3, 3, padding=1)])
        m.rp_shared_encoder = [types.SimpleNamespace(attention_map=None) for _ in range(3)]
        m._sort, m._shuffle, m._shuffle_layers = False, False, 1
        m.config = {"use_mask": False}
    elif kind == "ldms":              # constant width 8 (network/adain_rp.py:538-553)
        for i in range(3):
            setattr(m, f"rp_dec{i}", torch.nn.Conv2d(8, 8 if i < 2 else 3, 3, padding=1))
        m.stylized_layers = 3
    elif kind == "ldcat":             # concat doubles the width (network/adain_rp.py:780-799)
        m.rp_dec0 = torch.nn.Conv2d(8, 8, 3, padding=1)
        m.rp_dec1 = torch.nn.Conv2d(16, 8, 3, padding=1)
        m.rp_dec2 = torch.nn.Conv2d(16, 3, 3, padding=1)
    return m


def decode_inputs(kind):
    chans = [4, 8, 16] if kind == "multiscale" else [8, 8, 8]
    cs, ss = [], []
    for i, ch in enumerate(chans):
        c, s = synth_features((2, ch, 12, 10), cfg=40 + i, signed=True)
        cs.append(c); ss.append(s)
    g = torch.Generator().manual_seed(4242)
    atts = [torch.rand(2, ch, 1, 1, generator=g) for ch in chans]
    return cs, ss, atts


def gen_decode(net):
    """The reference's own decode()/test() methods, bound to the stub modules above."""
    import types
    adain_rp = sys.modules["network.adain_rp"]
    arrays = {}
    for kind, cls in (("multiscale", adain_rp.MultiScaleAdaINRPNet), ("ldms", adain_rp.LDMSAdaINRPNet),
                      ("ldcat", adain_rp.LDMSAdaINRPNet4)):
        m = decode_stub(kind)
        cs, ss, atts = decode_inputs(kind)
        arrays.update({f"{kind}.{k}": v for k, v in m.state_dict().items()})
        arrays[f"{kind}.out"] = cls.decode(m, cs, ss)
        if kind == "multiscale":
            for enc, a in zip(m.rp_shared_encoder, atts):
                enc.attention_map = a
            m._sort = True
            m.sort_by_weights = types.MethodType(cls.sort_by_weights, m)
            arrays["multiscale.out_sorted"] = cls.decode(m, cs, ss)
            # test(): shuffle levels <= _shuffle_layers, then decode (sort still on)
            m._shuffle = True
            m.shuffle = types.MethodType(cls.shuffle, m)
            m.decode = types.MethodType(cls.decode, m)
            feats = {"c": cs, "s": ss}
            m.encode_rp_intermediate = lambda x: feats[x]
            arrays["multiscale.test_out"] = cls.test(m, "c", "s")
            for i, a in enumerate(atts):
                arrays[f"multiscale.att{i}"] = a
    _save("decode", **arrays)


def gen_sanet_grad(net):
    """Autograd through the reference SANet module (what SAModel.forward trains, network/sanet.py:253-276):
    gradients of the inputs and of every parameter for a weighted-sum loss."""
    sanet = sys.modules["network.sanet"]
    torch.manual_seed(7)
    m = sanet.SANet(16)
    c, s = synth_features((2, 16, 8, 8), cfg=12, signed=True)
    g = torch.Generator().manual_seed(13)
    w = torch.randn(2, 16, 8, 8, generator=g)
    c.requires_grad_(); s.requires_grad_()
    with torch.enable_grad():
        out = m(c, s)
        (out * w).sum().backward()
    arrays = {"content": c.detach(), "style": s.detach(), "w": w, "out": out.detach(),
              "grad_content": c.grad, "grad_style": s.grad}
    arrays.update({"param." + k: v.detach() for k, v in m.state_dict().items()})
    arrays.update({"grad." + k: p.grad for k, p in m.named_parameters()})
    _save("sanet_grad", **arrays)


def gen_adaptive_grad(net):
    """Autograd through the reference AdaptiveSANet (AdaptiveSAModel.forward trains it, network/sanet.py:373-384):
    input and parameter gradients, both clamp modules."""
    sanet = sys.modules["network.sanet"]
    arrays = {}
    for mode in ("aea", "relu"):
        torch.manual_seed(21)
        m = sanet.AdaptiveSANet(16, 64, ada_module=mode)
        c, s = synth_features((2, 16, 8, 8), cfg=14, signed=True)
        g = torch.Generator().manual_seed(15)
        w = torch.randn(2, 16, 8, 8, generator=g)
        c.requires_grad_(); s.requires_grad_()
        with torch.enable_grad():
            out = m(c, s)
            (out * w).sum().backward()
        arrays.update({f"{mode}.content": c.detach(), f"{mode}.style": s.detach(), f"{mode}.w": w, f"{mode}.out": out.detach(),
                       f"{mode}.grad_content": c.grad, f"{mode}.grad_style": s.grad,
                       f"{mode}.clamp": m.claim_value.detach()})
        arrays.update({f"{mode}.param." + k: v.detach() for k, v in m.state_dict().items()})
        arrays.update({f"{mode}.grad." + k: p.grad for k, p in m.named_parameters()})
    _save("adaptive_grad", **arrays)


def main():
    net = load_reference()
    with torch.no_grad():
        gen_adain(net)
        gen_seg(net)
        gen_wct(net)
        gen_sanet(net)
        gen_mrf(net)
        gen_se(net)
        gen_losses(net)
        gen_channel_maps(net)
        gen_decode(net)
    gen_sanet_grad(net)
    gen_adaptive_grad(net)


if __name__ == "__main__":
    main()
