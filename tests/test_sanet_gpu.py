"""GPU parity tests for SANet / AdaptiveSANet / Transform."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL32 = 1e-3    # fp32 contract (bf16x3 tensor-core products)
TOL16 = 1e-2    # bf16 contract


@pytest.fixture(scope="module")
def rpst():
    import rpst as m
    return m


def _sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def test_golden_sanet_and_transform(rpst, golden):
    g = golden("sanet")
    with torch.no_grad():
        m = rpst.SANet(16).cuda()
        m.load_state_dict(_sub(g, "sanet."))
        assert R.rel_l2(m(g["c4"].cuda(), g["s4"].cuda()), g["sanet_out"]) < TOL32
        m.precision = "bf16"
        assert R.rel_l2(m(g["c4"].cuda(), g["s4"].cuda()), g["sanet_out"]) < TOL16
        tr = rpst.Transform(16).cuda()
        tr.load_state_dict(_sub(g, "transform."))
        out = tr(g["c4"].cuda(), g["s4"].cuda(), g["c5"].cuda(), g["s5"].cuda())
        assert R.rel_l2(out, g["transform_out"]) < TOL32
        assert R.rel_l2(rpst.cal_affinity_matrix(g["c4"].cuda(), g["s4"].cuda()), g["affinity"]) < 1e-5


@pytest.mark.parametrize("mode", ["aea", "relu"])
def test_golden_adaptive(rpst, golden, mode):
    g = golden("sanet")
    with torch.no_grad():
        m = rpst.AdaptiveSANet(16, 64, ada_module=mode).cuda()
        m.load_state_dict(_sub(g, f"ada_{mode}."))
        out = m(g["c4"].cuda(), g["s4"].cuda())
        assert R.rel_l2(m.claim_before, g[f"ada_{mode}_before"]) < TOL32
        assert R.rel_l2(m.claim_value, g[f"ada_{mode}_clamp"]) < TOL32
        assert R.rel_l2(m.claim_after, g[f"ada_{mode}_after"]) < 5e-3   # sigmoid(50 x) amplifies S errors
        assert R.rel_l2(out, g[f"ada_{mode}_out"]) < 2e-3
        at = rpst.AdaptiveTransform(16, 64, 16, ada_module=mode).cuda()
        at.load_state_dict(_sub(g, f"adatr_{mode}."))
        out = at(g["c4"].cuda(), g["s4"].cuda(), g["c5"].cuda(), g["s5"].cuda())
        assert R.rel_l2(out, g[f"adatr_{mode}_out"]) < 2e-3


@pytest.mark.parametrize("b,c,hc,wc,hs,ws", [(2, 32, 12, 12, 12, 12), (1, 64, 20, 24, 16, 18), (1, 512, 32, 32, 32, 32)])
def test_attention_core_vs_oracle(rpst, b, c, hc, wc, hs, ws):
    g = torch.Generator().manual_seed(c)
    f = torch.randn(b, c, hc, wc, generator=g) * 0.5
    k = torch.randn(b, c, hs, ws, generator=g) * 0.5
    v = torch.randn(b, c, hs, ws, generator=g)
    want = R.attention_core(f.reshape(b, c, -1).double(), k.reshape(b, c, -1).double(), v.reshape(b, c, -1).double())
    got, attn = rpst.attention_core(f.cuda(), k.cuda(), v.cuda(), return_attn=True)
    assert R.rel_l2(got.reshape(b, c, -1), want) < TOL32
    # rows of the attention map are probability distributions
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-4
    got16 = rpst.attention_core(f.cuda(), k.cuda(), v.cuda(), precision="bf16")
    assert R.rel_l2(got16.reshape(b, c, -1), want) < 5e-2


def test_sanet_relu4_1_size_vs_fp64(rpst):
    """in_planes 512 at 64x64 (relu4_1 @512^2, L=4096): full module against an fp64 evaluation on the GPU."""
    torch.manual_seed(0)
    m = rpst.SANet(512).cuda()
    c, s = R.synth_features((1, 512, 64, 64), cfg=4, device="cuda")
    with torch.no_grad():
        got = m(c, s)
        sd = {k: v.double() for k, v in m.state_dict().items()}

        def mvn(x):
            f = x.reshape(1, 512, -1)
            return ((f - f.mean(2, keepdim=True)) / (f.var(2, keepdim=True) + 1e-5).sqrt())
        conv = lambda x, n: torch.einsum("oc,ncl->nol", sd[n + ".weight"].reshape(512, 512), x) + sd[n + ".bias"].reshape(1, -1, 1)
        cd, sdd = c.double(), s.double()
        F = conv(mvn(cd), "f")
        G = conv(mvn(sdd), "g")
        H = conv(sdd.reshape(1, 512, -1), "h")
        P = torch.softmax(torch.bmm(F.transpose(1, 2), G), dim=-1)
        O = torch.bmm(H, P.transpose(1, 2))
        want = (conv(O, "out_conv") + cd.reshape(1, 512, -1)).reshape(1, 512, 64, 64)
    assert R.rel_l2(got, want) < TOL32


@pytest.mark.parametrize("b,c,hc,wc,hs,ws", [(2, 32, 12, 12, 12, 12), (1, 64, 20, 24, 16, 18), (1, 24, 7, 9, 11, 5),
                                             (1, 512, 32, 32, 32, 32)])
def test_attention_core_backward_vs_fp64_autograd(rpst, b, c, hc, wc, hs, ws):
    """dF, dG, dH of softmax(F^T G) applied to H against fp64 autograd of the
    reference formula (network/sanet.py:85-94); ragged tiles, lc != ls, and the relu5_1 training size."""
    g = torch.Generator().manual_seed(100 + c)
    f = torch.randn(b, c, hc, wc, generator=g) * 0.5
    k = torch.randn(b, c, hs, ws, generator=g) * 0.5
    v = torch.randn(b, c, hs, ws, generator=g)
    w = torch.randn(b, c, hc, wc, generator=g)
    fd, kd, vd = (t.double().requires_grad_() for t in (f, k, v))
    S = torch.softmax(torch.bmm(fd.reshape(b, c, -1).transpose(1, 2), kd.reshape(b, c, -1)), dim=-1)
    O = torch.bmm(vd.reshape(b, c, -1), S.transpose(1, 2)).reshape(b, c, hc, wc)
    (O * w.double()).sum().backward()
    fg, kg, vg = (t.cuda().requires_grad_() for t in (f, k, v))
    out = rpst.attention_core(fg, kg, vg)
    assert R.rel_l2(out, O) < TOL32
    (out * w.cuda()).sum().backward()
    assert R.rel_l2(vg.grad, vd.grad) < TOL32, R.rel_l2(vg.grad, vd.grad)
    assert R.rel_l2(fg.grad, fd.grad) < TOL32, R.rel_l2(fg.grad, fd.grad)
    assert R.rel_l2(kg.grad, kd.grad) < TOL32, R.rel_l2(kg.grad, kd.grad)
    # bf16 operands: contract 1e-2 on the output; gradients pass through two more bf16 products
    fb, kb, vb = (t.cuda().requires_grad_() for t in (f, k, v))
    (rpst.attention_core(fb, kb, vb, precision="bf16") * w.cuda()).sum().backward()
    assert R.rel_l2(vb.grad, vd.grad) < 5e-2 and R.rel_l2(fb.grad, fd.grad) < 5e-2 and R.rel_l2(kb.grad, kd.grad) < 5e-2


def test_sanet_module_training_step_matches_fp64(rpst):
    """Whole SANet module (mvn -> 1x1 convs -> attention -> out_conv + residual) forward+backward:
    parameter and input gradients against the same module evaluated with fp64 torch ops."""
    torch.manual_seed(3)
    m = rpst.SANet(32).cuda()
    c, s = R.synth_features((2, 32, 16, 16), cfg=9, device="cuda", signed=True)
    w = torch.randn(2, 32, 16, 16, device="cuda")
    cg, sg = c.clone().requires_grad_(), s.clone().requires_grad_()
    (m(cg, sg) * w).sum().backward()
    got = {n: p.grad.clone() for n, p in m.named_parameters()}

    m64 = rpst.SANet(32).cuda().double()
    m64.load_state_dict({k_: v_.double() for k_, v_ in m.state_dict().items()})
    cd, sd = c.double().requires_grad_(), s.double().requires_grad_()

    def mvn(x):
        mu = x.mean((2, 3), keepdim=True)
        return (x - mu) / (x.var((2, 3), keepdim=True) + 1e-5).sqrt()
    F, G, H = m64.f(mvn(cd)), m64.g(mvn(sd)), m64.h(sd)
    bsz, ch, hh, ww = F.shape
    S = torch.softmax(torch.bmm(F.view(bsz, ch, -1).permute(0, 2, 1), G.view(bsz, ch, -1)), dim=-1)
    O = torch.bmm(H.view(bsz, ch, -1), S.permute(0, 2, 1)).view(bsz, ch, hh, ww)
    O = m64.out_conv(O) + cd
    (O * w.double()).sum().backward()
    for n, p in m64.named_parameters():
        if n == "g.bias":   # softmax is invariant to a per-row shift of S: the true gradient is exactly zero
            assert float(got[n].abs().max()) < 1e-3 * float(m64.g.weight.grad.abs().max())
            continue
        assert R.rel_l2(got[n], p.grad) < TOL32, (n, R.rel_l2(got[n], p.grad))
    assert R.rel_l2(cg.grad, cd.grad) < TOL32 and R.rel_l2(sg.grad, sd.grad) < TOL32


@pytest.mark.parametrize("mode", ["aea", "relu"])
def test_golden_adaptive_gradients(rpst, golden, mode):
    """AdaptiveSANet under autograd (the AdaptiveSAModel training path, network/sanet.py:373-384): output, input
    gradients and every parameter gradient — including the clamp MLP f_psi — against what the REFERENCE module
    produced through autograd on CPU (tests/golden/adaptive_grad.npz)."""
    g = golden("adaptive_grad")
    pre = mode + "."
    m = rpst.AdaptiveSANet(16, 64, ada_module=mode).cuda()
    m.load_state_dict({k[len(pre) + 6:]: v for k, v in g.items() if k.startswith(pre + "param.")})
    c, s = g[pre + "content"].cuda().requires_grad_(), g[pre + "style"].cuda().requires_grad_()
    out = m(c, s)
    assert R.rel_l2(m.claim_value, g[pre + "clamp"]) < TOL32
    assert R.rel_l2(out, g[pre + "out"]) < 2e-3             # sigmoid(50 x) amplifies S errors, as in the forward test
    (out * g[pre + "w"].cuda()).sum().backward()
    tol = 5e-3 if mode == "aea" else 2e-3
    assert R.rel_l2(c.grad, g[pre + "grad_content"]) < tol, R.rel_l2(c.grad, g[pre + "grad_content"])
    assert R.rel_l2(s.grad, g[pre + "grad_style"]) < tol, R.rel_l2(s.grad, g[pre + "grad_style"])
    scale = float(g[pre + "grad.g.weight"].abs().max())
    for name, p in m.named_parameters():
        want = g[pre + "grad." + name]
        assert p.grad is not None, name
        if float(want.abs().max()) < 1e-6 * max(scale, 1e-30):     # exactly-zero gradients (e.g. g.bias under plain softmax)
            assert float((p.grad.cpu() - want).abs().max()) < 1e-3 * scale, name
            continue
        assert R.rel_l2(p.grad, want) < tol, (name, R.rel_l2(p.grad, want))



def test_sample_groups_match_per_sample_launches(rpst, monkeypatch):
    """k samples per launch (workspace = k x per-sample size) must give the per-sample results bit for bit:
    group sizes 1, 2+1 and 3 on a ragged shape, forward and backward."""
    b, c, hc, wc, hs, ws = 3, 40, 9, 11, 10, 7
    g = torch.Generator().manual_seed(7)
    f, k, v, w = (torch.randn(b, c, *hw, generator=g).cuda() for hw in ((hc, wc), (hs, ws), (hs, ws), (hc, wc)))
    L = rpst._lib.lib()
    per_f = L.rpst_sanet_attn_workspace_bytes(c, hc * wc, hs * ws)
    per_b = L.rpst_sanet_attn_bwd_workspace_bytes(c, hc * wc, hs * ws)
    results = []
    for k_group in (1, 2, 3):
        monkeypatch.setattr(rpst.sanet, "_GROUP_BYTES", k_group * max(per_f, per_b))
        fg, kg, vg = (t.clone().requires_grad_() for t in (f, k, v))
        out, attn = rpst.attention_core(fg, kg, vg, return_attn=True)
        (out * w).sum().backward()
        results.append((out.detach(), attn, fg.grad, kg.grad, vg.grad))
    for other in results[1:]:
        for a, b_ in zip(results[0], other):
            assert torch.equal(a, b_)


def test_long_rows_take_the_register_resident_row_kernels(rpst):
    """L = 8192 (rows of 8192 columns -> attn_rows_reg_kernel / attn_bwd_rows_reg_kernel), forward and
    backward against fp64 torch ops evaluated on the GPU (the fp64 formula of network/sanet.py:85-94)."""
    b, c, hc, wc = 1, 64, 64, 128
    g = torch.Generator().manual_seed(8192)
    f, k, v, w = (torch.randn(b, c, hc, wc, generator=g).cuda() * sc for sc in (0.4, 0.4, 1.0, 1.0))
    fd, kd, vd = (t.double().requires_grad_() for t in (f, k, v))
    S = torch.softmax(torch.bmm(fd.reshape(b, c, -1).transpose(1, 2), kd.reshape(b, c, -1)), dim=-1)
    O = torch.bmm(vd.reshape(b, c, -1), S.transpose(1, 2)).reshape(b, c, hc, wc)
    (O * w.double()).sum().backward()
    fg, kg, vg = (t.clone().requires_grad_() for t in (f, k, v))
    out, attn = rpst.attention_core(fg, kg, vg, return_attn=True)
    assert R.rel_l2(out, O.detach()) < TOL32
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-4
    assert R.rel_l2(attn, S.detach()) < TOL32
    (out * w).sum().backward()
    for got, want in ((vg.grad, vd.grad), (fg.grad, fd.grad), (kg.grad, kd.grad)):
        assert R.rel_l2(got, want) < TOL32, R.rel_l2(got, want)


def test_golden_sanet_gradients(rpst, golden):
    """Input and parameter gradients the REFERENCE SANet module produced through autograd on CPU
    (tests/golden/sanet_grad.npz) against the rpst module (rpst_sanet_attn_bwd + AdaIN backward)."""
    g = golden("sanet_grad")
    m = rpst.SANet(16).cuda()
    m.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("param.")})
    c, s = g["content"].cuda().requires_grad_(), g["style"].cuda().requires_grad_()
    out = m(c, s)
    assert R.rel_l2(out, g["out"]) < TOL32
    (out * g["w"].cuda()).sum().backward()
    assert R.rel_l2(c.grad, g["grad_content"]) < TOL32 and R.rel_l2(s.grad, g["grad_style"]) < TOL32
    scale = float(g["grad.g.weight"].abs().max())
    for name, p in m.named_parameters():
        want = g["grad." + name]
        if name == "g.bias":      # exactly zero in exact arithmetic (softmax shift invariance): compare absolutely
            assert float((p.grad.cpu() - want).abs().max()) < 1e-3 * scale
            continue
        assert R.rel_l2(p.grad, want) < TOL32, (name, R.rel_l2(p.grad, want))


def test_cal_affinity_matrix_is_differentiable(rpst):
    """network/sanet.py:12-18 under autograd: gradient of the cosine affinity w.r.t. both feature maps."""
    c, s = R.synth_features((2, 12, 6, 5), cfg=16, signed=True)
    w = torch.randn(2, 30, 30, generator=torch.Generator().manual_seed(17))
    cd, sd = c.double().requires_grad_(), s.double().requires_grad_()
    nc = torch.nn.functional.normalize(cd.view(2, 12, -1), dim=1)
    ns = torch.nn.functional.normalize(sd.view(2, 12, -1), dim=1)
    (torch.bmm(nc.permute(0, 2, 1), ns) * w.double()).sum().backward()
    cg, sg = c.cuda().requires_grad_(), s.cuda().requires_grad_()
    (rpst.cal_affinity_matrix(cg, sg) * w.cuda()).sum().backward()
    assert R.rel_l2(cg.grad, cd.grad) < 1e-4 and R.rel_l2(sg.grad, sd.grad) < 1e-4


@pytest.mark.parametrize("mode", ["aea", "relu"])
@pytest.mark.parametrize("c,h,w", [(48, 15, 13), (64, 16, 24)])
def test_adaptive_ragged_shapes_fused_and_unfused_row_pass(rpst, mode, c, h, w):
    """AdaptiveSANet at L = 195 / 384 (partial row tiles, partial 64-wide K tiles): the affinity leaves its GEMM as packed
    operand tiles (zero padding written by the epilogue) and softmax + clamp run as ONE row pass when the intermediate
    maps are not kept; with `keep_claims` the two-pass form runs.  Both against the fp64 oracle (network/sanet.py:114-138)."""
    torch.manual_seed(3)
    L = h * w
    m = rpst.AdaptiveSANet(c, L, ada_module=mode).cuda()
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ct, st = R.synth_features((2, c, h, w), cfg=4)
    want, before, after, clamp = R.adaptive_sanet_forward(ct, st, params, mode, dtype=torch.float64)
    with torch.no_grad():
        m.keep_claims = False
        fused = m(ct.cuda(), st.cuda())
        m.keep_claims = True
        kept = m(ct.cuda(), st.cuda())
    assert R.rel_l2(fused, want) < 2e-3, R.rel_l2(fused, want)
    assert R.rel_l2(kept, want) < 2e-3, R.rel_l2(kept, want)
    assert R.rel_l2(fused, kept) < 1e-5            # same arithmetic up to the order of the two softmax passes
    assert R.rel_l2(m.claim_before, before) < 1e-3
    assert R.rel_l2(m.claim_value.reshape(-1), clamp.reshape(-1)) < 1e-3
