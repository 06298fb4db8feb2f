"""GPU parity tests of the fused pointwise convolution (csrc/pwconv.cu) — SURVEY.md section 8f rank 2 (SANet projections
with the instance normalisation folded in, network/sanet.py:82-99) and rank 4 (RP-encoder 1x1 conv + LeakyReLU with the
AdaIN statistics from the conv epilogue, network/base.py:170-198 + :399-418)."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL32 = 1e-3
TOL16 = 1e-2


@pytest.fixture(scope="module")
def rpst():
    import rpst as m
    return m


def _ref_conv(x, w, b=None, sub=None, mul=None, residual=None, slope=None):
    x = x.double()
    if sub is not None:
        x = (x - sub.double()) * (mul.double() if mul is not None else 1.0)
    y = torch.einsum("oc,nchw->nohw", w.double().reshape(w.shape[0], -1), x)
    if b is not None:
        y = y + b.double().view(1, -1, 1, 1)
    if slope is not None:
        y = torch.where(y >= 0, y, slope * y)
    if residual is not None:
        y = y + residual.double()
    return y


@pytest.mark.parametrize("b,cin,cout,h,w", [(1, 64, 64, 16, 16), (2, 256, 256, 32, 32), (1, 512, 512, 16, 24), (2, 48, 200, 9, 7),
                                            (1, 3, 16, 33, 31), (1, 300, 512, 12, 12)])
def test_conv1x1_vs_fp64(rpst, b, cin, cout, h, w):
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g)
    res = torch.randn(b, cout, h, w, generator=g)
    sub = torch.randn(b, cin, 1, 1, generator=g)
    mul = torch.rand(b, cin, 1, 1, generator=g) + 0.5
    got = rpst.conv1x1(x.cuda(), wt.cuda(), bias.cuda())
    assert R.rel_l2(got, _ref_conv(x, wt, bias)) < 2e-5
    got = rpst.conv1x1(x.cuda(), wt.cuda(), bias.cuda(), sub=sub.cuda(), mul=mul.cuda(), residual=res.cuda(), act="lrelu", slope=0.2)
    assert R.rel_l2(got, _ref_conv(x, wt, bias, sub, mul, res, 0.2)) < 2e-5
    got16 = rpst.conv1x1(x.cuda(), wt.cuda(), bias.cuda(), precision="bf16")
    assert R.rel_l2(got16, _ref_conv(x, wt, bias)) < TOL16


@pytest.mark.parametrize("b,c,h,w", [(2, 32, 64, 64), (1, 256, 48, 40), (1, 16, 13, 11)])
def test_conv_epilogue_statistics_feed_the_transform(rpst, b, c, h, w):
    """f4: the last RP-encoder op (1x1 conv + LeakyReLU(0.2)) emits calc_mean_std of its output from the epilogue;
    `adain_from_stats` then equals the reference `[prev +] AdaIN(content_feat, style_feat)` on the conv outputs."""
    g = torch.Generator().manual_seed(7 + c)
    xc, xs = torch.randn(b, c, h, w, generator=g), torch.randn(b, c, h, w, generator=g) * 2 + 0.5
    wt = torch.randn(c, c, 1, 1, generator=g) / c ** 0.5
    bias = torch.randn(c, generator=g) * 0.1
    prev = torch.randn(b, c, h, w, generator=g)
    fc, (mu_c, sd_c) = rpst.conv1x1(xc.cuda(), wt.cuda(), bias.cuda(), act="lrelu", want_stats=True)
    fs, (mu_s, sd_s) = rpst.conv1x1(xs.cuda(), wt.cuda(), bias.cuda(), act="lrelu", want_stats=True)
    ref_c, ref_s = _ref_conv(xc, wt, bias, slope=0.2), _ref_conv(xs, wt, bias, slope=0.2)
    wm, wsd = R.plane_stats(ref_c, dtype=torch.float64)
    assert mu_c.shape == (b, c, 1, 1)
    assert R.rel_l2(mu_c, wm) < 1e-5 and R.rel_l2(sd_c, wsd) < 1e-5
    gm, gs = rpst.calc_mean_std(fc)
    assert R.rel_l2(mu_c, gm) < 1e-5 and R.rel_l2(sd_c, gs) < 1e-5
    want = R.adain(ref_c, ref_s, dtype=torch.float64)
    assert R.rel_l2(rpst.adain_from_stats(fc, (mu_c, sd_c), (mu_s, sd_s)), want) < 1e-4
    assert R.rel_l2(rpst.adain_from_stats(fc, (mu_c, sd_c), (mu_s, sd_s), prev=prev.cuda()), want + prev.double()) < 1e-4


@pytest.mark.parametrize("side,b", [(16, 2), (32, 1), (64, 1)])
def test_fused_sanet_forward_matches_unfused_and_fp64(rpst, side, b):
    """f2: SANet(512) inference with the projections, the folded mean_variance_norm and the residual on the tcgen05
    block (Q / K handed to the attention kernel as packed operands) vs the cuDNN-convolution path and vs fp64."""
    torch.manual_seed(0)
    m = rpst.SANet(512).cuda()
    c, s = R.synth_features((b, 512, side, side), cfg=43, device="cuda")
    with torch.no_grad():
        fused = m(c, s)
        m.fused = False
        plain = m(c, s)
        m.fused = True
        sd = {k: v.double() for k, v in m.state_dict().items()}

        def mvn(x):
            f = x.reshape(b, 512, -1)
            return (f - f.mean(2, keepdim=True)) / (f.var(2, keepdim=True) + 1e-5).sqrt()
        conv = lambda x, n: torch.einsum("oc,ncl->nol", sd[n + ".weight"].reshape(512, 512), x) + sd[n + ".bias"].reshape(1, -1, 1)
        cd, sdd = c.double(), s.double()
        F, G, H = conv(mvn(cd), "f"), conv(mvn(sdd), "g"), conv(sdd.reshape(b, 512, -1), "h")
        P = torch.softmax(torch.bmm(F.transpose(1, 2), G), dim=-1)
        want = (conv(torch.bmm(H, P.transpose(1, 2)), "out_conv") + cd.reshape(b, 512, -1)).reshape(b, 512, side, side)
        m.precision = "bf16"
        fused16 = m(c, s)
    assert R.rel_l2(fused, want) < TOL32, R.rel_l2(fused, want)
    assert R.rel_l2(plain, want) < TOL32
    assert R.rel_l2(fused16, want) < TOL16, R.rel_l2(fused16, want)


def test_fused_path_is_inference_only(rpst):
    """with autograd on, SANet keeps the differentiable path (parameters train through cuDNN + the attention backward)"""
    torch.manual_seed(1)
    m = rpst.SANet(512).cuda()
    c, s = R.synth_features((1, 512, 16, 16), cfg=44, device="cuda")
    out = m(c, s)
    assert out.requires_grad
    out.sum().backward()
    assert m.f.weight.grad is not None and torch.isfinite(m.f.weight.grad).all()
