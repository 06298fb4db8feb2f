"""GPU parity of the fused loss statistics through the C ABI: calc_style_loss and
calc_content_loss(norm=True) against the oracle, the reference-generated golden vector, and fp64
autograd of the reference formula.  Contract tolerance 1e-3 (fp32); asserted tighter where the
arithmetic allows."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rpst():
    import rpst as m
    return m


def rel(a, b):
    a, b = float(a.detach()) if torch.is_tensor(a) else float(a), float(b.detach()) if torch.is_tensor(b) else float(b)
    return abs(a - b) / max(abs(b), 1e-30)


def test_golden(rpst, golden):
    g = golden("losses")
    x, y, near = g["x"].cuda(), g["y"].cuda(), g["near"].cuda()
    assert rel(rpst.calc_style_loss(x, y), g["style"]) < 1e-5
    assert rel(rpst.calc_content_loss(x, y, norm=True), g["content_norm"]) < 1e-5
    assert rel(rpst.calc_content_loss(near, x, norm=True), g["content_norm_near"]) < 1e-3
    assert rel(rpst.calc_content_loss(x, y), g["content_plain"]) < 1e-5


# scalar path (odd hw), vector path with ragged last chunk, multi-chunk planes, one big plane
SHAPES = [(2, 3, 7, 9), (1, 64, 64, 64), (2, 5, 100, 100), (3, 4, 129, 131), (1, 8, 512, 512), (1, 2, 1024, 2048)]


@pytest.mark.parametrize("shape", SHAPES)
def test_vs_oracle(rpst, shape):
    x, y = R.synth_features(shape, cfg=40)
    style, content, stats = rpst.losses.pair_statistics(x.cuda(), y.cuda())
    assert rel(style, R.style_loss(x, y, dtype=torch.float64)) < 1e-5
    assert rel(content, R.content_loss(x, y, norm=True, dtype=torch.float64)) < 1e-5
    n, c = shape[:2]
    mx, sx = R.plane_stats(x, dtype=torch.float64)
    my, sy = R.plane_stats(y, dtype=torch.float64)
    st = stats.cpu().double()
    assert R.rel_l2(st[:, 0], mx.reshape(-1)) < 2e-6 and R.rel_l2(st[:, 1], sx.reshape(-1)) < 2e-6
    assert R.rel_l2(st[:, 2], my.reshape(-1)) < 2e-6 and R.rel_l2(st[:, 3], sy.reshape(-1)) < 2e-6
    # the entry points agree with the combined call and with calc_mean_std
    assert float(rpst.calc_style_loss(x.cuda(), y.cuda())) == float(style)
    m2, s2 = rpst.calc_mean_std(x.cuda())
    assert R.rel_l2(st[:, 0], m2.reshape(-1)) < 2e-6 and R.rel_l2(st[:, 1], s2.reshape(-1)) < 2e-6


def test_identical_inputs_give_zero(rpst):
    x, _ = R.synth_features((2, 4, 96, 96), cfg=41)
    xc = x.cuda()
    assert float(rpst.calc_style_loss(xc, xc.clone())) == 0.0
    # normalised content loss of identical tensors: cancellation leaves at most fp32 rounding of O(1) terms
    assert abs(float(rpst.calc_content_loss(xc, xc.clone(), norm=True))) < 1e-5


def test_hw1_is_nan_like_torch(rpst):
    x = torch.randn(1, 2, 1, 1).cuda()
    assert torch.isnan(rpst.calc_style_loss(x, x + 1))


def test_shape_mismatch_asserts(rpst):
    x = torch.randn(1, 2, 8, 8).cuda()
    with pytest.raises(AssertionError):
        rpst.calc_style_loss(x, x[:, :, :4])
    with pytest.raises(RuntimeError):
        rpst.calc_style_loss(x.cpu(), x.cpu())


@pytest.mark.parametrize("shape", [(2, 3, 9, 11), (1, 4, 128, 160)])
@pytest.mark.parametrize("which", ["style", "content"])
def test_autograd(rpst, shape, which):
    x, y = R.synth_features(shape, cfg=42, signed=True)
    xd, yd = x.double().requires_grad_(), y.double().requires_grad_()
    # fp64 autograd of the reference formula (network/base.py:399-407 + nn.MSELoss)
    def stats(t):
        n, c = t.shape[:2]
        v = t.reshape(n, c, -1).var(dim=2) + 1e-5
        return t.reshape(n, c, -1).mean(dim=2), v.sqrt()
    mxd, sxd = stats(xd)
    myd, syd = stats(yd)
    if which == "style":
        want = ((mxd - myd) ** 2).mean() + ((sxd - syd) ** 2).mean()
    else:
        nx = (xd - mxd[:, :, None, None]) / sxd[:, :, None, None]
        ny = (yd - myd[:, :, None, None]) / syd[:, :, None, None]
        want = ((nx - ny) ** 2).mean()
    (want * 3.0).backward()
    xg, yg = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    got = rpst.calc_style_loss(xg, yg) if which == "style" else rpst.calc_content_loss(xg, yg, norm=True)
    (got * 3.0).backward()
    assert rel(got, want) < 1e-5
    assert R.rel_l2(xg.grad, xd.grad) < 1e-4, R.rel_l2(xg.grad, xd.grad)
    assert R.rel_l2(yg.grad, yd.grad) < 1e-4, R.rel_l2(yg.grad, yd.grad)
    # target without grad (the usual case: style features come from the frozen VGG)
    xg2 = x.cuda().requires_grad_()
    got2 = rpst.calc_style_loss(xg2, y.cuda()) if which == "style" else rpst.calc_content_loss(xg2, y.cuda(), norm=True)
    got2.backward()
    assert R.rel_l2(xg2.grad * 3.0, xd.grad) < 1e-4
