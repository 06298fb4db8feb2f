"""Training-step parity on the GPU: the reference trains the decoder (and RP encoder) through AdaIN and
the style/content losses with plain autograd (network/adain_rp.py:321-345).  Here the same step is run
once with rpst's transform + statistics (custom backward kernels) and once in pure fp64 torch; losses and
every parameter gradient must agree, and the flat gradient bucket must reproduce them."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _stats64(x):
    n, c = x.shape[:2]
    f = x.reshape(n, c, -1)
    return f.mean(2).view(n, c, 1, 1), (f.var(2) + 1e-5).sqrt().view(n, c, 1, 1)


def _adain64(c, s):
    mc, sc = _stats64(c)
    ms, ss = _stats64(s)
    return (c - mc) / sc * ss + ms


class TinyRP(nn.Module):
    """Two-level resolution-preserving encoder/decoder in the shape of MultiScaleAdaINRPNet."""

    def __init__(self):
        super().__init__()
        self.enc = nn.ModuleList([nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU()),
                                  nn.Sequential(nn.Conv2d(8, 16, 3, padding=1), nn.ReLU())])
        self.dec = nn.ModuleList([nn.Sequential(nn.Conv2d(16, 8, 3, padding=1), nn.ReLU()),
                                  nn.Sequential(nn.Conv2d(8, 3, 3, padding=1))])

    def feats(self, x):
        out = []
        for m in self.enc:
            x = m(x)
            out.append(x)
        return out

    def forward(self, content, style, adain, blend, stats):
        cf, sf = self.feats(content), self.feats(style)
        y = self.dec[0](adain(cf[1], sf[1]))
        y = self.dec[1](blend(y, cf[0], sf[0]))
        # style loss on the (shared) encoder features of the stylized image + content loss
        yf = self.feats(y)
        loss = ((yf[1] - cf[1].detach()) ** 2).mean()
        for a, b in zip(yf, sf):
            ma, sa = stats(a)
            mb, sb = stats(b.detach())
            loss = loss + ((ma - mb) ** 2).mean() + ((sa - sb) ** 2).mean()
        return loss


@pytest.mark.parametrize("hw", [24, 160])   # 160x160 = 25600 > 16384: pipelined kernels; 24x24: direct kernels
def test_training_step_matches_fp64_autograd(hw):
    import rpst
    from rpst.dist import GradBucket
    # the convolutions around the transform run in cuDNN; keep them in true fp32 so that the comparison
    # with the fp64 graph measures the transform kernels, not TF32 convolution rounding
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = TinyRP().cuda()
    ref = TinyRP().double().cuda()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    g = torch.Generator(device="cuda").manual_seed(1)
    content = torch.rand(2, 3, hw, hw, device="cuda", generator=g)
    style = torch.rand(2, 3, hw, hw, device="cuda", generator=g)

    loss = net(content, style, rpst.adaptive_instance_normalization, rpst.adain_blend, rpst.calc_mean_std)
    loss.backward()
    loss64 = ref(content.double(), style.double(), _adain64, lambda p, c, s: p + _adain64(c, s), _stats64)
    loss64.backward()
    assert abs(float(loss) - float(loss64)) / abs(float(loss64)) < 2e-4
    for (name, p), q in zip(net.named_parameters(), ref.parameters()):
        err = float((p.grad.double() - q.grad).norm() / q.grad.norm().clamp_min(1e-30))
        assert err < 2e-3, (name, err)
    # one flat bucket (world size 1 here; NCCL all-reduce when launched under torchrun)
    before = [p.grad.clone() for p in net.parameters()]
    out = GradBucket(net.parameters()).allreduce_mean({"loss": loss})
    assert all(torch.equal(a, p.grad) for a, p in zip(before, net.parameters()))
    assert abs(float(out["loss"]) - float(loss)) < 1e-6
