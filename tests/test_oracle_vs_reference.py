"""Pin the oracle against the real reference functions imported from /root/reference.
Development-container only (skipped where the reference tree is absent, e.g. on the GPU box)."""
import sys

import pytest
import torch

from oracle import restate as R
from oracle.reference_loader import load_reference, reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    load_reference()
    return {n: sys.modules[f"network.{n}"] for n in ("base", "wct_rp", "sanet", "mrf_rp", "attention", "adain_rp")}


@pytest.mark.parametrize("shape", [(1, 3, 5, 7), (2, 8, 16, 12), (1, 4, 33, 31)])
def test_adain_bitwise_shapes(ref, shape):
    c, s = R.synth_features(shape, cfg=1)
    want = ref["base"].adaptive_instance_normalization(c, s)
    assert R.rel_l2(R.adain(c, s), want) <= 1e-6
    mu, sd = ref["base"].calc_mean_std(c)
    got_mu, got_sd = R.plane_stats(c)
    assert R.rel_l2(got_mu, mu) <= 1e-6 and R.rel_l2(got_sd, sd) <= 1e-6


def test_adain_shape_assert(ref):
    c = torch.randn(1, 2, 4, 4)
    with pytest.raises(AssertionError):
        ref["base"].adaptive_instance_normalization(c, c[:, :, :2])
    with pytest.raises(AssertionError):
        R.adain(c, c[:, :, :2])


def test_seg_random(ref):
    base = ref["base"]
    for seed in range(3):
        c, s = R.synth_features((1, 4, 32, 32), cfg=20 + seed)
        cl = R.synth_labels(1, 32, 32, classes=4, block=8, seed=4100 + seed)[0]
        sl = R.synth_labels(1, 32, 32, classes=4, block=8, seed=5100 + seed)[0]
        saved = base.get_segment_and_info
        base.get_segment_and_info = lambda *a: (cl.numpy(), sl.numpy(), *base.compute_label_info(cl.numpy(), sl.numpy()))
        try:
            want = base.adaptive_instance_normalization_with_segment(c, s, None, None)
        finally:
            base.get_segment_and_info = saved
        assert R.rel_l2(R.seg_adain(c, s, cl, sl), want) <= 1e-6


def test_wct_random(ref):
    wct = ref["wct_rp"]
    dummy = object.__new__(wct.WCTRPNet)
    c, s = R.synth_features((2, 12, 16, 16), cfg=30)
    want = wct.WCTRPNet.fuse(dummy, c, s)
    assert R.rel_l2(R.wct_fuse(c, s), want) <= 1e-6


def test_sanet_random(ref):
    sanet = ref["sanet"]
    torch.manual_seed(5)
    m = sanet.SANet(8)
    c, s = R.synth_features((2, 8, 6, 6), cfg=31)
    with torch.no_grad():
        want = m(c, s)
    assert R.rel_l2(R.sanet_forward(c, s, dict(m.state_dict())), want) <= 1e-5


def test_losses_random(ref):
    import types
    stub = types.SimpleNamespace(mse_loss=torch.nn.MSELoss())
    x, y = R.synth_features((2, 5, 17, 13), cfg=33)
    sanet, adain_rp = ref["sanet"], ref["adain_rp"]
    assert R.rel_l2(R.style_loss(x, y), adain_rp.AdaINRPNet.calc_style_loss(stub, x, y)) <= 1e-6
    assert R.rel_l2(R.content_loss(x, y, norm=True), sanet.SAModel.calc_content_loss(stub, x, y, norm=True)) <= 1e-6
    assert R.rel_l2(R.content_loss(x, y), adain_rp.AdaINRPNet.calc_content_loss(stub, x, y)) <= 1e-6
