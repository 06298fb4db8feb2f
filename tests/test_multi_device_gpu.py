"""One process driving two devices (SURVEY.md §8b threading contract): kernels with opt-in shared memory must
be configured per device, workspaces and streams belong to the current device.  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_second_device_runs_every_kernel_family():
    import rpst
    c, s = R.synth_features((2, 4, 256, 256), cfg=70)            # TMA AdaIN kernel (opt-in shared memory)
    cl = R.synth_labels(2, 256, 256, classes=6, block=16, seed=4200)
    sl = R.synth_labels(2, 256, 256, classes=6, block=16, seed=5200)
    f, k, v = (torch.randn(1, 32, 12, 12, generator=torch.Generator().manual_seed(i)) * 0.5 for i in (1, 2, 3))
    want_adain = R.adain(c, s, dtype=torch.float64)
    want_seg = R.seg_adain_batch(c, s, cl, sl, dtype=torch.float64)
    want_attn = R.attention_core(f.reshape(1, 32, -1).double(), k.reshape(1, 32, -1).double(), v.reshape(1, 32, -1).double())
    want_wct = R.wct_fuse(c[:, :, :32, :32], s[:, :, :32, :32])
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            got = rpst.adaptive_instance_normalization(c.to(dev), s.to(dev))
            assert got.device == torch.device(dev) and R.rel_l2(got, want_adain) < 2e-6
            assert R.rel_l2(rpst.seg_adain_batch(c.to(dev), s.to(dev), cl.to(dev), sl.to(dev)), want_seg) < 2e-6
            assert R.rel_l2(rpst.attention_core(f.to(dev), k.to(dev), v.to(dev)).reshape(1, 32, -1), want_attn) < 1e-3
            assert R.rel_l2(rpst.wct_fuse(c[:, :, :32, :32].contiguous().to(dev), s[:, :, :32, :32].contiguous().to(dev)), want_wct) < 1e-3
            assert float(rpst.calc_style_loss(c.to(dev), s.to(dev))) > 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_tensors_on_a_non_current_device_need_no_device_context():
    """a model on cuda:1 while cuda:0 is current must work like the reference's torch ops do."""
    import rpst
    torch.cuda.set_device(0)
    c, s = R.synth_features((2, 4, 256, 256), cfg=71)
    want = R.adain(c, s, dtype=torch.float64)
    got = rpst.adaptive_instance_normalization(c.to("cuda:1"), s.to("cuda:1"))
    assert got.device == torch.device("cuda:1") and R.rel_l2(got, want) < 2e-6
    assert torch.cuda.current_device() == 0
    small = rpst.adaptive_instance_normalization(c[:, :, :32, :32].contiguous().to("cuda:1"), s[:, :, :32, :32].contiguous().to("cuda:1"))
    assert R.rel_l2(small, R.adain(c[:, :, :32, :32], s[:, :, :32, :32], dtype=torch.float64)) < 2e-6
    with pytest.raises(RuntimeError, match="different devices"):
        rpst.adaptive_instance_normalization(c.to("cuda:0"), s.to("cuda:1"))
