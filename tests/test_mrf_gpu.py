"""GPU parity tests: tcgen05 GEMM building block and MRF patch matching.
MRF patch indices must be bit-exact on tie-free inputs (BASELINE.json north_star)."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rpst():
    import rpst as m
    return m


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 384, 512), (200, 100, 70), (1, 5, 3), (4096, 256, 192)])
def test_packed_gemm_bf16x3_is_fp32_grade(rpst, m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g)
    want = a.double() @ b.double().t()
    got3 = rpst.packed_gemm(a.cuda(), b.cuda(), passes=3)
    got1 = rpst.packed_gemm(a.cuda(), b.cuda(), passes=1, alpha=-2.0)
    assert R.rel_l2(got3, want) < 2e-5          # ~2^-16: fp32-grade
    assert R.rel_l2(got1, -2.0 * want) < 1e-2   # plain bf16 operands


def test_cal_dist_golden_and_random(rpst, golden):
    g = golden("mrf")
    a = g["content"].reshape(16, -1).cuda()
    b = g["style"].reshape(16, -1).cuda()
    assert R.rel_l2(rpst.cal_dist(a, b), g["dist"]) < 1e-5
    x = torch.randn(48, 300).cuda()
    y = torch.randn(48, 170).cuda()
    assert R.rel_l2(rpst.cal_dist(x, y), R.pairwise_sqdist(x.cpu().double(), y.cpu().double())) < 1e-5


def test_golden_indices_are_bit_exact(rpst, golden):
    g = golden("mrf")
    k = int(g["k"])
    idx0, idx1, aff, loss = rpst.mrf_match(g["content"].cuda(), g["style"].cuda(), k, want_affinity=True, want_loss=True)
    assert torch.equal(idx0.cpu(), g["idx0"]) and torch.equal(idx1.cpu(), g["idx1"])
    assert torch.equal(aff.cpu(), g["affinity"])
    assert R.rel_l2(loss, g["loss"]) < 1e-5
    assert torch.equal(rpst.cal_affinity_map(g["content"].cuda(), g["style"].cuda(), k).cpu(), g["affinity"])
    m = rpst.MRFLoss(k, mean="all")
    assert R.rel_l2(m(g["content"].cuda(), g["style"].cuda()), g["loss_all"]) < 1e-5


def _tie_free_gap(ncc, k):
    top1 = torch.topk(ncc, k + 1, dim=1).values
    top0 = torch.topk(ncc, k + 1, dim=0).values
    return min(float((top1[:, k - 1] - top1[:, k]).min()), float((top0[k - 1] - top0[k]).min()))


@pytest.mark.parametrize("c,h,w,k,reverse", [(16, 10, 12, 3, False), (64, 32, 32, 5, False), (40, 17, 9, 2, True)])
def test_vs_oracle_small(rpst, c, h, w, k, reverse):
    ct, st = R.synth_features((1, c, h, w), cfg=6)
    idx0_w, idx1_w, ncc = R.mrf_topk_indices(ct, st, k, reverse, dtype=torch.float64)
    assert _tie_free_gap(ncc, k) > 1e-6
    idx0, idx1, aff, loss = rpst.mrf_match(ct.cuda(), st.cuda(), k, reverse, want_affinity=True, want_loss=True)
    assert torch.equal(idx0.cpu(), idx0_w) and torch.equal(idx1.cpu(), idx1_w)
    want_aff = torch.zeros(h * w, h * w, dtype=torch.float64)
    want_aff.scatter_(0, idx0_w, 1.0)
    want_aff.scatter_(1, idx1_w, 1.0)
    assert torch.equal(aff.cpu().double(), want_aff)
    dist = R.pairwise_sqdist(ct.reshape(c, -1).double(), st.reshape(c, -1).double())
    assert R.rel_l2(loss, (want_aff * dist).sum() / (h * w * k)) < 1e-5


def _check_topk(got_idx, ncc, k, dim, eps=2e-6):
    """Indices must equal the fp64 top-k except where the fp64 scores are closer than `eps` (near-ties:
    16.7M random cosines always contain some); a deviating pick must still score within eps of the
    entry it replaced, and every list must be sorted."""
    want = torch.topk(ncc, k, dim=dim)
    got_val = torch.gather(ncc, dim, got_idx)
    exact = (got_idx == want.indices)
    assert float(exact.double().mean()) > 0.999
    assert float((want.values - got_val).abs().max()) < eps
    d = got_val.narrow(dim, 0, k - 1) - got_val.narrow(dim, 1, k - 1)
    assert float(d.min()) > -eps
    return float(exact.double().mean())


def test_full_size_relu4_1(rpst):
    """C=512, L=4096 (relu4_1 @512^2), k=5 (config['k']) against an fp64 cosine map computed on the GPU
    from the same inputs; loss against the sparse oracle evaluated on the kernel's own index sets."""
    c, h, w, k = 512, 64, 64, 5
    ct, st = R.synth_features((1, c, h, w), cfg=6, device="cuda")
    a = ct.view(c, -1).double()
    b = st.view(c, -1).double()
    an = a / a.norm(dim=0, keepdim=True).clamp_min(1e-12)
    bn = b / b.norm(dim=0, keepdim=True).clamp_min(1e-12)
    ncc = an.t() @ bn
    idx0, idx1, _, loss = rpst.mrf_match(ct, st, k, want_loss=True)
    _check_topk(idx1, ncc, k, 1)
    _check_topk(idx0, ncc, k, 0)
    # rows/columns whose k-th and (k+1)-th scores are separated by > 1e-5 must be bit-exact
    top1 = torch.topk(ncc, k + 1, dim=1)
    safe = (top1.values[:, :k] - top1.values[:, 1:k + 1]).min(dim=1).values > 1e-5
    assert torch.equal(idx1[safe], top1.indices[safe][:, :k]) and int(safe.sum()) > 3000
    want = R.mrf_loss_from_indices(ct.cpu(), st.cpu(), idx0.cpu(), idx1.cpu(), k, dtype=torch.float64)
    assert R.rel_l2(loss, want) < 1e-5


def test_k_out_of_range_is_rejected(rpst):
    x = torch.zeros(1, 4, 4, 4).cuda()
    with pytest.raises(rpst.RpstError):
        rpst.mrf_match(x, x, k=9)


def test_mrf_loss_gradient_flows_through_the_distances(rpst):
    """network/mrf_rp.py:12-23 under autograd: the affinity map is piecewise constant, so the gradient is that of
    sum(A o cal_dist) with A fixed; compared with fp64 autograd of exactly that expression (A from the oracle)."""
    c, s = R.synth_features((1, 32, 12, 12), cfg=6, signed=True)
    k = 3
    aff = R.mrf_affinity_map(c, s, k).double()
    cd, sd = c.double().requires_grad_(), s.double().requires_grad_()
    a, b = cd.view(32, -1), sd.view(32, -1)
    dist = (a * a).sum(0)[:, None] + (b * b).sum(0)[None, :] - 2 * a.t() @ b
    want = (aff * dist).sum() / (144 * k)
    (want * 1.7).backward()
    cg, sg = c.cuda().requires_grad_(), s.cuda().requires_grad_()
    got = rpst.MRFLoss(k)(cg, sg)
    (got * 1.7).backward()
    assert abs(float(got.detach()) - float(want.detach())) / abs(float(want.detach())) < 1e-4
    assert R.rel_l2(cg.grad, cd.grad) < 1e-4 and R.rel_l2(sg.grad, sd.grad) < 1e-4


def test_cal_dist_is_differentiable(rpst):
    """the reference's cal_dist is plain torch ops; the drop-in must not detach silently."""
    g = torch.Generator().manual_seed(21)
    a = torch.randn(24, 70, generator=g); b = torch.randn(24, 45, generator=g); w = torch.randn(70, 45, generator=g)
    ag, bg = a.cuda().requires_grad_(), b.cuda().requires_grad_()
    ga, gb = torch.autograd.grad(rpst.cal_dist(ag, bg), (ag, bg), w.cuda())
    a64, b64 = a.double().requires_grad_(), b.double().requires_grad_()
    dist = (a64 * a64).sum(0)[:, None] + (b64 * b64).sum(0)[None, :] - 2.0 * a64.t() @ b64   # network/base.py:349-360
    ra, rb = torch.autograd.grad(dist, (a64, b64), w.double())
    assert R.rel_l2(ga, ra) < 1e-4 and R.rel_l2(gb, rb) < 1e-4


def test_ccam_gram_attention_vs_oracle(rpst):
    """CCAM's channel Gram + attention on the tcgen05 GEMM block (network/adain_rp.py:358-385)."""
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 32, 24, 20, generator=g) * 0.05
    y = torch.randn(2, 32, 24, 20, generator=g) * 0.05
    m = rpst.CCAMDec()
    assert torch.equal(m(x.cuda(), y.cuda()).cpu(), x)                     # as shipped: scale == 0 -> identity
    m.scale = torch.tensor([0.7])
    want = R.ccam(x, y, scale=0.7, dtype=torch.float64)
    assert R.rel_l2(m(x.cuda(), y.cuda()), want) < 1e-4
