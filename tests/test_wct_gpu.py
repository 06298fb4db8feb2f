"""GPU parity tests for WCT: Jacobi matrix functions and whitening/colouring."""
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rpst():
    import rpst as m
    return m


def test_matfn_golden(rpst, golden):
    g = golden("wct_matfn")
    a = g["a"].cuda()
    assert R.rel_l2(rpst.matrix_sqrt(a), g["sqrt"]) < 1e-10
    assert R.rel_l2(rpst.matrix_inv_sqrt(a), g["inv_sqrt"]) < 1e-10


@pytest.mark.parametrize("n,batch", [(8, 1), (32, 3), (48, 2), (64, 2), (128, 2), (256, 3), (512, 1)])
def test_matfn_random_spd(rpst, n, batch):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(batch, n, 3 * n, generator=g, dtype=torch.float64)
    x[:, : n // 4] *= 30.0            # spread the spectrum
    a = x @ x.transpose(1, 2) / (3 * n - 1)
    rs, ri = rpst.wct._sym_fn(a.cuda(), True, True)
    for b in range(batch):
        assert R.rel_l2(rs[b], R.matrix_sqrt(a[b])) < 1e-9
        assert R.rel_l2(ri[b], R.matrix_inv_sqrt(a[b])) < 1e-9
    # defining property: sqrt @ sqrt == A + 1e-4 I
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert R.rel_l2(rs[0] @ rs[0], a[0].cuda() + 1e-4 * eye) < 1e-10


def test_matfn_rank_deficient_is_truncation_safe(rpst):
    # rank-4 PSD matrix: after +1e-4 every eigenvalue is >= 1e-4 > 1e-5, nothing is cut
    g = torch.Generator().manual_seed(0)
    x = torch.randn(24, 4, generator=g, dtype=torch.float64)
    a = x @ x.t()
    assert R.rel_l2(rpst.matrix_inv_sqrt(a.cuda()), R.matrix_inv_sqrt(a)) < 1e-8


def test_wct_golden(rpst, golden):
    g = golden("wct")
    out = rpst.wct_fuse(g["content"].cuda(), g["style"].cuda())
    assert R.rel_l2(out, g["fuse"]) < 1e-3
    cf = g["content"][0].reshape(16, -1).double().cuda()
    sf = g["style"][0].reshape(16, -1).double().cuda()
    assert R.rel_l2(rpst.whiten_and_color(cf, sf), g["closed_form"]) < 1e-3
    assert R.rel_l2(rpst.whiten_and_color(cf, sf, method="original"), g["original"]) < 1e-3


@pytest.mark.parametrize("n,c,h,w,hs,ws", [(2, 16, 20, 20, 20, 20), (1, 64, 48, 40, 30, 52), (2, 128, 32, 32, 32, 32),
                                            (1, 256, 64, 64, 64, 64), (1, 512, 32, 32, 32, 32)])
def test_wct_vs_oracle(rpst, n, c, h, w, hs, ws):
    ct = torch.relu(torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(1)) + 0.5)
    st = torch.relu(torch.randn(n, c, hs, ws, generator=torch.Generator().manual_seed(2)) * 2 + 1)
    # correlated channels (a random mixing) so the covariances are far from diagonal
    mix = torch.randn(c, c, generator=torch.Generator().manual_seed(3)) / c ** 0.5
    ct = torch.einsum("oc,nchw->nohw", mix, ct)
    st = torch.einsum("oc,nchw->nohw", mix.t(), st)
    for method in ("closed-form", "original"):
        want = R.wct_fuse(ct, st, method)
        got, tr = rpst.wct_fuse(ct.cuda(), st.cuda(), method, return_transform=True)
        assert R.rel_l2(got, want) < 1e-3, method
        # output statistics: the coloured features carry the style means exactly
        assert R.rel_l2(got.mean(dim=(2, 3)), st.mean(dim=(2, 3))) < 1e-3


def test_wct_full_plane_properties(rpst):
    """config #3 shape per sample (256 x 512x512): with method 'original' cov(out) == cov(style)+1e-4-ish,
    checked through the per-channel variance and mean; one sample, bounded runtime."""
    c, h, w = 256, 512, 512
    ct, st = R.synth_features((1, c, h, w), cfg=3, device="cuda")
    out = rpst.wct_fuse(ct, st)
    assert torch.isfinite(out).all()
    assert R.rel_l2(out.mean(dim=(2, 3)), st.mean(dim=(2, 3))) < 1e-3
    # closed form maps cov_c + I onto cov_s: per-channel variances of out ~ those of style when var_c >> 1 is not
    # guaranteed, so compare against the fp64 reference on a channel subset instead
    xo = out[0].reshape(c, -1).double()
    cov_o = torch.cov(xo)
    xs = st[0].reshape(c, -1).double()
    xc = ct[0].reshape(c, -1).double()
    cov_c = torch.cov(xc) + torch.eye(c, dtype=torch.float64, device="cuda")
    tr = rpst.wct_fuse(ct, st, return_transform=True)[1][0]
    want_cov_o = tr @ (cov_c - torch.eye(c, dtype=torch.float64, device="cuda")) @ tr.t()
    assert R.rel_l2(cov_o, want_cov_o) < 1e-3
    # T (C_c) T^T == C_s + 1e-4-regularised root product: defining identity of the closed form
    assert R.rel_l2(tr @ cov_c @ tr.t(), torch.cov(xs)) < 5e-3


# ---- round 2: fused convert + centre + SYRK covariance kernel (csrc/cov.cu) -------------------
@pytest.mark.parametrize("n,c,h,w", [(1, 256, 64, 64), (2, 200, 20, 17 * 4), (1, 64, 9, 8), (1, 128, 33, 4), (2, 16, 8, 8)])
def test_fused_covariance_matches_packed_path_and_oracle(rpst, n, c, h, w):
    """The transform matrix T depends on the two covariances only: fused kernel (shifted operands + rank-1
    correction, SYRK symmetry, ragged last k-tile, padded channel rows) vs the pack + GEMM path vs the fp64 oracle."""
    ct = torch.relu(torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(11)) + 0.5)
    st = torch.relu(torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(12)) * 2 + 1)
    mix = torch.randn(c, c, generator=torch.Generator().manual_seed(13)) / c ** 0.5
    ct = torch.einsum("oc,nchw->nohw", mix, ct)
    st = torch.einsum("oc,nchw->nohw", mix.t(), st)
    try:
        rpst.set_tuning("wct_fused_cov", 0)
        rpst.set_tuning("wct_fused_apply", 0)
        out0, t0 = rpst.wct_fuse(ct.cuda(), st.cuda(), return_transform=True)
    finally:
        rpst.set_tuning("wct_fused_cov", 1)
        rpst.set_tuning("wct_fused_apply", 1)
    out1, t1 = rpst.wct_fuse(ct.cuda(), st.cuda(), return_transform=True)
    want = R.wct_fuse(ct, st)
    assert R.rel_l2(out1, want) < 1e-3, R.rel_l2(out1, want)
    assert R.rel_l2(out0, want) < 1e-3
    assert R.rel_l2(t1, t0) < 2e-4, R.rel_l2(t1, t0)
    assert R.rel_l2(out1, out0) < 2e-4, R.rel_l2(out1, out0)      # fused apply (wct_apply.cu) vs pack + GEMM
    out16 = rpst.wct_fuse(ct.cuda(), st.cuda(), precision="bf16")
    assert R.rel_l2(out16, want) < 1e-2
    assert R.rel_l2(out1.mean(dim=(2, 3)), st.mean(dim=(2, 3))) < 1e-3


def test_fused_covariance_with_mean_far_from_zero(rpst):
    """|mean| >> std: the sub-sampled shift must keep the rank-1 centring correction free of cancellation."""
    n, c, h, w = 1, 64, 64, 64
    g = torch.Generator().manual_seed(21)
    ct = torch.randn(n, c, h, w, generator=g) * 0.05 + 40.0 + torch.arange(c).view(1, c, 1, 1) * 3.0
    st = torch.randn(n, c, h, w, generator=g) * 0.2 - 25.0
    want = R.wct_fuse(ct, st)
    got = rpst.wct_fuse(ct.cuda(), st.cuda())
    assert R.rel_l2(got, want) < 1e-3, R.rel_l2(got, want)


@pytest.mark.parametrize("c,h,w", [(64, 7, 9), (200, 31, 5), (256, 33, 31), (3, 5, 5)])
def test_fused_apply_ragged_shapes(rpst, c, h, w):
    """odd H*W (scalar loads, ragged last position tile) and channel counts that are not tile multiples"""
    ct = torch.relu(torch.randn(1, c, h, w, generator=torch.Generator().manual_seed(31)) + 0.5)
    st = torch.relu(torch.randn(1, c, h + 2, w + 1, generator=torch.Generator().manual_seed(32)) * 2 + 1)
    want = R.wct_fuse(ct, st)
    got = rpst.wct_fuse(ct.cuda(), st.cuda())
    assert R.rel_l2(got, want) < 1e-3, R.rel_l2(got, want)
