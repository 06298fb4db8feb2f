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


# ---- round 2: Newton-Schulz matrix roots (csrc/nsroot.cu) with the Jacobi solver for flagged matrices ----------------
def _mixed(n, c, h, w, seed, gain=1.0, offset=0.5):
    x = torch.relu(torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(seed)) * gain + offset)
    mix = torch.randn(c, c, generator=torch.Generator().manual_seed(seed + 100)) / c ** 0.5
    return torch.einsum("oc,nchw->nohw", mix, x)


@pytest.mark.parametrize("n,c,h,w,hs,ws", [(2, 128, 32, 32, 32, 32),      # full rank
                                            (1, 200, 31, 5, 33, 6),        # H*W < C: rank-deficient, eigenvalues at the 1e-4 floor
                                            (1, 256, 10, 20, 12, 20),      # condition number ~1e6 on the style side
                                            (2, 129, 16, 16, 16, 16),      # odd order: register-staged GEMM path
                                            (1, 512, 24, 24, 24, 24)])
def test_newton_schulz_roots_match_jacobi_and_oracle(rpst, n, c, h, w, hs, ws):
    """network/wct_rp.py:7-38: the iteration computes the same (A + 1e-4 I)^(+-1/2) as the eigensolver whenever the matrix
    is positive definite as constructed; knob 2 flags every matrix, which must reproduce the Jacobi-only path bit for bit."""
    ct, st = _mixed(n, c, h, w, 41), _mixed(n, c, hs, ws, 42, gain=2.0, offset=1.0)
    for method in ("closed-form", "original"):
        res = {}
        try:
            for knob in (0, 1, 2):
                rpst.set_tuning("wct_roots_ns", knob)
                before = rpst.get_tuning("wct_ns_flagged")
                res[knob] = rpst.wct_fuse(ct.cuda(), st.cuda(), method, return_transform=True)
                if knob == 1:
                    assert rpst.get_tuning("wct_ns_flagged") == before, "a positive definite matrix was flagged"
        finally:
            rpst.set_tuning("wct_roots_ns", 1)
        # both solvers see the same fp32-grade covariance; at the 1e-4 eigenvalue floor d sqrt(x) / dx = 50 amplifies their
        # own 1e-8-level differences, hence 5e-5 and not 1e-9 (the contract against the reference is 1e-3, below)
        assert R.rel_l2(res[1][1], res[0][1]) < 5e-5, (method, R.rel_l2(res[1][1], res[0][1]))
        assert torch.equal(res[2][1], res[0][1]) and torch.equal(res[2][0], res[0][0])
        assert R.rel_l2(res[1][0], R.wct_fuse(ct, st, method)) < 1e-3


def _spectrum_matrix(ev, seed):
    n = len(ev)
    q, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, generator=torch.Generator().manual_seed(seed)))
    a = (q * torch.as_tensor(ev, dtype=torch.float64)) @ q.t()
    return (a + a.t()) / 2, q


def test_spd_roots_against_known_spectra(rpst):
    """`rpst_spd_roots` on matrices with prescribed spectra: clustered at the floor (rank-deficient covariance + 1e-4 I),
    condition number 1e8, odd order; and the acceptance flag on an indefinite matrix and on a violated eigenvalue bound."""
    import numpy as np
    rng = np.random.default_rng(5)
    cases = []
    ev = np.exp(rng.uniform(np.log(1e-4), np.log(1e3), 256)); ev[:64] = 1e-4
    cases.append(ev)                                                       # clustered floor, kappa 1e7
    cases.append(np.exp(rng.uniform(np.log(1e-4), np.log(1e4), 255)))      # odd order, kappa 1e8
    cases.append(1.0 + np.exp(rng.uniform(np.log(1e-3), np.log(50.0), 128)))   # content side: eigenvalues >= 1
    for i, ev in enumerate(cases):
        a, q = _spectrum_matrix(ev, 60 + i)
        lmin = float(ev.min())
        root, iroot, flags = rpst.spd_roots(a.cuda(), lmin=lmin, diag_add=0.0)
        assert int(flags.sum()) == 0
        want_r = (q * torch.as_tensor(ev).sqrt()) @ q.t()
        want_i = (q / torch.as_tensor(ev).sqrt()) @ q.t()
        assert R.rel_l2(root, want_r) < 1e-9, (i, R.rel_l2(root, want_r))
        assert R.rel_l2(iroot, want_i) < 1e-6, (i, R.rel_l2(iroot, want_i))    # kappa * eps
    # batch with one indefinite member and one whose smallest eigenvalue is 1000x below the stated bound
    good, _ = _spectrum_matrix(np.linspace(0.5, 3.0, 96), 70)
    ev_bad = np.linspace(0.5, 3.0, 96); ev_bad[0] = -0.2
    bad, _ = _spectrum_matrix(ev_bad, 71)
    ev_low = np.linspace(0.5, 3.0, 96); ev_low[0] = 1e-7
    low, _ = _spectrum_matrix(ev_low, 72)
    batch = torch.stack([good, bad, low, good]).cuda()
    before = rpst.get_tuning("wct_ns_flagged")
    root, iroot, flags = rpst.spd_roots(batch, lmin=1e-4, diag_add=0.0)
    assert flags.tolist() == [0, 1, 1, 0]
    assert rpst.get_tuning("wct_ns_flagged") - before == 2
    assert R.rel_l2(root[0] @ root[0], good) < 1e-12 and torch.equal(root[0], root[3])


def test_newton_schulz_flags_indefinite_input_and_jacobi_takes_over(rpst):
    """Huge variances on a rank-deficient style map: the fp32-grade covariance has eigenvalues below -1e-4 after rounding,
    the iteration cannot converge, the matrix is flagged on the device and solved by the Jacobi path (|s| semantics of
    torch.svd in network/wct_rp.py:11) — no host synchronisation, the result stays finite and matches the Jacobi-only run."""
    c = 160
    ct = _mixed(1, c, 12, 12, 51)
    st = _mixed(1, c, 10, 12, 52, gain=3000.0, offset=1000.0)          # H*W = 120 < C, variance ~1e7
    try:
        rpst.set_tuning("wct_roots_ns", 0)
        want, tw = rpst.wct_fuse(ct.cuda(), st.cuda(), "original", return_transform=True)
    finally:
        rpst.set_tuning("wct_roots_ns", 1)
    before = rpst.get_tuning("wct_ns_flagged")
    got, tg = rpst.wct_fuse(ct.cuda(), st.cuda(), "original", return_transform=True)
    flagged = rpst.get_tuning("wct_ns_flagged") - before
    assert torch.isfinite(got).all()
    assert R.rel_l2(tg, tw) < 5e-5, R.rel_l2(tg, tw)
    # whether rounding really pushed an eigenvalue below the bound depends on the data; either way the result agrees with
    # the eigensolver (asserted above); the flag logic itself is pinned by test_spd_roots_against_known_spectra
    assert flagged in (0, 1)


def test_wct_is_deterministic_and_safe_on_two_streams(rpst):
    """Every reduction in the WCT path has a fixed order (per-CTA partial Grams, column sums, Newton-Schulz products), so
    repeated calls are bit-identical; the covariance is a cooperative launch with grid barriers — two streams issuing it
    concurrently must serialise, not deadlock, and give the same bits."""
    c, s = R.synth_features((2, 256, 128, 128), cfg=3, device="cuda")
    ref = rpst.wct_fuse(c, s)
    for _ in range(10):
        assert torch.equal(rpst.wct_fuse(c, s), ref)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(5):
                o = rpst.wct_fuse(c, s)
            outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)


def test_wct_cuda_graph_capture_and_replay(rpst):
    """The whole fuse (memsets, cooperative covariance launches with tensor maps in their parameters, Newton-Schulz launches
    whose step counts live on the device, colouring) only enqueues work on the given stream: it captures into a CUDA graph
    and replays bit-identically (CUDA graphs)."""
    c, s = R.synth_features((2, 256, 128, 128), cfg=3, device="cuda")
    ref = rpst.wct_fuse(c, s)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rpst.wct_fuse(c, s)                      # warm-up outside the capture (kernel attributes, driver entry point)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = rpst.wct_fuse(c, s)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
