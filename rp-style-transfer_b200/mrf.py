"""MRF patch matching — host-side mirror of network/base.py:317-360 (`cal_affinity_map`, `cal_dist`)
and network/mrf_rp.py:4-23 (`MRFLoss`).  The reference hard-codes `.cuda()` and N == 1; both hold here
by construction (CUDA only, one sample per call)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from . import _lib
from .functional import _prep, _ptr, _stream


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def packed_gemm(a: torch.Tensor, b: torch.Tensor, passes: int = 3, alpha: float = 1.0) -> torch.Tensor:
    """alpha * a @ b.T for fp32 a [M,K], b [N,K] on tcgen05 (passes=3: bf16x3, fp32-grade)."""
    a, b = _prep(a, "a"), _prep(b, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    m, k = a.shape
    n = b.shape[0]
    L = _lib.lib()
    na, nb_ = L.rpst_packed_operand_bytes(m, k), L.rpst_packed_operand_bytes(n, k)
    ta = [_ws(na, a.device) for _ in range(2)]
    tb = [_ws(nb_, a.device) for _ in range(2)]
    _lib.check(L.rpst_pack_operand(a.data_ptr(), m, k, k, 1, None, ta[0].data_ptr(), ta[1].data_ptr(), _stream()))
    _lib.check(L.rpst_pack_operand(b.data_ptr(), n, k, k, 1, None, tb[0].data_ptr(), tb[1].data_ptr(), _stream()))
    out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    _lib.check(L.rpst_gemm_packed(ta[0].data_ptr(), ta[1].data_ptr(), tb[0].data_ptr(), tb[1].data_ptr(),
                                  out.data_ptr(), m, n, k, n, passes, alpha, None, None, _stream()))
    return out


def cal_dist(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """Drop-in for network/base.py:349 — A (d,m), B (d,n) column vectors -> (m,n) squared distances."""
    A, B = _prep(A, "A"), _prep(B, "B")
    assert A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0]
    d, m = A.shape
    n = B.shape[1]
    L = _lib.lib()
    ws = _ws(L.rpst_pairwise_sqdist_workspace_bytes(d, m, n), A.device)
    out = torch.empty(m, n, dtype=torch.float32, device=A.device)
    _lib.check(L.rpst_pairwise_sqdist(A.data_ptr(), B.data_ptr(), d, m, n, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


def mrf_match(content_feat: torch.Tensor, style_feat: torch.Tensor, k: int = 3, reverse: bool = False,
              want_affinity: bool = False, want_loss: bool = False, mean: str = "mean", precision: str = "fp32"):
    """Top-k matching of one content/style pair.  Returns (idx_dim0 [k,L], idx_dim1 [L,k], affinity|None, loss|None)
    — the two index tensors are exactly `torch.topk(map, k, 0)[1]` / `torch.topk(map, k, 1)[1]` of the
    reference's cosine map (network/base.py:338-343)."""
    assert content_feat.size() == style_feat.size()
    n, c, h, w = content_feat.size()
    assert n == 1, "MRF matching is defined for a single sample (the reference squeezes the batch axis)"
    cf = _prep(content_feat, "content_feat").view(c, h * w)
    sf = _prep(style_feat, "style_feat").view(c, h * w)
    l = h * w
    dev = cf.device
    idx0 = torch.empty(k, l, dtype=torch.int64, device=dev)
    idx1 = torch.empty(l, k, dtype=torch.int64, device=dev)
    aff = torch.empty(l, l, dtype=torch.float32, device=dev) if want_affinity else None
    loss = torch.empty(1, dtype=torch.float32, device=dev) if want_loss else None
    L = _lib.lib()
    ws = _ws(L.rpst_mrf_workspace_bytes(c, l, k), dev)
    passes = {"fp32": 3, "bf16": 1}[precision]
    _lib.check(L.rpst_mrf_match(cf.data_ptr(), sf.data_ptr(), c, l, int(k), int(bool(reverse)), passes,
                                idx0.data_ptr(), idx1.data_ptr(), _ptr(aff), _ptr(loss), int(mean != "mean"),
                                ws.data_ptr(), ws.numel(), _stream()))
    return idx0, idx1, aff, (loss[0] if loss is not None else None)


def cal_affinity_map(content_feat, style_feat, k=3, reverse=False, c_mask=None, s_mask=None) -> torch.Tensor:
    """Drop-in for network/base.py:317 — the dense binary [HW,HW] affinity map."""
    return mrf_match(content_feat, style_feat, k, reverse, want_affinity=True)[2]


class _MRFLossFn(torch.autograd.Function):
    """loss = sum(A o dist) / norm with the (non-differentiable) top-k affinity A from librpst; like the
    reference, the gradient flows through `cal_dist` alone: d/da_i = 2/norm * sum_j A_ij (a_i - b_j)."""

    @staticmethod
    def forward(ctx, content_feat, style_feat, k, mean):
        _, _, aff, loss = mrf_match(content_feat, style_feat, k, want_affinity=True, want_loss=True, mean=mean)
        ctx.save_for_backward(content_feat, style_feat, aff)
        n, c, h, w = content_feat.shape
        ctx.norm = float(h * w * k) if mean == "mean" else float(h * w) ** 2
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        cf, sf, aff = ctx.saved_tensors
        c = cf.shape[1]
        a, b = cf.reshape(c, -1), sf.reshape(c, -1)
        s = 2.0 * g / ctx.norm
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = (s * (a * aff.sum(1)[None, :] - b @ aff.t())).view_as(cf)
        if ctx.needs_input_grad[1]:
            gb = (s * (b * aff.sum(0)[None, :] - a @ aff)).view_as(sf)
        return ga, gb, None, None


class MRFLoss(nn.Module):
    """network/mrf_rp.py:4-23.  Matching, distances and the loss are one librpst call; under autograd the
    gradient flows through `cal_dist` alone, as in the reference (the top-k affinity is piecewise constant)."""

    def __init__(self, k, mask=None, mean='mean') -> None:
        super().__init__()
        self.mask = mask
        self.k = k
        self.mean = mean

    def forward(self, content_feat, style_feat):
        if torch.is_grad_enabled() and (content_feat.requires_grad or style_feat.requires_grad):
            return _MRFLossFn.apply(_prep(content_feat, "content_feat"), _prep(style_feat, "style_feat"), self.k, self.mean)
        return mrf_match(content_feat, style_feat, self.k, want_loss=True, mean=self.mean)[3]
