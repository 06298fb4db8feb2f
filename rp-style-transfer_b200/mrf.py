"""MRF patch matching — host-side mirror of network/base.py:317-360 (`cal_affinity_map`, `cal_dist`)
and network/mrf_rp.py:4-23 (`MRFLoss`).  The reference hard-codes `.cuda()` and N == 1; both hold here
by construction (CUDA only, one sample per call)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from . import _lib
from .functional import _prep, _ptr, _stream, device_guard


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


@device_guard
def packed_gemm(a: torch.Tensor, b: torch.Tensor, passes: int = 3, alpha: float = 1.0,
                row_add: torch.Tensor = None, col_add: torch.Tensor = None) -> torch.Tensor:
    """alpha * a @ b.T (+ row_add[:, None] + col_add[None, :]) for fp32 a [M,K], b [N,K] on tcgen05
    (passes=3: bf16x3, fp32-grade; 1: plain bf16).  `a` and `b` may be arbitrary 2-D strided views (a transposed
    view costs nothing: the pack kernel reads through the strides)."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    for t, name in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.float32:
            raise TypeError(f"rpst.packed_gemm: `{name}` must be a CUDA float32 tensor")
    m, k = a.shape
    n = b.shape[0]
    L = _lib.lib()
    na, nb_ = L.rpst_packed_operand_bytes(m, k), L.rpst_packed_operand_bytes(n, k)
    lo = passes == 3
    ta = [_ws(na, a.device) for _ in range(2 if lo else 1)]
    tb = [_ws(nb_, a.device) for _ in range(2 if lo else 1)]
    _lib.check(L.rpst_pack_operand(a.data_ptr(), m, k, a.stride(0), a.stride(1), None, ta[0].data_ptr(),
                                   ta[1].data_ptr() if lo else None, _stream()))
    _lib.check(L.rpst_pack_operand(b.data_ptr(), n, k, b.stride(0), b.stride(1), None, tb[0].data_ptr(),
                                   tb[1].data_ptr() if lo else None, _stream()))
    out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ra = None if row_add is None else _prep(row_add.reshape(-1), "row_add")
    ca = None if col_add is None else _prep(col_add.reshape(-1), "col_add")
    assert ra is None or ra.numel() == m
    assert ca is None or ca.numel() == n
    _lib.check(L.rpst_gemm_packed(ta[0].data_ptr(), ta[1].data_ptr() if lo else None, tb[0].data_ptr(),
                                  tb[1].data_ptr() if lo else None, out.data_ptr(), m, n, k, n, passes, alpha,
                                  _ptr(ra), _ptr(ca), _stream()))
    return out


def _cal_dist_raw(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    d, m = A.shape
    n = B.shape[1]
    L = _lib.lib()
    ws = _ws(L.rpst_pairwise_sqdist_workspace_bytes(d, m, n), A.device)
    out = torch.empty(m, n, dtype=torch.float32, device=A.device)
    _lib.check(L.rpst_pairwise_sqdist(A.data_ptr(), B.data_ptr(), d, m, n, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


class _CalDistFn(torch.autograd.Function):
    """dist[i,j] = |a_i|^2 + |b_j|^2 - 2 a_i.b_j  =>  dA = 2 (A o rowsum(g) - B g^T), dB = 2 (B o colsum(g) - A g);
    both products run on the tcgen05 GEMM block."""

    @staticmethod
    def forward(ctx, A, B):
        ctx.save_for_backward(A, B)
        return _cal_dist_raw(A, B)

    @staticmethod
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        g = g.contiguous()
        dA = dB = None
        if ctx.needs_input_grad[0]:
            dA = 2.0 * (A * g.sum(1)[None, :] - packed_gemm(B, g))          # [d,n] x [m,n]^T
        if ctx.needs_input_grad[1]:
            dB = 2.0 * (B * g.sum(0)[None, :] - packed_gemm(A, g.t()))      # [d,m] x [n,m]^T
        return dA, dB


@device_guard
def cal_dist(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """Drop-in for network/base.py:349 — A (d,m), B (d,n) column vectors -> (m,n) squared distances
    (differentiable, like the reference's torch ops)."""
    A, B = _prep(A, "A"), _prep(B, "B")
    assert A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0]
    if torch.is_grad_enabled() and (A.requires_grad or B.requires_grad):
        return _CalDistFn.apply(A, B)
    return _cal_dist_raw(A, B)


@device_guard
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


@device_guard
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
            ga = (s * (a * aff.sum(1)[None, :] - packed_gemm(b, aff))).view_as(cf)        # b [c,L] x aff[i,:]
        if ctx.needs_input_grad[1]:
            gb = (s * (b * aff.sum(0)[None, :] - packed_gemm(a, aff.t()))).view_as(sf)
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
