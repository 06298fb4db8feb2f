"""Host-side mirrors of the reference's transform MODULES (same constructor signatures, forward
arguments and state-dict names), calling librpst for the transform math."""
from __future__ import annotations

import torch
from torch import nn

from . import functional as F
from .losses import _pair_stats_raw


class _GateFn(torch.autograd.Function):
    """out[n,c,:,:] = x[n,c,:,:] * gate[n,c] with both gradients from librpst: dx = g * gate (plane-affine kernel),
    dgate[n,c] = sum_hw g*x = C_gx + HW*mean_g*mean_x (one pass of the pair-moments kernel over (g, x))."""

    @staticmethod
    def forward(ctx, x, gate):
        ctx.save_for_backward(x, gate)
        return F.plane_affine(x, gate)

    @staticmethod
    def backward(ctx, g):
        x, gate = ctx.saved_tensors
        g = g.contiguous()
        dx = F.plane_affine(g, gate) if ctx.needs_input_grad[0] else None
        dgate = None
        if ctx.needs_input_grad[1]:
            _, st = _pair_stats_raw(g, x, F.EPS, True)
            hw = x[0, 0].numel()
            dgate = (st[:, 6] + hw * st[:, 0] * st[:, 2]).view_as(gate)
        return dx, dgate


class SELayer(nn.Module):
    """network/attention.py:5-22.  Global average pool (= per-(n,c) mean, the statistics kernel) ->
    FC/ReLU/FC/Sigmoid -> per-plane scale (plane-affine kernel).  Parameters: fc.0.weight,
    fc.2.weight (bias-free), as in the reference."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)  # kept for state-dict / attribute compatibility
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )
        self.attention_map = None

    def forward(self, x):
        b, c, _, _ = x.size()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training keeps autograd: pooled statistics through the differentiable stats op
            mean, _ = F.calc_mean_std(x) if x.requires_grad else (F.calc_mean_std(x.detach())[0], None)
            y = self.fc(mean.view(b, c)).view(b, c, 1, 1)
            self.attention_map = y
            return _GateFn.apply(F._prep(x, "x"), y.contiguous())
        mean, _ = F.calc_mean_std(x)
        y = self.fc(mean.view(b, c)).view(b, c, 1, 1)
        self.attention_map = y
        return F.plane_affine(x, y)


class CCAMDec(nn.Module):
    """network/adain_rp.py:347-385 (the channel Gram `X Y^T` over H*W is the same dense
    contraction as the WCT covariance).  Both products run on the tcgen05 GEMM block (bf16x3, fp32-grade); the C x K
    softmax in between is O(C*K).  As in the reference, `scale` is a plain zero tensor (`nn.Parameter(...).cuda()` does
    not register), so the shipped module returns its input unchanged; with `scale == 0` the products are skipped."""

    def __init__(self):
        super().__init__()
        self.softmax = nn.Softmax(dim=-1)
        self.scale = torch.zeros(1)

    def forward(self, x, y):
        x, y = x.detach(), y.detach()
        scale = float(self.scale)
        if scale == 0.0:
            return x + 0.0
        return x + scale * ccam_attention(x, y)


def ccam_attention(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """softmax(rowmax(X Y^T) - X Y^T) Y for x [B,C,H,W], y [B,K,H,W] (network/adain_rp.py:370-381)."""
    from .mrf import packed_gemm
    x, y = F._prep(x, "x"), F._prep(y, "y")
    b, c = x.shape[:2]
    k = y.shape[1]
    outs = []
    for i in range(b):
        xr, yr = x[i].reshape(c, -1), y[i].reshape(k, -1)
        energy = packed_gemm(xr, yr)                                     # [C,K]: Gram over H*W
        att = torch.softmax(energy.max(dim=-1, keepdim=True)[0] - energy, dim=-1)
        outs.append(packed_gemm(att, yr.t()).reshape(x.shape[1:]))      # [C,K] x [K,N]
    return torch.stack(outs)
