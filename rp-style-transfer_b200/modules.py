"""Host-side mirrors of the reference's transform MODULES (same constructor signatures, forward
arguments and state-dict names), calling librpst for the transform math."""
from __future__ import annotations

import torch
from torch import nn

from . import functional as F


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
            return x * y.expand_as(x)
        mean, _ = F.calc_mean_std(x)
        y = self.fc(mean.view(b, c)).view(b, c, 1, 1)
        self.attention_map = y
        return F.plane_affine(x, y)
