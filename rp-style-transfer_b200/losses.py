"""Loss statistics of the training step, one streaming pass per tensor pair.

Reference interfaces mirrored (methods of every *Net class; paths relative to the reference root):
  calc_style_loss(self, input, target)                 network/adain_rp.py:84-88, network/sanet.py:232-236
  calc_content_loss(self, input, target, norm=False)   network/adain_rp.py:81-82, network/sanet.py:226-230

`rpst_pair_stats` reads input and target ONCE (2*E*4 bytes) and returns both
`mse(mean_i, mean_t) + mse(std_i, std_t)` and `mse(mean_variance_norm(i), mean_variance_norm(t))`
without materialising the normalised tensors; the backward pass is one two-input plane-affine kernel."""
from __future__ import annotations

import torch

from . import _lib
from .functional import EPS, _prep, _stream, _workspace


def _pair_stats_raw(x: torch.Tensor, y: torch.Tensor, eps: float, want_stats: bool):
    n, c = x.shape[:2]
    hw = x[0, 0].numel()
    planes = n * c
    losses = torch.empty(2, dtype=torch.float32, device=x.device)
    stats = torch.empty(planes, 8, dtype=torch.float32, device=x.device) if want_stats else None
    L = _lib.lib()
    ws = _workspace(L.rpst_pair_stats_workspace_bytes(planes, hw), x.device)
    _lib.check(L.rpst_pair_stats(x.data_ptr(), y.data_ptr(), planes, hw, eps,
                                 None if stats is None else stats.data_ptr(), losses.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()))
    return losses, stats


def _affine(x, scale, shift):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().rpst_plane_affine(x.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
                                            x.shape[0] * x.shape[1], x[0, 0].numel(), _stream()))
    return out


def _affine2(x, y, ax, ay, b):
    out = torch.empty_like(x)
    _lib.check(_lib.lib().rpst_plane_affine2(x.data_ptr(), y.data_ptr(), ax.data_ptr(), ay.data_ptr(), b.data_ptr(),
                                             out.data_ptr(), x.shape[0] * x.shape[1], x[0, 0].numel(), _stream()))
    return out


class _PairLossFn(torch.autograd.Function):
    """which = 0: style loss, 1: normalised content loss.  Returns a 0-dim tensor."""

    @staticmethod
    def forward(ctx, x, y, which):
        need = x.requires_grad or y.requires_grad
        losses, stats = _pair_stats_raw(x, y, EPS, need)
        if need:
            ctx.save_for_backward(x, y, stats)
        ctx.which = which
        return losses[which].clone()

    @staticmethod
    def backward(ctx, g):
        x, y, st = ctx.saved_tensors
        planes, hw = st.shape[0], x[0, 0].numel()
        mx, sx, my, sy, m2x, m2y, cxy = (st[:, i] for i in range(7))
        g = g.to(torch.float32)
        gx = gy = None
        if ctx.which == 0:
            # d/dx_i [ (mx-my)^2 + (sx-sy)^2 ] / P  with  d mx/dx_i = 1/HW,  d sx/dx_i = (x_i-mx)/((HW-1) sx)
            dm = 2.0 * g * (mx - my) / (planes * hw)
            dsd = 2.0 * g * (sx - sy) / (planes * (hw - 1))
            if ctx.needs_input_grad[0]:
                scale = (dsd / sx).contiguous()
                gx = _affine(x, scale, (dm - scale * mx).contiguous())
            if ctx.needs_input_grad[1]:
                scale = (-dsd / sy).contiguous()
                gy = _affine(y, scale, (-dm - scale * my).contiguous())
        else:
            # L = sum (x^ - y^)^2 / (P HW);  dL/dx_i = s/sd_x (x^_i - y^_i - x^_i K_x),
            # K_x = (M2_x/sd_x^2 - C_xy/(sd_x sd_y)) / (HW-1)   (sum of x^ and of y^ are both zero)
            s = 2.0 * g / (planes * hw)
            cross = cxy / (sx * sy)
            if ctx.needs_input_grad[0]:
                kx = (m2x / (sx * sx) - cross) / (hw - 1)
                ax = s * (1.0 - kx) / (sx * sx)
                ay = -s / (sx * sy)
                gx = _affine2(x, y, ax.contiguous(), ay.contiguous(), (-ax * mx - ay * my).contiguous())
            if ctx.needs_input_grad[1]:
                ky = (m2y / (sy * sy) - cross) / (hw - 1)
                ay = s * (1.0 - ky) / (sy * sy)
                ax = -s / (sx * sy)
                gy = _affine2(y, x, ay.contiguous(), ax.contiguous(), (-ay * my - ax * mx).contiguous())
        return gx, gy, None


def _pair(input: torch.Tensor, target: torch.Tensor, which: int) -> torch.Tensor:
    assert input.dim() == 4 and target.dim() == 4
    assert (input.size() == target.size())
    x, y = _prep(input, "input"), _prep(target, "target")
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        return _PairLossFn.apply(x, y, which)
    losses, _ = _pair_stats_raw(x, y, EPS, False)
    return losses[which]


def calc_style_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Drop-in for the `calc_style_loss` methods (network/adain_rp.py:84-88)."""
    return _pair(input, target, 0)


def calc_content_loss(input: torch.Tensor, target: torch.Tensor, norm: bool = False) -> torch.Tensor:
    """Drop-in for `calc_content_loss` (network/sanet.py:226-230).  `norm=False` is a plain MSE with
    no statistics in it (network/adain_rp.py:81-82) and stays torch's own `mse_loss`."""
    if not norm:
        return torch.nn.functional.mse_loss(input, target)
    return _pair(input, target, 1)


def pair_statistics(input: torch.Tensor, target: torch.Tensor):
    """(style_loss, content_norm_loss, stats[N*C, 8]) from one pass; see include/rpst.h."""
    x, y = _prep(input, "input"), _prep(target, "target")
    assert x.shape == y.shape and x.dim() == 4
    losses, stats = _pair_stats_raw(x, y, EPS, True)
    return losses[0], losses[1], stats
