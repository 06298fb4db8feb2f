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
from .functional import EPS, _prep, _stream, _workspace, device_guard


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
        g = g.to(torch.float32).reshape(1).contiguous()
        L = _lib.lib()
        grads = [None, None]
        for wrt, t in enumerate((x, y)):
            if ctx.needs_input_grad[wrt]:
                out = torch.empty_like(t)
                _lib.check(L.rpst_pair_loss_bwd(x.data_ptr(), y.data_ptr(), st.data_ptr(), g.data_ptr(), ctx.which, wrt,
                                                out.data_ptr(), planes, hw, _stream()))
                grads[wrt] = out
        return grads[0], grads[1], None


def _pair(input: torch.Tensor, target: torch.Tensor, which: int) -> torch.Tensor:
    assert input.dim() == 4 and target.dim() == 4
    assert (input.size() == target.size())
    x, y = _prep(input, "input"), _prep(target, "target")
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        return _PairLossFn.apply(x, y, which)
    losses, _ = _pair_stats_raw(x, y, EPS, False)
    return losses[which]


@device_guard
def calc_style_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Drop-in for the `calc_style_loss` methods (network/adain_rp.py:84-88)."""
    return _pair(input, target, 0)


@device_guard
def calc_content_loss(input: torch.Tensor, target: torch.Tensor, norm: bool = False) -> torch.Tensor:
    """Drop-in for `calc_content_loss` (network/sanet.py:226-230).  `norm=False` is a plain MSE with
    no statistics in it (network/adain_rp.py:81-82) and stays torch's own `mse_loss`."""
    if not norm:
        return torch.nn.functional.mse_loss(input, target)
    return _pair(input, target, 1)


@device_guard
def pair_statistics(input: torch.Tensor, target: torch.Tensor):
    """(style_loss, content_norm_loss, stats[N*C, 8]) from one pass; see include/rpst.h."""
    x, y = _prep(input, "input"), _prep(target, "target")
    assert x.shape == y.shape and x.dim() == 4
    losses, stats = _pair_stats_raw(x, y, EPS, True)
    return losses[0], losses[1], stats
