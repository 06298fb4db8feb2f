"""Pointwise (1x1) convolutions on the tcgen05 GEMM block with fused prologue / epilogue (csrc/pwconv.cu).

SURVEY.md section 8f:
  rank 2  SANet's f / g / h / out_conv projections with `mean_variance_norm` folded into the operand conversion and the
          residual add in the epilogue; Q and K leave the kernel as packed attention operands  (network/sanet.py:82-99)
  rank 4  RP-encoder 1x1 conv + LeakyReLU with the AdaIN statistics emitted from the conv epilogue, and the transform
          that consumes them without a statistics pass (network/base.py:170-198, :399-418)

Inference-side operators: inputs are detached (the training graph keeps cuDNN convolutions and the differentiable
transform ops)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .functional import EPS, _prep, _ptr, _stream, device_guard, plane_affine

PRECISION = {"fp32": 3, "bf16": 1}


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def pack_weight(weight: torch.Tensor, precision: str = "fp32"):
    """[cout, cin(,1,1)] fp32 weight -> packed bf16 hi (+ lo) operand tiles (rows = output channels)."""
    w = _prep(weight.detach().reshape(weight.shape[0], -1), "weight")
    cout, cin = w.shape
    L = _lib.lib()
    nbytes = L.rpst_packed_operand_bytes(cout, cin)
    hi = _ws(nbytes, w.device)
    lo = _ws(nbytes, w.device) if PRECISION[precision] == 3 else None
    _lib.check(L.rpst_pack_operand(w.data_ptr(), cout, cin, cin, 1, None, hi.data_ptr(), _ptr(lo), _stream()))
    return hi, lo


@device_guard
def conv1x1(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, sub: Optional[torch.Tensor] = None,
            mul: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: Optional[str] = None,
            slope: float = 0.2, want_stats: bool = False, packed: bool = False, precision: str = "fp32", eps: float = EPS):
    """y = act(W ((x - sub) * mul) + bias) (+ residual) for x [b,cin,h,w], W [cout,cin(,1,1)], sub / mul [b,cin(,1,1)].

    Returns y [b,cout,h,w]; with `want_stats` also (mean, std) of y as `calc_mean_std` would compute them; with
    `packed` the result leaves as packed attention operands (hi, lo) instead of fp32 (rpst_sanet_attn_fwd_packed)."""
    assert act in (None, "lrelu")
    x = _prep(x.detach(), "x")
    b, cin = x.shape[:2]
    hw = x[0, 0].numel()
    cout = weight.shape[0]
    w_hi, w_lo = pack_weight(weight, precision)
    prep = lambda t, n: None if t is None else _prep(t.detach().reshape(b, -1), n)
    sub, mul = prep(sub, "sub"), prep(mul, "mul")
    assert sub is None or sub.shape[1] == cin
    assert mul is None or mul.shape[1] == cin
    bias_ = None if bias is None else _prep(bias.detach().reshape(-1), "bias")
    res = None if residual is None else _prep(residual.detach(), "residual")
    assert res is None or res.shape == (b, cout) + tuple(x.shape[2:])
    L = _lib.lib()
    out = out_hi = out_lo = None
    if packed:
        tile_bytes = (L.rpst_packed_operand_bytes(hw, cout) + 255) // 256 * 256
        out_hi = _ws(b * tile_bytes, x.device)
        out_lo = _ws(b * tile_bytes, x.device) if PRECISION[precision] == 3 else None
    else:
        out = torch.empty((b, cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    partial = torch.empty(L.rpst_conv1x1_stats_bytes(b, cout, hw) // 4, dtype=torch.float32, device=x.device) if want_stats else None
    _lib.check(L.rpst_conv1x1(x.data_ptr(), w_hi.data_ptr(), _ptr(w_lo), _ptr(bias_), _ptr(sub), _ptr(mul), _ptr(res), _ptr(out),
                              _ptr(out_hi), _ptr(out_lo), _ptr(partial), b, cin, cout, hw, 1 if act == "lrelu" else 0,
                              float(slope), PRECISION[precision], _stream()))
    result = (out_hi, out_lo) if packed else out
    if not want_stats:
        return result
    mean = torch.empty(b, cout, dtype=torch.float32, device=x.device)
    std = torch.empty_like(mean)
    _lib.check(L.rpst_conv1x1_stats_finalize(partial.data_ptr(), b, cout, hw, eps, mean.data_ptr(), std.data_ptr(), _stream()))
    return result, (mean.view(b, cout, 1, 1), std.view(b, cout, 1, 1))


@device_guard
def adain_from_stats(content: torch.Tensor, content_stats: Tuple[torch.Tensor, torch.Tensor],
                     style_stats: Tuple[torch.Tensor, torch.Tensor], prev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`[prev +] AdaIN(content, style)` (network/base.py:410-418, network/adain_rp.py:300-301) when both tensors'
    statistics are already known (conv-epilogue statistics): the style tensor is never read and the content is read
    once — 2E (3E with prev) bytes instead of 3E (4E)."""
    mu_c, sd_c = content_stats
    mu_s, sd_s = style_stats
    scale = (sd_s / sd_c).reshape(-1)
    shift = (mu_s.reshape(-1) - mu_c.reshape(-1) * scale)
    if prev is None:
        return plane_affine(content, scale, shift)
    x, p = _prep(content.detach(), "content"), _prep(prev.detach(), "prev")
    n, c = x.shape[:2]
    out = torch.empty_like(x)
    one = torch.ones_like(scale)
    _lib.check(_lib.lib().rpst_plane_affine2(x.data_ptr(), p.data_ptr(), scale.contiguous().data_ptr(), one.data_ptr(),
                                             shift.contiguous().data_ptr(), out.data_ptr(), n * c, x[0, 0].numel(), _stream()))
    return out


def sanet_forward_fused(module, content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """Inference forward of `SANet` (network/sanet.py:82-99) for in_planes = 512 with every 1x1 convolution on the
    tcgen05 block: statistics (one pass per tensor) -> Q / K projections with the instance normalisation folded in,
    emitted as packed operands -> H projection -> flash attention -> out_conv with the residual in the epilogue."""
    from .functional import calc_mean_std
    c, s = _prep(content.detach(), "content"), _prep(style.detach(), "style")
    b, ch, hc, wc = c.shape
    lc, ls = hc * wc, s.shape[2] * s.shape[3]
    prec = module.precision
    mu_c, sd_c = calc_mean_std(c)
    mu_s, sd_s = calc_mean_std(s)
    q_hi, q_lo = conv1x1(c, module.f.weight, module.f.bias, sub=mu_c, mul=1.0 / sd_c, packed=True, precision=prec)
    k_hi, k_lo = conv1x1(s, module.g.weight, module.g.bias, sub=mu_s, mul=1.0 / sd_s, packed=True, precision=prec)
    h = conv1x1(s, module.h.weight, module.h.bias, precision=prec)
    L = _lib.lib()
    o = torch.empty(b, ch, hc, wc, dtype=torch.float32, device=c.device)
    ws = _ws(L.rpst_sanet_attn_packed_workspace_bytes(b, lc, ls), c.device)
    _lib.check(L.rpst_sanet_attn_fwd_packed(q_hi.data_ptr(), _ptr(q_lo), k_hi.data_ptr(), _ptr(k_lo), h.data_ptr(), o.data_ptr(),
                                            b, ch, lc, ls, PRECISION[prec], ws.data_ptr(), ws.numel(), _stream()))
    return conv1x1(o, module.out_conv.weight, module.out_conv.bias, residual=c, precision=prec)


def sanet_fused_supported(content: torch.Tensor, style: torch.Tensor) -> bool:
    if content.dim() != 4 or style.dim() != 4 or content.shape[1] != 512 or style.shape[1] != 512:
        return False
    lc, ls = content.shape[2] * content.shape[3], style.shape[2] * style.shape[3]
    return lc % 128 == 0 and ls % 256 == 0
