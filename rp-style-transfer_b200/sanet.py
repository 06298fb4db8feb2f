"""SANet / AdaptiveSANet / Transform — host-side mirrors of network/sanet.py:12-160 with identical
constructor signatures, forward arguments and state-dict names (f, g, h, out_conv,
attention_layer.f_psi.{0,2}, sanet4_1, sanet5_1, merge_conv), so reference checkpoints load unchanged.
The 1x1 / 3x3 convolutions run in cuDNN through torch (they are out of the transform's scope,
SURVEY.md §2 #9-10); normalisation and the attention core call librpst."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .functional import _prep, _ptr, _stream, device_guard, mean_variance_norm

PRECISION = {"fp32": 3, "bf16": 1}


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_GROUP_BYTES = 4 << 30   # workspace budget that buys multi-sample launches (see include/rpst.h)


def _group_ws(per_sample: int, b: int, device) -> torch.Tensor:
    """The attention entry points take `k * per_sample` bytes and then run k samples per launch."""
    k = max(1, min(b, _GROUP_BYTES // max(int(per_sample), 1)))
    return _ws(k * int(per_sample), device)


@device_guard
def cal_affinity_matrix(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    """Drop-in for network/sanet.py:12 — [b,c,h,w] x2 -> [b,hw,hw] cosine affinity (differentiable)."""
    assert content_feat.size() == style_feat.size()
    if torch.is_grad_enabled() and (content_feat.requires_grad or style_feat.requires_grad):
        return _AffinityFn.apply(_prep(content_feat, "content_feat"), _prep(style_feat, "style_feat"))
    return _cal_affinity_raw(content_feat, style_feat)


def _cal_affinity_raw(content_feat: torch.Tensor, style_feat: torch.Tensor) -> torch.Tensor:
    c4, s4 = _prep(content_feat.detach(), "content_feat"), _prep(style_feat.detach(), "style_feat")
    b, c, h, w = c4.shape
    l = h * w
    out = torch.empty(b, l, l, dtype=torch.float32, device=c4.device)
    L = _lib.lib()
    ws = _ws(L.rpst_cosine_affinity_workspace_bytes(c, l, l), c4.device)
    _lib.check(L.rpst_cosine_affinity(c4.data_ptr(), s4.data_ptr(), out.data_ptr(), b, c, l, l, ws.data_ptr(), ws.numel(), _stream()))
    return out


def _attn_fwd_raw(F, G, H, precision, return_attn):
    b, c, hc, wc = F.shape
    ls = G.shape[2] * G.shape[3]
    lc = hc * wc
    out = torch.empty(b, c, hc, wc, dtype=torch.float32, device=F.device)
    attn = torch.empty(b, lc, ls, dtype=torch.float32, device=F.device) if return_attn else None
    L = _lib.lib()
    ws = _group_ws(L.rpst_sanet_attn_workspace_bytes(c, lc, ls), b, F.device)
    _lib.check(L.rpst_sanet_attn_fwd(F.data_ptr(), G.data_ptr(), H.data_ptr(), out.data_ptr(), b, c, lc, ls,
                                     PRECISION[precision], _ptr(attn), ws.data_ptr(), ws.numel(), _stream()))
    return out, attn


class _AttnFn(torch.autograd.Function):
    """Differentiable attention core: the backward pass recomputes the attention matrix
    (rpst_sanet_attn_bwd) instead of keeping L x L state alive."""

    @staticmethod
    def forward(ctx, F, G, H, precision):
        out, _ = _attn_fwd_raw(F, G, H, precision, False)
        ctx.save_for_backward(F, G, H)
        ctx.precision = precision
        return out

    @staticmethod
    def backward(ctx, grad_out):
        F, G, H = ctx.saved_tensors
        b, c, hc, wc = F.shape
        lc, ls = hc * wc, G.shape[2] * G.shape[3]
        go = _prep(grad_out, "grad_out")
        dF, dG, dH = torch.empty_like(F), torch.empty_like(G), torch.empty_like(H)
        L = _lib.lib()
        ws = _group_ws(L.rpst_sanet_attn_bwd_workspace_bytes(c, lc, ls), b, F.device)
        _lib.check(L.rpst_sanet_attn_bwd(F.data_ptr(), G.data_ptr(), H.data_ptr(), go.data_ptr(), dF.data_ptr(),
                                         dG.data_ptr(), dH.data_ptr(), b, c, lc, ls, PRECISION[ctx.precision],
                                         ws.data_ptr(), ws.numel(), _stream()))
        return dF, dG, dH, None


@device_guard
def attention_core(F: torch.Tensor, G: torch.Tensor, H: torch.Tensor, precision: str = "fp32", return_attn: bool = False):
    """softmax(F^T G) applied to H: F [b,c,hc,wc], G/H [b,c,hs,ws] -> [b,c,hc,wc] (network/sanet.py:85-94).
    Differentiable w.r.t. F, G and H (the attention matrix itself is returned detached)."""
    F, G, H = _prep(F, "F"), _prep(G, "G"), _prep(H, "H")
    assert G.shape[:2] == F.shape[:2] and H.shape == G.shape
    if torch.is_grad_enabled() and (F.requires_grad or G.requires_grad or H.requires_grad):
        out = _AttnFn.apply(F, G, H, precision)
        if not return_attn:
            return out
        with torch.no_grad():
            return out, _attn_fwd_raw(F, G, H, precision, True)[1]
    out, attn = _attn_fwd_raw(F, G, H, precision, return_attn)
    return (out, attn) if return_attn else out


class _AffinityFn(torch.autograd.Function):
    """Differentiable cosine affinity (network/sanet.py:12-18): forward is the tcgen05 kernel; the backward
    pass is two [C,L]x[L,L] products per sample on the same tcgen05 GEMM block (bf16x3) and the normalisation's
    closed-form Jacobian."""

    @staticmethod
    def forward(ctx, content, style):
        ctx.save_for_backward(content, style)
        return _cal_affinity_raw(content, style)

    @staticmethod
    def backward(ctx, daff):
        from .mrf import packed_gemm
        content, style = ctx.saved_tensors
        b, c = content.shape[:2]
        cv, sv = content.reshape(b, c, -1), style.reshape(b, c, -1)
        nc, ns = cv.norm(dim=1, keepdim=True).clamp_min(1e-12), sv.norm(dim=1, keepdim=True).clamp_min(1e-12)
        ch, sh = cv / nc, sv / ns
        daff = daff.contiguous()
        # d c^[c,i] = sum_j s^[c,j] daff[i,j] ;  d s^[c,j] = sum_i c^[c,i] daff[i,j]
        dch = torch.stack([packed_gemm(sh[i], daff[i]) for i in range(b)])
        dsh = torch.stack([packed_gemm(ch[i], daff[i].t()) for i in range(b)])
        dc = (dch - ch * (ch * dch).sum(1, keepdim=True)) / nc
        ds = (dsh - sh * (sh * dsh).sum(1, keepdim=True)) / ns
        return dc.view_as(content), ds.view_as(style)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM block (bias in the epilogue); dx = g W, dW = g^T x, db = colsum(g).
    Used for f_psi's Linear(L -> L/16) — 550 GFLOP per sample at L = 16384 (network/sanet.py:34-39)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from .mrf import packed_gemm
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return packed_gemm(x, weight, col_add=bias)

    @staticmethod
    def backward(ctx, g):
        from .mrf import packed_gemm
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        dx = packed_gemm(g, weight.t()) if ctx.needs_input_grad[0] else None
        dw = packed_gemm(g.t(), x.t()) if ctx.needs_input_grad[1] else None
        db = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def _f_psi(al, x2d: torch.Tensor) -> torch.Tensor:
    """The clamp MLP of AEAModule / AEALReluModule on [rows, L] affinity rows with the module's own parameters:
    Linear(L -> L/16) on tensor cores, the rest (LeakyReLU, Linear(L/16 -> 1), sigmoid/tanh) is O(rows * L/16)."""
    lin0 = al.f_psi[0]
    hid = _LinearFn.apply(x2d.contiguous(), lin0.weight, lin0.bias)
    return al.f_psi[3](al.f_psi[2](torch.nn.functional.leaky_relu(hid, 0.2)))


class _ClampedAttnFn(torch.autograd.Function):
    """out = H S'^T with S' = clamp-activation(softmax(F^T G), clamp) (network/sanet.py:41-46, 66-71, 114-138),
    forward and backward in librpst (rpst_sanet_attn_clamped_fwd / _bwd); differentiable w.r.t. F, G, H and the
    per-row clamp, so the clamp MLP trains through ordinary autograd."""

    @staticmethod
    def forward(ctx, F, G, H, clamp, mode, scale, precision):
        b, c, hc, wc = F.shape
        lc, ls = hc * wc, G.shape[2] * G.shape[3]
        out = torch.empty(b, c, hc, wc, dtype=torch.float32, device=F.device)
        L = _lib.lib()
        ws = _ws(L.rpst_sanet_attn_clamped_workspace_bytes(c, lc, ls), F.device)
        _lib.check(L.rpst_sanet_attn_clamped_fwd(F.data_ptr(), G.data_ptr(), H.data_ptr(), clamp.data_ptr(), mode, scale,
                                                 out.data_ptr(), b, c, lc, ls, PRECISION[precision], ws.data_ptr(),
                                                 ws.numel(), _stream()))
        ctx.save_for_backward(F, G, H, clamp)
        ctx.mode, ctx.scale, ctx.precision = mode, scale, precision
        return out

    @staticmethod
    def backward(ctx, grad_out):
        F, G, H, clamp = ctx.saved_tensors
        b, c, hc, wc = F.shape
        lc, ls = hc * wc, G.shape[2] * G.shape[3]
        go = _prep(grad_out, "grad_out")
        dF, dG, dH, dcl = torch.empty_like(F), torch.empty_like(G), torch.empty_like(H), torch.empty_like(clamp)
        L = _lib.lib()
        ws = _ws(L.rpst_sanet_attn_clamped_workspace_bytes(c, lc, ls), F.device)
        _lib.check(L.rpst_sanet_attn_clamped_bwd(F.data_ptr(), G.data_ptr(), H.data_ptr(), clamp.data_ptr(), ctx.mode,
                                                 ctx.scale, go.data_ptr(), dF.data_ptr(), dG.data_ptr(), dH.data_ptr(),
                                                 dcl.data_ptr(), b, c, lc, ls, PRECISION[ctx.precision], ws.data_ptr(),
                                                 ws.numel(), _stream()))
        return dF, dG, dH, dcl, None, None, None


class AEAModule(nn.Module):
    """network/sanet.py:26-46."""

    def __init__(self, inplanes, scale_value=50, from_value=0.4, value_interval=0.5):
        super().__init__()
        self.inplanes = inplanes
        self.scale_value = scale_value
        self.from_value = from_value
        self.value_interval = value_interval
        self.f_psi = nn.Sequential(
            nn.Linear(self.inplanes, self.inplanes // 16),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Linear(self.inplanes // 16, 1),
            nn.Sigmoid(),
        )
    mode = 1

    def forward(self, x, f_x):
        # stand-alone use (the fused path is AdaptiveSANet.forward): same algebra on the module's own layers
        b, hw, c = x.size()
        clamp_value = _f_psi(self, x.reshape(b * hw, c)) * self.value_interval + self.from_value
        clamp_value = clamp_value.view(b, hw, 1)
        return torch.sigmoid(self.scale_value * (f_x - clamp_value)), clamp_value


class AEALReluModule(nn.Module):
    """network/sanet.py:49-71."""

    def __init__(self, inplanes, scale_value=50, from_value=0.4, value_interval=0.5):
        super().__init__()
        self.inplanes = inplanes
        self.scale_value = scale_value
        self.from_value = from_value
        self.value_interval = value_interval
        self.f_psi = nn.Sequential(
            nn.Linear(self.inplanes, self.inplanes // 16),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Linear(self.inplanes // 16, 1),
            nn.Tanh(),
        )
        self.clamp_sig = nn.Sequential(nn.ReLU(inplace=True), nn.Softmax(dim=-1))
    mode = 2

    def forward(self, x, f_x):
        b, hw, c = x.size()
        clamp_value = ((_f_psi(self, x.reshape(b * hw, c)) + 1) / 2).view(b, hw, 1)
        return self.clamp_sig(f_x - clamp_value), clamp_value


class SANet(nn.Module):
    """network/sanet.py:73-99."""

    def __init__(self, in_planes):
        super().__init__()
        self.f = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.g = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.h = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.sm = nn.Softmax(dim=-1)
        self.out_conv = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.precision = "fp32"
        self.fused = True      # inference at in_planes = 512: projections + normalisation + residual on the tcgen05 block

    def forward(self, content, style):
        from .conv import sanet_forward_fused, sanet_fused_supported
        needs_grad = torch.is_grad_enabled() and (content.requires_grad or style.requires_grad or
                                                  any(p.requires_grad for p in self.parameters()))
        if getattr(self, "fused", True) and not needs_grad and sanet_fused_supported(content, style):
            return sanet_forward_fused(self, content, style)
        F = self.f(mean_variance_norm(content))
        G = self.g(mean_variance_norm(style))
        H = self.h(style)
        O = attention_core(F, G, H, self.precision)
        O = self.out_conv(O)
        O += content
        return O


class AdaptiveSANet(nn.Module):
    """network/sanet.py:100-138 (keeps claim_value / claim_before / claim_after for the visualisation
    code at network/sanet.py:346-348)."""

    def __init__(self, in_planes, spatial_dims, ada_module='aea'):
        super().__init__()
        self.f = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.g = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.h = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.sm = nn.Softmax(dim=-1)
        self.out_conv = nn.Conv2d(in_planes, in_planes, (1, 1))
        self.attention_layer = AEAModule(spatial_dims) if ada_module == 'aea' else AEALReluModule(spatial_dims)
        self.claim_value = 0
        self.claim_before = 0
        self.claim_after = 0
        self.claim_after_sm = 0
        self.precision = "fp32"
        self.keep_claims = True

    def forward(self, content, style):
        assert content.size() == style.size()
        F = self.f(mean_variance_norm(content))
        G = self.g(mean_variance_norm(style))
        H = self.h(style)
        if torch.is_grad_enabled() and any(t.requires_grad for t in (F, G, H, content, style)):
            return self._forward_training(content, style, F, G, H)
        F, G, H = (_prep(t.detach(), n) for t, n in ((F, "F"), (G, "G"), (H, "H")))
        c_raw, s_raw = _prep(content.detach(), "content"), _prep(style.detach(), "style")
        b, c, hh, ww = F.shape
        l = hh * ww
        al = self.attention_layer
        lin0, lin2 = al.f_psi[0], al.f_psi[2]
        assert lin0.in_features == l, f"spatial_dims={lin0.in_features} does not match H*W={l} (network/sanet.py:290-292)"
        lh = lin0.out_features
        dev = F.device
        out = torch.empty(b, c, hh, ww, dtype=torch.float32, device=dev)
        before = torch.empty(b, l, l, dtype=torch.float32, device=dev) if self.keep_claims else None
        after = torch.empty(b, l, l, dtype=torch.float32, device=dev) if self.keep_claims else None
        clamp = torch.empty(b, l, 1, dtype=torch.float32, device=dev)
        L = _lib.lib()
        ws = _ws(L.rpst_sanet_adaptive_workspace_bytes(c, l, l, lh), dev)
        w0 = lin0.weight.detach().contiguous()
        b0 = lin0.bias.detach().contiguous()
        w2 = lin2.weight.detach().reshape(-1).contiguous()
        b2 = lin2.bias.detach().contiguous()
        _lib.check(L.rpst_sanet_attn_adaptive_fwd(
            F.data_ptr(), G.data_ptr(), H.data_ptr(), c_raw.data_ptr(), s_raw.data_ptr(), c_raw.shape[1],
            w0.data_ptr(), b0.data_ptr(), w2.data_ptr(), b2.data_ptr(), al.mode, float(al.scale_value),
            float(al.from_value), float(al.value_interval), out.data_ptr(), _ptr(before), _ptr(after), clamp.data_ptr(),
            b, c, l, l, lh, PRECISION[self.precision], ws.data_ptr(), ws.numel(), _stream()))
        self.claim_before = before
        self.claim_after = after
        O = self.out_conv(out)
        O += content
        self.claim_value = clamp
        return O


def _adaptive_forward_training(self, content, style, F, G, H):
    """Training path (network/sanet.py:114-138 under autograd): the cosine affinity and the clamped attention
    are librpst kernels with their own backward passes; the small clamp MLP `f_psi` (Linear L -> L/16 -> 1 on
    the affinity rows) runs as the module's own layers so its parameters train through autograd."""
    b, c, hh, ww = F.shape
    l = hh * ww
    al = self.attention_layer
    assert al.f_psi[0].in_features == l, f"spatial_dims={al.f_psi[0].in_features} does not match H*W={l}"
    aff = _AffinityFn.apply(_prep(content, "content"), _prep(style, "style"))
    z = _f_psi(al, aff.view(b * l, l))
    clamp = z * al.value_interval + al.from_value if al.mode == 1 else (z + 1) / 2
    clamp = clamp.view(b, l)
    out = _ClampedAttnFn.apply(_prep(F, "F"), _prep(G, "G"), _prep(H, "H"), clamp.contiguous(), al.mode,
                               float(al.scale_value), self.precision)
    self.claim_before = self.claim_after = None      # not materialised in training
    self.claim_value = clamp.detach().view(b, l, 1)
    O = self.out_conv(out)
    O = O + content
    return O


AdaptiveSANet._forward_training = _adaptive_forward_training


class Transform(nn.Module):
    """network/sanet.py:140-149."""

    def __init__(self, in_planes):
        super().__init__()
        self.sanet4_1 = SANet(in_planes=in_planes)
        self.sanet5_1 = SANet(in_planes=in_planes)
        self.upsample5_1 = nn.Upsample(scale_factor=2, mode='nearest')
        self.merge_conv_pad = nn.ReflectionPad2d((1, 1, 1, 1))
        self.merge_conv = nn.Conv2d(in_planes, in_planes, (3, 3))

    def forward(self, content4_1, style4_1, content5_1, style5_1):
        return self.merge_conv(self.merge_conv_pad(
            self.sanet4_1(content4_1, style4_1) + self.upsample5_1(self.sanet5_1(content5_1, style5_1))))


class AdaptiveTransform(nn.Module):
    """network/sanet.py:151-160."""

    def __init__(self, in_planes, relu4_1_dims, relu5_1_dims, ada_module='aea'):
        super().__init__()
        self.sanet4_1 = AdaptiveSANet(in_planes=in_planes, spatial_dims=relu4_1_dims, ada_module=ada_module)
        self.sanet5_1 = AdaptiveSANet(in_planes=in_planes, spatial_dims=relu5_1_dims, ada_module=ada_module)
        self.upsample5_1 = nn.Upsample(scale_factor=2, mode='nearest')
        self.merge_conv_pad = nn.ReflectionPad2d((1, 1, 1, 1))
        self.merge_conv = nn.Conv2d(in_planes, in_planes, (3, 3))

    def forward(self, content4_1, style4_1, content5_1, style5_1):
        return self.merge_conv(self.merge_conv_pad(
            self.sanet4_1(content4_1, style4_1) + self.upsample5_1(self.sanet5_1(content5_1, style5_1))))
