"""WCT whitening/colouring — host-side mirror of network/wct_rp.py:7-40 (`matrix_sqrt`,
`matrix_inv_sqrt`), :82-114 (`WCTRPNet.whiten_and_color`) and :157-166 (`WCTRPNet.fuse`)."""
from __future__ import annotations

import torch

from . import _lib
from .functional import _prep, _ptr, _stream, device_guard

_METHODS = {"closed-form": 0, "original": 1}


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _sym_fn(A: torch.Tensor, want_sqrt: bool, want_inv: bool, diag_add: float = 1e-4):
    if not A.is_cuda:
        raise RuntimeError("rpst: matrix functions need a CUDA tensor (there is no CPU path)")
    squeeze = A.dim() == 2
    a = (A[None] if squeeze else A).to(torch.float64).contiguous()
    assert a.dim() == 3 and a.shape[1] == a.shape[2]
    b, n = a.shape[:2]
    L = _lib.lib()
    ws = _ws(L.rpst_sym_eig_fn_workspace_bytes(b, n), a.device)
    rs = torch.empty_like(a) if want_sqrt else None
    ri = torch.empty_like(a) if want_inv else None
    _lib.check(L.rpst_sym_eig_fn(a.data_ptr(), b, n, diag_add, _ptr(rs), _ptr(ri), None, None,
                                 ws.data_ptr(), ws.numel(), _stream()))
    fix = (lambda t: None if t is None else (t[0] if squeeze else t).to(A.dtype))
    return fix(rs), fix(ri)


@device_guard
def matrix_sqrt(A: torch.Tensor) -> torch.Tensor:
    """Drop-in for network/wct_rp.py:25 (also accepts a batch [B,n,n])."""
    return _sym_fn(A, True, False)[0]


@device_guard
def matrix_inv_sqrt(A: torch.Tensor) -> torch.Tensor:
    """Drop-in for network/wct_rp.py:7."""
    return _sym_fn(A, False, True)[1]


@device_guard
def wct_fuse(content_feats: torch.Tensor, style_feats: torch.Tensor, method: str = "closed-form",
             precision: str = "fp32", return_transform: bool = False):
    """Batched `WCTRPNet.fuse`: [N,C,H,W] x [N,C,Hs,Ws] -> [N,C,H,W] fp32, inputs detached like the
    reference (network/wct_rp.py:161-162)."""
    assert method in _METHODS, method
    c4 = _prep(content_feats.detach(), "content_feats")
    s4 = _prep(style_feats.detach(), "style_feats")
    assert c4.dim() == 4 and s4.dim() == 4 and c4.shape[:2] == s4.shape[:2]
    n, c = c4.shape[:2]
    hw_c, hw_s = c4[0, 0].numel(), s4[0, 0].numel()
    out = torch.empty_like(c4)
    tr = torch.empty(n, c, c, dtype=torch.float64, device=c4.device) if return_transform else None
    L = _lib.lib()
    ws = _ws(L.rpst_wct_workspace_bytes(n, c, hw_c, hw_s), c4.device)
    _lib.check(L.rpst_wct_fuse(c4.data_ptr(), s4.data_ptr(), out.data_ptr(), n, c, hw_c, hw_s, _METHODS[method],
                               {"fp32": 3, "bf16": 1}[precision], _ptr(tr), ws.data_ptr(), ws.numel(), _stream()))
    return (out, tr) if return_transform else out


@device_guard
def spd_roots(A: torch.Tensor, lmin: float, diag_add: float = 1e-4):
    """(A + diag_add I)^(1/2), (A + diag_add I)^(-1/2) and the per-matrix "not accepted" flags of the Newton-Schulz
    iteration (`rpst_spd_roots`); A [b,n,n] or [n,n] symmetric with every shifted eigenvalue >= lmin."""
    if not A.is_cuda:
        raise RuntimeError("rpst: matrix functions need a CUDA tensor (there is no CPU path)")
    squeeze = A.dim() == 2
    a = (A[None] if squeeze else A).to(torch.float64).contiguous()
    assert a.dim() == 3 and a.shape[1] == a.shape[2]
    b, n = a.shape[:2]
    L = _lib.lib()
    ws = _ws(L.rpst_spd_roots_workspace_bytes(b, n), a.device)
    rs, ri = torch.empty_like(a), torch.empty_like(a)
    flags = torch.zeros(b, dtype=torch.int32, device=a.device)
    _lib.check(L.rpst_spd_roots(a.data_ptr(), b, n, diag_add, lmin, rs.data_ptr(), ri.data_ptr(), flags.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream()))
    return (rs[0], ri[0], flags) if squeeze else (rs, ri, flags)


@device_guard
def whiten_and_color(cF: torch.Tensor, sF: torch.Tensor, method: str = "closed-form") -> torch.Tensor:
    """Function form of network/wct_rp.py:82 — cF [C,HWc], sF [C,HWs] (any float dtype; the reference
    passes fp64) -> [C,HWc] in cF's dtype."""
    assert cF.dim() == 2 and sF.dim() == 2 and cF.shape[0] == sF.shape[0]
    c = cF.shape[0]
    out = wct_fuse(cF.float().reshape(1, c, 1, -1), sF.float().reshape(1, c, 1, -1), method)
    return out.reshape(c, -1).to(cF.dtype)


def fuse(self, content_feats, style_feats):
    """Replacement for the bound method `WCTRPNet.fuse` (network/wct_rp.py:157)."""
    return wct_fuse(content_feats, style_feats)
