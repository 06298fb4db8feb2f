"""Decode loops of the reference's multiscale networks with the blend fused into the transform.

The reference spells the multiscale blend as two ops, `stylized + AdaIN(c_l, s_l)`
(network/adain_rp.py:300-301); these replacements make it one kernel call (`adain_blend`).  They are
installed over `<Class>.decode` by `rpst.install()` and also usable directly (`multiscale_transform`)."""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import functional as F
from . import segment as SEG


def multiscale_transform(content_feats: Sequence[torch.Tensor], style_feats: Sequence[torch.Tensor],
                         prevs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """The transform calls of `MultiScaleAdaINRPNet.decode` with the decoder convolutions factored
    out (features ordered shallow -> deep, `prevs[l]` = decoder state blended at level l); outputs
    ordered deep -> shallow.  This is the hot path bench.py times."""
    outs = [F.adaptive_instance_normalization(content_feats[-1], style_feats[-1])]
    for l in range(len(content_feats) - 2, -1, -1):
        outs.append(F.adain_blend(prevs[l], content_feats[l], style_feats[l]))
    return outs


def decode_multiscale(self, content_feats, style_feats, use_mask=False, c_mask_path=None, s_mask_path=None):
    """MultiScaleAdaINRPNet.decode (network/adain_rp.py:286-302): top level plain (seg-)AdaIN, every
    shallower level `decoder(prev + (seg-)AdaIN(c_l, s_l))` with the add fused into the transform."""
    if self._sort:
        content_feats = self.sort_by_weights(content_feats)
        style_feats = self.sort_by_weights(style_feats)
    if use_mask:
        top = SEG.do_mask_stylized(content_feats[-1], style_feats[-1], c_mask_path, s_mask_path)
    else:
        top = F.adaptive_instance_normalization(content_feats[-1], style_feats[-1])
    stylized = self.rp_decoder[0](top)
    lower = list(zip(content_feats[:-1], style_feats[:-1]))[::-1]
    for i, (cf, sf) in enumerate(lower):
        if use_mask:
            fused = SEG.do_mask_stylized(cf, sf, c_mask_path, s_mask_path, prev=stylized)
        else:
            fused = F.adain_blend(stylized, cf, sf)
        stylized = self.rp_decoder[i + 1](fused)
    return stylized


def _top(self, content_feats, style_feats, use_mask, c_mask_path, s_mask_path):
    if use_mask:
        return SEG.do_mask_stylized(content_feats[-1], style_feats[-1], c_mask_path, s_mask_path)
    return F.adaptive_instance_normalization(content_feats[-1], style_feats[-1])


def decode_ldms(self, content_feats, style_feats, use_mask=False, c_mask_path=None, s_mask_path=None):
    """LDMSAdaINRPNet.decode (network/adain_rp.py:538-553; inherited by LDMS 2/3): decoder blocks are
    attributes `rp_dec{i}`; lower levels use the running decoder state as content,
    `stylized + AdaIN(stylized, s_l)` (prev aliases content), only while `i < stylized_layers-1`."""
    stylized = self.rp_dec0(_top(self, content_feats, style_feats, use_mask, c_mask_path, s_mask_path))
    lower = list(zip(content_feats[:-1], style_feats[:-1]))[::-1]
    for i, (cf, sf) in enumerate(lower):
        if i < self.stylized_layers - 1:
            if use_mask:
                fused = SEG.do_mask_stylized(cf, sf, c_mask_path, s_mask_path, prev=stylized)
            else:
                fused = F.adain_blend(stylized, stylized, sf)
        else:
            fused = stylized + []   # the reference adds an empty list here (:543,552) and fails the same way
        stylized = getattr(self, f'rp_dec{i + 1}')(fused)
    return stylized


def decode_ld_concat(self, content_feats, style_feats, use_mask=False, c_mask_path=None, s_mask_path=None):
    """LDMSAdaINRPNet4.decode (network/adain_rp.py:780-799, inherited by LDMS 5):
    `cat([stylized, AdaIN(c_l, s_l)], 1)` with the AdaIN half written straight into the result."""
    stylized = self.rp_dec0(_top(self, content_feats, style_feats, use_mask, c_mask_path, s_mask_path))
    lower = list(zip(content_feats[:-1], style_feats[:-1]))[::-1]
    for i, (cf, sf) in enumerate(lower):
        if use_mask:
            fused = torch.cat([stylized, SEG.do_mask_stylized(cf, sf, c_mask_path, s_mask_path)], dim=1)
        else:
            fused = F.adain_concat(stylized, cf, sf)
        stylized = getattr(self, f'rp_dec{i + 1}')(fused)
    return stylized


PATCHED_DECODES = {
    "MultiScaleAdaINRPNet": decode_multiscale,
    "LDMSAdaINRPNet": decode_ldms,
    "LDMSAdaINRPNet4": decode_ld_concat,
}
