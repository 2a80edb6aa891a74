"""`compute_img_bkg_seg` with the reference's signature (data/utils/found_bkg_mask.py:4-85); the arithmetic runs
in `csrc/pseudo_label.cu` (one CTA per image, only the needed row of the cosine matrix)."""
from __future__ import annotations

from typing import Tuple

import torch

from ... import ops
from ..._lib import UcodError


def compute_img_bkg_seg(attentions, feats, featmap_dims, th_bkg, up_size: int = None, dim=64,
                        epsilon: float = 1e-10, apply_weights: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """attentions [B,nh,T,T] (last layer), feats [B,T,nh*dim] (last-layer keys incl. CLS).
    Returns (bkg_mask [B,h,w] float {0,1}, sim_map [B,h,w] float), exactly as the reference."""
    w_f, h_f = featmap_dims
    if dim != 64:
        raise UcodError("compute_img_bkg_seg: head dim 64 (ViT-B) only")
    att = attentions[:, :, 0, 1:] if attentions.dim() == 4 else attentions
    keys = feats[:, 1:] if feats.shape[1] == w_f * h_f + 1 else feats
    nb, nh = att.shape[:2]
    if up_size is not None and up_size != w_f:
        # found_bkg_mask.py:26-27,50-55: attention and descriptors are bilinearly resampled to up_size^2 first.  The
        # head weights beta are constant per channel, so weighting commutes with the interpolation and stays in the kernel.
        att = ops.upsample_bilinear(att.reshape(nb * nh, w_f, h_f).float().contiguous(), (up_size, up_size))
        att = att.reshape(nb, nh, up_size * up_size)
        keys, _ = ops.resize_tokens_bilinear(keys.float().contiguous(), (w_f, h_f), (up_size, up_size), want_bf16=False)
        w_f = h_f = up_size
    cos, bkg, _, sim = ops.pseudo_label_score(att.reshape(nb, nh, -1), keys, float(th_bkg), float(epsilon), want_sim=True,
                                              apply_weights=bool(apply_weights))
    return bkg.reshape(nb, w_f, h_f).float(), sim.reshape(nb, w_f, h_f)
