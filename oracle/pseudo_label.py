"""ORACLE (test infrastructure — never imported by the product path).

Fixed-strategy pseudo-label generation, restated on CPU:

* `compute_img_bkg_seg`  — data/utils/found_bkg_mask.py:4-85 (FOUND-style background discovery; with
  `up_size=None` both F.interpolate calls are identities and are omitted here)
* `refine_post_process`  — generate_pseudo_label.py:30-67 (cv2.connectedComponentsWithStats replaced by
  oracle/cc.py, which is pinned against cv2)
* `generate_mask`        — generate_pseudo_label.py:70-94 glue (mask = 1 - bkg, then cleanup)

Parity pin: tools/make_golden.py calls the reference's own functions (imported from /root/reference) on planted
inputs and stores results in tests/golden/pseudo_label_*.npz.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import cc


@torch.no_grad()
def compute_img_bkg_seg(attentions, feats, featmap_dims, th_bkg, dim=64, epsilon: float = 1e-10,
                        apply_weights: bool = True, id_ref_override=None, want_att_sum: bool = False, up_size=None):
    """attentions [B,nh,T,T] (or the CLS row [B,nh,P] directly), feats [B,T,C] incl. CLS (or [B,P,C]).
    Returns (bkg_mask [B,h,w] float {0,1}, sim_map [B,h,w] float).  `up_size` (found_bkg_mask.py:19-20, :26-27,
    :50-55): attention and descriptors are bilinearly resampled to up_size^2 before anything else."""
    w_f, h_f = featmap_dims
    P = w_f * h_f
    if attentions.dim() == 4:
        att = attentions[:, :, 0, 1:]
    else:
        att = attentions
    nb, nh = att.shape[:2]
    att = att.reshape(nb, nh, P).float()
    descs = feats[:, 1:] if feats.shape[1] == P + 1 else feats
    descs = descs.float()
    if up_size is not None and up_size != w_f:
        att = F.interpolate(att.reshape(nb, nh, w_f, h_f), size=(up_size, up_size), mode="bilinear")
        # the reference weights the descriptors by beta (constant per channel) and then interpolates; interpolation is
        # linear per channel, so resampling first and weighting afterwards is the same arithmetic up to rounding
        descs = F.interpolate(descs.reshape(nb, w_f, h_f, -1).permute(0, 3, 1, 2), size=(up_size, up_size), mode="bilinear")
        descs = descs.permute(0, 2, 3, 1).reshape(nb, up_size * up_size, -1)
        w_f = h_f = up_size
        P = up_size * up_size
        att = att.reshape(nb, nh, P)
    threshold = torch.mean(att.reshape(nb, -1), dim=1)
    Q = torch.sum(att > threshold[:, None, None], dim=2) / P
    beta = torch.log(torch.sum(Q + epsilon, dim=1)[:, None] / (Q + epsilon))
    if apply_weights:
        descs = (descs.reshape(nb, P, nh, dim) * beta[:, None, :, None]).reshape(nb, P, nh * dim)
    descs = F.normalize(descs, dim=-1, p=2)
    cos_sim = torch.bmm(descs, descs.permute(0, 2, 1))
    if apply_weights:
        att = att * beta[:, :, None]
    att_sum = torch.sum(att, dim=1)
    id_ref = torch.argmin(att_sum, dim=-1)
    if id_ref_override is not None:  # test hook: evaluate the map for a given reference patch (near-tie analysis)
        id_ref = torch.as_tensor(id_ref_override, dtype=torch.long).reshape(nb)
    row = cos_sim[torch.arange(nb), id_ref, :].reshape(nb, w_f, h_f)
    bkg = row > th_bkg
    sim = 1 - row.float()
    sim = sim / (sim.max() + 1e-10)
    if want_att_sum:
        return bkg.float(), (sim * (1 - bkg.float())).float(), row, id_ref, att_sum
    return bkg.float(), (sim * (1 - bkg.float())).float(), row, id_ref


def refine_post_process(mask, area_threshold: int = 4) -> np.ndarray:
    """mask [h,w] {0,1} -> refined [h,w] uint8 (generate_pseudo_label.py:30-67, sequential over labels)."""
    mask_np = np.asarray(mask).astype(np.uint8).squeeze()
    num_labels, labels = cc.connected_components_8(mask_np)
    st = cc.stats(labels, num_labels)
    refined = mask_np.copy()
    H, W = mask_np.shape
    for label in range(1, num_labels):
        x, y, width, height, area = (int(v) for v in st[label])
        if area >= area_threshold:
            continue
        comp = labels[y:y + height, x:x + width] == label
        x0, y0 = max(x - 1, 0), max(y - 1, 0)
        x1, y1 = min(x + width + 1, W), min(y + height + 1, H)
        surrounding = refined[y0:y1, x0:x1].copy()
        ring = np.ones_like(surrounding, dtype=bool)
        cy, cx = np.where(comp)
        ring[cy + (y - y0), cx + (x - x0)] = False
        component_label = int(refined[y + height // 2, x + width // 2])
        opposite = 1 - component_label
        if np.all(surrounding[ring] == opposite):
            view = refined[y:y + height, x:x + width]
            view[comp] = opposite % 256
    return refined


def generate_mask_from_outputs(attn_cls, key_patches, grid, th_bkg: float = 0.6):
    """generate_pseudo_label.py:79-92 for a batch evaluated image by image (the reference runs B=1)."""
    out = []
    for b in range(attn_cls.shape[0]):
        bkg, _, _, _ = compute_img_bkg_seg(attn_cls[b:b + 1], key_patches[b:b + 1], grid, th_bkg,
                                           dim=key_patches.shape[-1] // attn_cls.shape[1])
        out.append(refine_post_process((1 - bkg[0]).numpy()))
    return np.stack(out)
