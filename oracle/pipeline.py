"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's first-stage eval step for one batch (BASELINE.json configs 1/2), following
engine/runner/loop_UCOD_DPL.py:297-311 (`ValLoop_Look_Twice.run`: upsample cached features to feature_size,
student decoder) and :354-361 (`process_preds`: bilinear to image size, sigmoid > 0.5), with the feature cache
filled as data/datasets/base_dataset.py:113-121 does (backbone over the normalised image).
"""
from __future__ import annotations

import torch

from . import decoder as odec
from . import vit as ovit


@torch.no_grad()
def first_stage_eval(vit_sd, spec, dec_sd, images_u8: torch.Tensor, image_size, feature_size: int = 68):
    """images_u8 [B,3,S,S] uint8 -> dict(keys [B,768,g,g], logits [B,1,fs,fs], mask [B,S,S] uint8)."""
    x = ovit.normalize_u8(images_u8)
    out = ovit.vit_forward(vit_sd, spec, x)
    keys = ovit.keys_to_map(out["key_tokens"])
    feats = odec.upsample_bilinear(keys, (feature_size, feature_size))
    fg, _, _ = odec.baseline_forward(dec_sd, feats, want_ortho=False)
    up = odec.upsample_bilinear(fg, image_size)
    mask = (torch.sigmoid(up) > 0.5).squeeze(1).to(torch.uint8)
    return {"keys": keys, "logits": fg, "mask": mask}
