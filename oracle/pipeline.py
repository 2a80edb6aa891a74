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


@torch.no_grad()
def coral_eval(vit_sd, spec, dec_sd, ref_sd, image_u8_hwc, image_size: int, label_size, window_size: int = 3,
               window_length: int = 56, threshold: float = 0.0015):
    """Second-stage eval of ONE image (require_m_patches=False): data/datasets/lr_dataset.py:82-152 feature
    production + engine/runner/loop_CORAL.py:130-166.  image_u8_hwc: numpy [H0,W0,3] uint8."""
    import numpy as np

    from . import coral as oc
    from . import pil_resample as pr

    S, ws, g = image_size, window_size, window_length

    def features(img):
        low = pr.resize_u8(img, S, S, "bilinear")                      # transforms.Resize((S,S)) on the PIL image
        big = pr.resize_u8(img, S * ws, S * ws, "bilinear")            # self.resize_img, lr_dataset.py:57-63
        wins = [big[i * S:(i + 1) * S, j * S:(j + 1) * S] for i in range(ws) for j in range(ws)]
        batch = torch.from_numpy(np.stack([low] + wins)).permute(0, 3, 1, 2)
        keys = ovit.keys_to_map(ovit.vit_forward(vit_sd, spec, ovit.normalize_u8(batch))["key_tokens"])
        return keys[:1], keys[1:].unsqueeze(0)

    def prepare(l, h):
        lf = odec.upsample_bilinear(l, (g, g))
        hf = odec.upsample_bilinear(h.flatten(0, 1), (g, g)).reshape(1, ws * ws, -1, g, g)
        preds, _, _ = odec.baseline_forward(dec_sd, lf, want_ortho=False)
        return lf, hf, preds

    lf, hf, preds = prepare(*features(image_u8_hwc))
    coarse = preds
    crop = oc.should_crop_center(preds)
    if crop:
        H0, W0 = image_u8_hwc.shape[:2]
        nh, nw = H0 // 2, W0 // 2
        top, left = (H0 - nh) // 2, (W0 - nw) // 2
        lf, hf, preds = prepare(*features(image_u8_hwc[top:top + nh, left:left + nw]))
    out, opt = oc.sparse_refiner_forward(ref_sd, lf, hf, preds, threshold, ws)
    if crop:
        out = oc.center_pad(out)
    return {"coarse": coarse, "crop": crop, "refined": out, "mask": oc.process_preds(out, label_size), "opt": opt}


@torch.no_grad()
def look_twice_eval(vit_sd, spec, dec_sd, images_u8: torch.Tensor, image_size, feature_size: int = 68,
                    look_twice_th: float = 0.15, expand_type: str = "dynamic", first_logits=None, label_size=None):
    """First-stage eval WITH Look-Twice for a batch whose originals are the network-size images themselves
    (engine/runner/loop_UCOD_DPL.py:297-317, per image like the reference): first look -> process_preds -> second look
    over the boxes -> final bilinear resize to `label_size` -> `> 0.5`.
    first_logits (optional [B,1,fs,fs]): replaces the first-look prediction as process_preds' input — the SURVEY 8(d)
    "LT" benchmark plants two small objects per image this way; the first look is still computed.
    Returns dict(final uint8 [B,h,w], first uint8 [B,S,S], bboxes list)."""
    import torch.nn.functional as F

    from . import looktwice as olt

    S = tuple(image_size)
    label_size = S if label_size is None else tuple(label_size)

    def seg(x):  # backbone + student decoder on the raw token grid (:343-345)
        keys = ovit.keys_to_map(ovit.vit_forward(vit_sd, spec, x)["key_tokens"])
        return odec.baseline_forward(dec_sd, keys, want_ortho=False)[0]

    finals, firsts, boxes = [], [], []
    for i in range(images_u8.shape[0]):
        first = first_stage_eval(vit_sd, spec, dec_sd, images_u8[i:i + 1], S, feature_size)
        logits = first["logits"] if first_logits is None else first_logits[i:i + 1]
        up, bb = olt.process_preds(logits, S, look_twice_th, expand_type)
        firsts.append(up.to(torch.uint8))
        boxes.append(bb)
        if bb is not None:
            up = olt.look_twice(images_u8[i].permute(1, 2, 0).contiguous().numpy(), bb, up, S, seg)
        fin = F.interpolate(up.reshape(1, 1, *S).float(), size=label_size, mode="bilinear", align_corners=False)
        finals.append((fin[0, 0] > 0.5).to(torch.uint8))
    return {"final": torch.stack(finals), "first": torch.cat(firsts), "bboxes": boxes}
