"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the Look-Twice stage of `ValLoop_Look_Twice` (engine/runner/loop_UCOD_DPL.py):

* `process_preds`  — :354-384  (bilinear to image size, sigmoid > 0.5, 8-conn CC, area fractions, boxes)
* `expand_bbox`    — :399-417  (python float64 maths, int() truncation; the `br = h*y/(H*W)` quirk is kept)
* `resize_bbox`    — :387-397
* `look_twice`     — :326-352  (PIL crop -> Resize (antialiased bilinear) -> backbone -> decoder@37^2 ->
                                sigmoid > 0.5 -> u8 -> PIL bicubic resize -> paste)

cv2 / PIL calls are replaced by oracle/cc.py and oracle/pil_resample.py (both pinned against the real libraries).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cc, pil_resample as pr

DEFAULT_BOX = [129, 129, 259, 259]  # loop_UCOD_DPL.py:370 (hard-coded, also for 296^2 inputs)


def expand_bbox(mask: np.ndarray, bbox, img_width, img_height, expand_type="const", scale=1.3):
    x, y, w, h = bbox
    if expand_type == "dynamic":
        fr = mask[y:y + h, x:x + w].sum() / (h * w)
        br = (h * y) / (mask.shape[-2] * mask.shape[-1])
        scale = math.sqrt(1 - br / fr + 1)
    new_w = w * scale
    new_h = h * scale
    new_x = x - (new_w - w) / 2
    new_y = y - (new_h - h) / 2
    new_x = max(0, new_x)
    if new_x + new_w > img_width:
        new_x = img_width - new_w
    new_y = max(0, new_y)
    if new_y + new_h > img_height:
        new_y = img_height - new_h
    return [int(new_x), int(new_y), int(new_w), int(new_h)]


def resize_bbox(bbox, original_width, original_height, new_width, new_height):
    x, y, w, h = bbox
    ws = new_width / original_width
    hs = new_height / original_height
    return [int(x * ws), int(y * hs), int(w * ws), int(h * hs)]


@torch.no_grad()
def process_preds(preds: torch.Tensor, img_size, look_twice_th: float, expand_type: str = "dynamic"):
    """preds [1,1,fs,fs] logits -> (preds_up [1,S,S] float {0,1}, bboxes list | None)."""
    h, w = img_size
    up = F.interpolate(preds.float(), size=(h, w), mode="bilinear", align_corners=False)[..., :h, :w]
    preds_up = (torch.sigmoid(up) > 0.5).squeeze(0).float()
    return preds_up, boxes_from_mask(preds_up.numpy(), img_size, look_twice_th, expand_type)


def boxes_from_mask(mask01, img_size, look_twice_th: float, expand_type: str = "dynamic"):
    """The integer half of `process_preds` (:362-384) for a given binarised mask ([1,S,S] or [S,S], values {0,1}):
    8-connected components, area fractions, boundingRect + expand_bbox, sort.  Returns the box list or None."""
    h, w = img_size
    np_mask = (np.asarray(mask01, dtype=np.float32) * 255).astype(np.uint8)
    if np_mask.ndim == 3:
        np_mask = np_mask.squeeze(0)
    num_labels, labels = cc.connected_components_8(np_mask)
    p = [(labels == i).sum() / (h * w) for i in range(1, num_labels)]
    if len(p) == 0:
        return [list(DEFAULT_BOX)]
    if max(p) < look_twice_th:
        bboxes = []
        for i in range(1, num_labels):
            if p[i - 1] > 0.01:
                binary = (labels == i).astype(np.uint8)
                bboxes.append(expand_bbox(binary, cc.bounding_rect(binary), h, w, expand_type=expand_type))
        bboxes = sorted(bboxes, key=lambda b: -1 * b[2] * b[3])
        return bboxes
    return None


def crop_resize_normalize(image_u8_hwc: np.ndarray, box_xywh, out_size) -> torch.Tensor:
    """PIL crop -> torchvision Resize(out_size) (PIL bilinear, antialias) -> ToTensor -> Normalize; [3,S,S] fp32."""
    x, y, w, h = box_xywh
    crop = pr.crop_u8(image_u8_hwc, x, y, x + w, y + h)
    S_h, S_w = out_size
    res = pr.resize_u8(crop, S_w, S_h, "bilinear")
    t = torch.from_numpy(res).permute(2, 0, 1).float() / 255.0
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    return (t - mean) / std, res


@torch.no_grad()
def look_twice(image_u8_hwc: np.ndarray, bboxes, old_mask: torch.Tensor, img_size, segment_fn):
    """image_u8_hwc: original image [H0,W0,3]; old_mask [1,S,S] float {0,1};
    segment_fn(x [1,3,S,S] fp32 normalised) -> fg logits [1,1,g,g] (backbone + student decoder at the raw grid).
    Returns new mask [1,S,S] float in [0,1] (u8/255)."""
    ih, iw = img_size
    H0, W0 = image_u8_hwc.shape[:2]
    new_mask = (old_mask.squeeze(0).numpy() * 255).astype(np.uint8).copy()
    for bbox in bboxes:
        nb = resize_bbox(bbox, iw, ih, W0, H0)
        x, _ = crop_resize_normalize(image_u8_hwc, nb, (ih, iw))
        logits = segment_fn(x.unsqueeze(0))
        pred = (torch.sigmoid(logits.float().reshape(logits.shape[-2], logits.shape[-1])) > 0.5).float()
        pred_u8 = (pred * 255).to(torch.uint8).numpy()
        pasted = pr.resize_u8(pred_u8, bbox[-2], bbox[-1], "bicubic")
        pr.paste_u8(new_mask, pasted, bbox[0], bbox[1])
    return torch.from_numpy(new_mask).float().div(255.0).unsqueeze(0)
