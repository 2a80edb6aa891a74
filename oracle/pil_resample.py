"""ORACLE (test infrastructure — never imported by the product path).

numpy restatement of Pillow's `ImagingResample` for 8-bit images (libImaging/Resample.c; Pillow is an un-vendored
dependency of the reference, 12.2.0 installed here), as reached from

* torchvision `transforms.Resize((S,S))` on a PIL image = `Image.resize((S,S), BILINEAR)` — antialiased triangle
  filter (engine/runner/loop_UCOD_DPL.py:282-286,341; data/datasets/transforms.py:12-18)
* `pred_PIL.resize((w,h))` with Pillow's default BICUBIC (a = -0.5) (loop_UCOD_DPL.py:350)

Algorithm: per axis, coefficients in float64 (`precompute_coeffs`), converted to 22-bit fixed point
(`normalize_coeffs_8bpc`), horizontal pass then vertical pass, each pass rounding to uint8
(`clip8((1<<21) + sum(px*k)) >> 22`).  Pinned against PIL itself in tests/test_oracle_golden.py (test_pil_*_bit_exact).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int, filt: str):
    """-> (ksize, bounds int [out,2] = (xmin, count), kk int32 [out, ksize])."""
    fn, fsupport = FILTERS[filt]
    scale = float(np.float32(in1) - np.float32(in0)) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = float(np.float32(in0)) + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        xs = np.arange(n, dtype=np.float64)
        w = fn((xs + xmin - center + 0.5) * ss)
        ww = w.sum() if n > 0 else 0.0
        # Pillow accumulates ww sequentially in double
        ww = 0.0
        for v in w:
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :n] = w
        bounds[xx] = (xmin, n)
    ki = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64),
                  (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int64)
    return ksize, bounds, ki


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass_h(img, bounds, kk, out_w, row0, rows):
    """img [H,W,C] u8 -> [rows, out_w, C] u8 using source rows row0..row0+rows-1."""
    C = img.shape[2]
    out = np.empty((rows, out_w, C), dtype=np.uint8)
    src = img[row0:row0 + rows].astype(np.int64)
    for xx in range(out_w):
        xmin, n = bounds[xx]
        acc = np.full((rows, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        if n > 0:
            acc = acc + np.tensordot(src[:, xmin:xmin + n, :], kk[xx, :n], axes=([1], [0]))
        out[:, xx, :] = _clip8(acc)
    return out


def _pass_v(img, bounds, kk, out_h):
    W, C = img.shape[1], img.shape[2]
    out = np.empty((out_h, W, C), dtype=np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_h):
        ymin, n = bounds[yy]
        acc = np.full((W, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        if n > 0:
            acc = acc + np.tensordot(kk[yy, :n], src[ymin:ymin + n], axes=([0], [0]))
        out[yy] = _clip8(acc)
    return out


def resize_u8(img: np.ndarray, out_w: int, out_h: int, filt: str, box=None) -> np.ndarray:
    """`Image.resize((out_w,out_h), filt, box)` for uint8 [H,W] or [H,W,C] arrays."""
    squeeze = img.ndim == 2
    im = img[:, :, None] if squeeze else img
    H, W = im.shape[:2]
    if box is None:
        box = (0, 0, W, H)
    if (W, H) == (out_w, out_h) and tuple(box) == (0, 0, W, H):
        return img.copy()
    need_h = out_w != W or box[0] != 0 or box[2] != out_w
    need_v = out_h != H or box[1] != 0 or box[3] != out_h
    _, bh, kh = precompute_coeffs(W, box[0], box[2], out_w, filt)
    _, bv, kv = precompute_coeffs(H, box[1], box[3], out_h, filt)
    y_first = int(bv[0, 0])
    y_last = int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
    cur = im
    if need_h:
        bv = bv.copy()
        bv[:, 0] -= y_first
        cur = _pass_h(im, bh, kh, out_w, y_first, y_last - y_first)
    if need_v:
        cur = _pass_v(cur, bv, kv, out_h)
    if not need_h and not need_v:
        cur = im.copy()
    return cur[:, :, 0] if squeeze else cur


def crop_u8(img: np.ndarray, left: int, top: int, right: int, bottom: int) -> np.ndarray:
    """`Image.crop((l,t,r,b))`: out-of-image area is zero filled."""
    H, W = img.shape[:2]
    h, w = max(bottom - top, 0), max(right - left, 0)
    out = np.zeros((h, w) + img.shape[2:], dtype=img.dtype)
    y0, y1 = max(top, 0), min(bottom, H)
    x0, x1 = max(left, 0), min(right, W)
    if y1 > y0 and x1 > x0:
        out[y0 - top:y1 - top, x0 - left:x1 - left] = img[y0:y1, x0:x1]
    return out


def paste_u8(dst: np.ndarray, src: np.ndarray, x: int, y: int) -> None:
    """`Image.paste(src, (x,y))` in place (clipped to dst)."""
    H, W = dst.shape[:2]
    h, w = src.shape[:2]
    y0, y1 = max(y, 0), min(y + h, H)
    x0, x1 = max(x, 0), min(x + w, W)
    if y1 > y0 and x1 > x0:
        dst[y0:y1, x0:x1] = src[y0 - y:y1 - y, x0 - x:x1 - x]
