"""ORACLE (test infrastructure — never imported by the product path).

numpy float64 restatement of the COD metric suite the eval loops feed after the hot path
(engine/utils/metrics/metric.py: `_prepare_data` :128-136, ACC :139-158, IoU :160-181, MAE :183-201,
S-measure :203-304, E-measure :306-416, F-measure :418-477, weighted F-measure :479-531, `statistics` :19-74).
Per-image values are returned as a dict so that the device implementation can be checked measure by measure.

Parity pin: tools/make_golden_metrics.py runs the reference's own `statistics` on seeded (gt, pred) pairs.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import convolve, distance_transform_edt

EPS = np.spacing(1)


def prepare(gt: np.ndarray, pred: np.ndarray):
    gt = gt.astype(np.float64)
    pred = pred.astype(np.float64)
    if gt.max() != gt.min():
        gt = (gt - gt.min()) / (gt.max() - gt.min())
    gt = gt > 0.5
    if pred.max() != pred.min():
        pred = (pred - pred.min()) / (pred.max() - pred.min())
    else:
        pred = pred.astype(int)
    return pred, gt


def _ssim(p, g):
    n = p.size
    x, y = p.mean(), g.mean()
    sx = ((p - x) ** 2).sum() / (n - 1)
    sy = ((g - y) ** 2).sum() / (n - 1)
    sxy = ((p - x) * (g - y)).sum() / (n - 1)
    a = 4 * x * y * sxy
    b = (x ** 2 + y ** 2) * (sx + sy)
    if a != 0:
        return a / (b + EPS)
    return 1.0 if b == 0 else 0.0


def _s_object(p, g):
    sel = p[g == 1]
    x = sel.mean()
    sigma = sel.std(ddof=1)
    return 2 * x / (x ** 2 + 1 + sigma + EPS)


def s_measure(p, g, alpha=0.5):
    y = g.mean()
    if y == 0:
        return 1 - p.mean()
    if y == 1:
        return p.mean()
    gf = g.astype(np.float64)
    obj = y * _s_object(p * gf, gf) + (1 - y) * _s_object((1 - p) * (1 - gf), 1 - gf)
    h, w = g.shape
    if np.count_nonzero(g) == 0:
        cx, cy = np.round(w / 2), np.round(h / 2)
    else:
        cy, cx = np.argwhere(g).mean(axis=0).round()
    cx, cy = int(cx) + 1, int(cy) + 1
    area = h * w
    w1, w2, w3 = cx * cy / area, cy * (w - cx) / area, (h - cy) * cx / area
    w4 = 1 - w1 - w2 - w3
    reg = (w1 * _ssim(p[:cy, :cx], gf[:cy, :cx]) + w2 * _ssim(p[:cy, cx:], gf[:cy, cx:]) +
           w3 * _ssim(p[cy:, :cx], gf[cy:, :cx]) + w4 * _ssim(p[cy:, cx:], gf[cy:, cx:]))
    return max(0, alpha * obj + (1 - alpha) * reg)


def _em_from_counts(fg_fg, fg_bg, gt_fg, size):
    """fg_fg / fg_bg: scalars or arrays (per threshold)."""
    pred_fg = fg_fg + fg_bg
    pred_bg = size - pred_fg
    if gt_fg == 0:
        s = pred_bg
    elif gt_fg == size:
        s = pred_fg
    else:
        bg_fg = gt_fg - fg_fg
        bg_bg = pred_bg - bg_fg
        mp, mg = pred_fg / size, gt_fg / size
        combos = [(1 - mp, 1 - mg), (1 - mp, 0 - mg), (0 - mp, 1 - mg), (0 - mp, 0 - mg)]
        s = 0
        for part, (a, b) in zip([fg_fg, fg_bg, bg_fg, bg_bg], combos):
            align = 2 * (a * b) / (a ** 2 + b ** 2 + EPS)
            s = s + ((align + 1) ** 2 / 4) * part
    return s / (size - 1 + EPS)


def _hists(p, g):
    q = (p * 255).astype(np.uint8)
    bins = np.linspace(0, 256, 257)
    fg, _ = np.histogram(q[g], bins=bins)
    bg, _ = np.histogram(q[~g], bins=bins)
    return np.cumsum(np.flip(fg)), np.cumsum(np.flip(bg))


def e_measure(p, g):
    size, gt_fg = g.size, np.count_nonzero(g)
    fg_c, bg_c = _hists(p, g)
    curve = _em_from_counts(fg_c, bg_c, gt_fg, size)
    thr = min(2 * p.mean(), 1.0)
    b = p >= thr
    adp = _em_from_counts(np.count_nonzero(b & g), np.count_nonzero(b & ~g), gt_fg, size)
    return np.asarray(curve, np.float64) * np.ones(256), float(adp)


def f_measure(p, g, beta=0.3):
    thr = min(2 * p.mean(), 1.0)
    b = p >= thr
    inter = b[g].sum()
    if inter == 0:
        adp = 0.0
    else:
        pre, rec = inter / np.count_nonzero(b), inter / np.count_nonzero(g)
        adp = (1 + beta) * pre * rec / (beta * pre + rec)
    fg_c, bg_c = _hists(p, g)
    ps = fg_c + bg_c
    ps[ps == 0] = 1
    t = max(np.count_nonzero(g), 1)
    prec, rec = fg_c / ps, fg_c / t
    num = (1 + beta) * prec * rec
    den = np.where(num == 0, 1, beta * prec + rec)
    return num / den, float(adp), prec, rec


def _gauss7(sigma=5):
    y, x = np.ogrid[-3:4, -3:4]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h / h.sum()


def weighted_f_measure(p, g, beta=1):
    if np.all(~g):
        return 0.0
    dst, idx = distance_transform_edt(g == 0, return_indices=True)
    e = np.abs(p - g)
    et = e.copy()
    et[g == 0] = et[idx[0][g == 0], idx[1][g == 0]]
    ea = convolve(et, weights=_gauss7(), mode="constant", cval=0)
    mn = np.where(g & (ea < e), ea, e)
    bw = np.where(g == 0, 2 - np.exp(np.log(0.5) / 5 * dst), np.ones_like(g, dtype=np.float64))
    ew = mn * bw
    tpw = g.sum() - ew[g == 1].sum()
    fpw = ew[g == 0].sum()
    r = 1 - ew[g == 1].mean()
    pr = tpw / (tpw + fpw + EPS)
    return (1 + beta) * r * pr / (r + beta * pr + EPS)


def per_image(gt: np.ndarray, pred: np.ndarray) -> dict:
    p, g = prepare(gt, pred)
    em_curve, em_adp = e_measure(p, g)
    fm_curve, fm_adp, prec, rec = f_measure(p, g)
    inter, union = np.logical_and(p, g).sum(), np.logical_or(p, g).sum()
    return {"acc": float(np.sum(p == g) / g.size), "iou": float(1.0 if union == 0 else inter / union),
            "mae": float(np.mean(np.abs(p - g))), "sm": float(s_measure(p, g)), "em_curve": em_curve, "em_adp": em_adp,
            "fm_curve": fm_curve, "fm_adp": fm_adp, "wfm": float(weighted_f_measure(p, g))}


def aggregate(items: list) -> dict:
    """`statistics.get_result` (metric.py:61-74)."""
    em = np.mean(np.array([i["em_curve"] for i in items], np.float64), axis=0)
    fm = np.mean(np.array([i["fm_curve"] for i in items], np.float64), axis=0)
    mean = lambda k: float(np.mean(np.array([i[k] for i in items], np.float64)))  # noqa: E731
    return {"ACC": mean("acc"), "mIOU": mean("iou"), "E_MAX": float(em.max()), "E_MEAN": float(em.mean()),
            "F_MAX": float(fm.max()), "F_MEAN": float(fm.mean()), "SMeasure": mean("sm"), "MAE": mean("mae"),
            "WFM": mean("wfm")}
