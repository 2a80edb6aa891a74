"""Golden vectors for the COD metric suite: the REFERENCE's `engine.utils.metrics.metric.statistics` (imported from
/root/reference) on seeded (gt, pred) pairs.  Called by tools/make_golden.py."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).resolve().parents[1] / "tests" / "golden"


def metric_cases(seed: int = 0):
    """List of (gt uint8 {0,1}|{0,255}, pred float) pairs of assorted sizes, incl. the degenerate cases."""
    rng = np.random.default_rng(seed)
    cases = []
    for k, (h, w) in enumerate([(97, 131), (240, 180), (64, 64), (150, 333), (200, 200), (77, 90)]):
        yy, xx = np.mgrid[0:h, 0:w]
        gt = np.zeros((h, w), np.float64)
        for _ in range(1 + k % 3):
            cy, cx, ry, rx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w, rng.uniform(0.08, 0.3) * h, rng.uniform(0.08, 0.3) * w
            gt = np.maximum(gt, (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1).astype(np.float64))
        noise = rng.normal(0, 0.35, (h, w))
        soft = np.clip(gt * 0.7 + 0.15 + noise, 0, 1)
        binary = (soft > 0.5).astype(np.float64)
        cases.append((gt * (255 if k % 2 else 1), binary))           # the eval loops feed binary masks
        cases.append((gt, soft))                                      # soft predictions exercise the 256-bin curves
    h, w = 50, 60
    cases.append((np.zeros((h, w)), (rng.random((h, w)) > 0.7).astype(np.float64)))   # empty ground truth
    cases.append((np.ones((h, w)), rng.random((h, w))))                                # full ground truth
    g = np.zeros((h, w)); g[10:30, 20:45] = 1
    cases.append((g, np.zeros((h, w))))                                                # constant prediction (0)
    cases.append((g, np.ones((h, w))))                                                 # constant prediction (1)
    return cases


def gold_metrics():
    from engine.utils.metrics.metric import statistics
    res = {}
    cases = metric_cases()
    st = statistics()
    for i, (gt, pred) in enumerate(cases):
        one = statistics()
        one.step(torch.from_numpy(gt)[None], torch.from_numpy(pred)[None])
        st.step(torch.from_numpy(gt)[None], torch.from_numpy(pred)[None])
        res[f"acc_{i}"] = np.float64(one.ACC.accs[0])
        res[f"iou_{i}"] = np.float64(one.MIOU.ious[0])
        res[f"mae_{i}"] = np.float64(one.MAE.maes[0])
        res[f"sm_{i}"] = np.float64(one.SM.sms[0])
        res[f"em_curve_{i}"] = np.asarray(one.EM.changeable_ems[0], np.float64) * np.ones(256)
        res[f"em_adp_{i}"] = np.float64(one.EM.adaptive_ems[0])
        res[f"fm_curve_{i}"] = np.asarray(one.FM.changeable_fms[0], np.float64)
        res[f"fm_adp_{i}"] = np.float64(one.FM.adaptive_fms[0])
        res[f"wfm_{i}"] = np.float64(one.WFM.weighted_fms[0])
    for k, v in st.get_result().items():
        res["final_" + k] = np.float64(v)
    res["n_cases"] = np.int64(len(cases))
    np.savez_compressed(GOLD / "metrics.npz", **res)
