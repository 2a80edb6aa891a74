"""Pins oracle/metrics.py against the reference's own `statistics` (tests/golden/metrics.npz). CPU-only."""
from pathlib import Path

import numpy as np

from oracle import metrics as om
from tools.make_golden_metrics import metric_cases

GOLD = Path(__file__).resolve().parents[1] / "tests" / "golden"


def test_metrics_match_reference():
    gold = np.load(GOLD / "metrics.npz")
    cases = metric_cases()
    assert len(cases) == int(gold["n_cases"])
    items = []
    for i, (gt, pred) in enumerate(cases):
        r = om.per_image(gt, pred)
        items.append(r)
        for k in ("acc", "iou", "mae", "sm", "em_adp", "fm_adp", "wfm"):
            np.testing.assert_allclose(r[k], float(gold[f"{k}_{i}"]), rtol=1e-12, atol=1e-14, err_msg=f"{k} case {i}")
        np.testing.assert_allclose(r["em_curve"], gold[f"em_curve_{i}"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(r["fm_curve"], gold[f"fm_curve_{i}"], rtol=1e-12, atol=1e-14)
    final = om.aggregate(items)
    for k, v in final.items():
        np.testing.assert_allclose(v, float(gold["final_" + k]), rtol=1e-12, err_msg=k)
