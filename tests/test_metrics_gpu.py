"""COD metric suite on the device (fp64) vs the reference's `statistics` (tests/golden/metrics.npz) and the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import metrics as om
from tools.make_golden_metrics import metric_cases
from ucod_dpl_b200.engine.utils.metrics.metric import cod_metrics, statistics

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parents[1] / "tests" / "golden"
KEYS = ("acc", "iou", "mae", "sm", "em_adp", "fm_adp", "wfm")


def test_per_image_measures_match_reference():
    gold = np.load(GOLD / "metrics.npz")
    cases = metric_cases()
    st = statistics()
    for i, (gt, pred) in enumerate(cases):
        g = torch.from_numpy(gt).float()[None].cuda()     # the eval loops hold fp32 tensors
        p = torch.from_numpy(pred).float()[None].cuda()
        row = cod_metrics(g, p)[0].cpu().numpy()
        st.step(g, p)
        # the reference sees float64 copies of the same fp32 values
        ref = om.per_image(gt.astype(np.float32), pred.astype(np.float32))
        for k, name in enumerate(KEYS):
            np.testing.assert_allclose(row[k], ref[name], rtol=1e-9, atol=1e-12, err_msg=f"{name} case {i}")
        np.testing.assert_allclose(row[7:263], ref["em_curve"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(row[263:], ref["fm_curve"], rtol=1e-9, atol=1e-12)
        # and the reference-generated golden (float64 inputs; binary / {0,255} cases are exactly representable)
        if (i % 2 == 0 and i < 12) or i in (12, 14, 15):
            for k, name in enumerate(KEYS):
                np.testing.assert_allclose(row[k], float(gold[f"{name}_{i}"]), rtol=1e-9, atol=1e-12,
                                           err_msg=f"gold {name} case {i}")
    res = st.get_result()
    assert set(res) == {"ACC", "mIOU", "E_MAX", "E_MEAN", "F_MAX", "F_MEAN", "SMeasure", "MAE", "WFM"}


def test_batched_call_equals_single_calls_and_feature_transform_ties():
    rng = np.random.default_rng(3)
    B, h, w = 5, 120, 150
    gt = (rng.random((B, h, w)) < 0.02).astype(np.float32)          # sparse points: many equidistant ties
    gt[:, 40:70, 50:90] = 1
    pred = rng.random((B, h, w)).astype(np.float32)
    rows = cod_metrics(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda()).cpu().numpy()
    for b in range(B):
        ref = om.per_image(gt[b], pred[b])
        np.testing.assert_allclose(rows[b, 6], ref["wfm"], rtol=1e-9)   # scipy tie rule reproduced exactly
        np.testing.assert_allclose(rows[b, 3], ref["sm"], rtol=1e-9)
        one = cod_metrics(torch.from_numpy(gt[b:b + 1]).cuda(), torch.from_numpy(pred[b:b + 1]).cuda())[0].cpu().numpy()
        np.testing.assert_allclose(one, rows[b], rtol=1e-12, atol=1e-15)
