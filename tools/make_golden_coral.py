"""Golden vectors for the CORAL SparseRefiner: the REFERENCE's own `models.UDLR.SparseRefiner` (imported from
/root/reference through the shims of tools/make_golden.py) with seeded weights / inputs.  Called by make_golden.py."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def gold_coral():
    from engine.config.config import CfgNode as RefCfg
    from engine.runner.loop_CORAL import LocalRefineValidationLoop as Loop
    from models.UDLR import SparseRefiner

    from ucod_dpl_b200.synth import random_refiner_state_dict, synth_coral_inputs
    res = {}
    sd = random_refiner_state_dict(seed=0)
    ref = SparseRefiner.from_config(RefCfg({"window_size": 3, "threshold": 0.0015}))
    missing, unexpected = ref.load_state_dict(sd, strict=True)
    ref.eval()
    res["state_dict_keys"] = np.array(sorted(ref.state_dict().keys()))
    for tag, seed, batch, unc in (("a", 5, 1, ((0, 1), (1, 1), (2, 0))), ("b", 6, 2, ((1, 2),))):
        l, h, preds = synth_coral_inputs(seed, batch=batch, uncertain=unc)
        if tag == "b":
            preds[1] = -12.0          # second image: nothing selected at all
        with torch.no_grad():
            out, ex_loss, opt = ref(l, h, preds)
        assert ex_loss == 0
        res[tag + "_out"] = out.numpy()
        res[tag + "_mask"] = opt["mask"].numpy()
        res[tag + "_coords"] = opt["coords_list"].numpy()
        res[tag + "_window_preds"] = opt["window_preds"].numpy()
        res[tag + "_h_preds"] = opt["h_preds"].numpy()
        res[tag + "_ge_w"] = opt["GE_w"].numpy()
        res[tag + "_entropy"] = opt["entropy"].numpy()
    # probabilities instead of logits take the other branch of the selector (ASR.py:42-45)
    l, h, preds = synth_coral_inputs(7, uncertain=((2, 2),))
    with torch.no_grad():
        out, _, opt = ref(l, h, preds.sigmoid())
    res["c_out"], res["c_mask"] = out.numpy(), opt["mask"].numpy()
    # loop glue
    g = torch.Generator().manual_seed(9)
    p4 = torch.randn(2, 4, 1, 68, 68, generator=g)
    res["concate_preds_out"] = Loop.concate_preds(None, p4).numpy()
    x = torch.randn(1, 1, 168, 168, generator=g)
    res["center_pad_out"] = Loop._center_pad(None, x).numpy()
    res["process_preds_out"] = Loop.process_preds(None, x, (300, 417)).numpy()
    np.savez_compressed(GOLD / "coral.npz", **res)
