"""The oracle's restatement of the Look-Twice loop against goldens produced by the REFERENCE'S OWN LOOP CODE as the
caller (tools/make_golden_refcaller.py: `ValLoop_Look_Twice.process_preds` / `look_twice` / the final resize of
`run`, over the reference's `backbone.forward` and `baseline`)."""
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from safetensors.torch import load_file

from oracle import decoder as odec
from oracle import looktwice as olt
from oracle import vit as ovit
from ucod_dpl_b200.synth import random_vit_state_dict
from ucod_dpl_b200.vit import spec_for

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden" / "refcaller_looktwice.npz"
S = 224


def test_oracle_loop_matches_reference_loop():
    g = np.load(GOLD)
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
    spec = ovit.spec_for("dinov2")
    for ci in range(2):
        img, logits = g[f"c{ci}_image"], torch.from_numpy(g[f"c{ci}_logits"])
        up, boxes = olt.process_preds(logits, (S, S), 0.15, "dynamic")
        assert boxes == g[f"c{ci}_boxes"].tolist()
        assert np.array_equal(np.packbits(up[0].numpy().astype(np.uint8)), g[f"c{ci}_first"])
        seen = []

        def seg(x):
            keys = ovit.keys_to_map(ovit.vit_forward(vit_sd, spec, x)["key_tokens"])
            out = odec.baseline_forward(dec_sd, keys, want_ortho=False)[0]
            seen.append(out)
            return out

        new = olt.look_twice(img, boxes, up, (S, S), seg)
        second = torch.cat(seen).numpy()
        assert np.abs(second - g[f"c{ci}_second_logits"]).max() < 2e-3          # HF modules vs the oracle ViT, both fp32
        got = np.rint(new[0].numpy() * 255).astype(np.uint8)
        assert (got == g[f"c{ci}_new_mask"]).mean() >= 0.9999
        h, w = img.shape[:2]
        final = (F.interpolate(new.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0) > 0.5)[0].numpy()
        want = np.unpackbits(g[f"c{ci}_final"]).reshape(-1)[: h * w].reshape(h, w).astype(bool)
        assert (final == want).mean() >= 0.9999
