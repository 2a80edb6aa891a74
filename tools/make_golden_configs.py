#!/usr/bin/env python
"""Golden first-stage-eval outputs of the fp32 CPU oracle for the BASELINE configurations, committed so that the GPU
tests and tools compare against fixed vectors (the oracle itself is pinned to reference-generated goldens by
tests/test_oracle_*.py).  Writes tests/golden/configs_eval.npz:
  c0_logits [8,1,68,68], c0_mask bits — configs[0]: DINO ViT-B/8, 8 synthetic images @296^2, UCOD_DPL_dinov1 weights
  c1_logits [4,1,68,68], c1_mask bits — configs[1]: DINOv2 ViT-B/14, images 0,21,42,63 of the 64-image batch @518^2
Run on the build box: python tools/make_golden_configs.py"""
import sys
from pathlib import Path

import numpy as np
import torch
from safetensors.torch import load_file

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import pipeline as opipe  # noqa: E402
from oracle import vit as ovit  # noqa: E402
from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8  # noqa: E402
from ucod_dpl_b200.vit import spec_for  # noqa: E402

C1_IMAGES = (0, 21, 42, 63)
out = {}
for tag, kind, S, idx in (("c0", "dinov1", 296, tuple(range(8))), ("c1", "dinov2", 518, C1_IMAGES)):
    vit_sd = random_vit_state_dict(spec_for(kind), seed=0)
    dec_sd = load_file(str(ROOT / "weights" / f"UCOD_DPL_{kind}.safetensors"))
    imgs = torch.stack([synth_image_u8(i, S, S) for i in idx])
    logits, masks = [], []
    for i in range(len(idx)):
        r = opipe.first_stage_eval(vit_sd, ovit.spec_for(kind), dec_sd, imgs[i:i + 1], (S, S), 68)
        logits.append(r["logits"])
        masks.append(r["mask"])
        print(tag, idx[i], float(r["mask"].float().mean()), flush=True)
    out[tag + "_logits"] = torch.cat(logits).numpy().astype(np.float32)
    out[tag + "_mask"] = np.packbits(torch.cat(masks).numpy().astype(np.uint8), axis=-1)
    out[tag + "_images"] = np.asarray(idx, np.int32)
np.savez_compressed(ROOT / "tests" / "golden" / "configs_eval.npz", **out)
print("written", {k: v.shape for k, v in out.items()})
