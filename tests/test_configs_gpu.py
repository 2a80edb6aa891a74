"""BASELINE.json configurations as parity / property tests.

configs[0] (the reference's CPU-runnable case): UCOD-DPL_dinov1 first-stage eval, 8 synthetic images @296^2, shipped
weights/UCOD_DPL_dinov1.safetensors — CUDA path vs the fp32 CPU oracle.
configs[1] at full size (64 images @518^2): size-independent properties — determinism and per-image independence
(an image's mask must not depend on what else is in the batch: the data-parallel sharding relies on it)."""
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch
from safetensors.torch import load_file

from oracle import pipeline as opipe
from oracle import vit as ovit
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.pipeline import FirstStageEval
from ucod_dpl_b200.synth import random_vit_state_dict, synth_batch_u8
from ucod_dpl_b200.vit import spec_for

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _pipe(kind: str, S: int):
    vit_sd = random_vit_state_dict(spec_for(kind), seed=0)
    dec_sd = load_file(str(ROOT / "weights" / f"UCOD_DPL_{kind}.safetensors"))
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(dec_sd, strict=True)
    return FirstStageEval(vit_sd, spec_for(kind), model, (S, S), 68, device="cuda"), vit_sd, dec_sd


def test_config0_dinov1_first_stage_eval_matches_oracle():
    S = 296
    pipe, vit_sd, dec_sd = _pipe("dinov1", S)
    imgs = synth_batch_u8(0, 8, S, S)
    fg = pipe.logits(imgs.cuda()).cpu()
    masks = pipe(imgs.cuda()).cpu()
    ref = opipe.first_stage_eval(vit_sd, ovit.spec_for("dinov1"), dec_sd, imgs, (S, S), 68)
    # bf16 backbone vs fp32 reference.  The random-init ViT-B/8 (no LayerScale) gives low-contrast keys (token std
    # 0.28 vs 0.39 for the DINOv2 config, same absolute bf16 error ~6e-3 rms), so its decoder logits sit closer to
    # the sigmoid = 0.5 boundary than a trained model's: measured max 2.3e-2 / mean 3.7e-3 on the sigmoid outputs and
    # 99.55 % identical mask pixels; the DINOv2 config (smoke(), test_vit_gpu) meets 1e-2 / 99.9 %.
    d = (torch.sigmoid(fg) - torch.sigmoid(ref["logits"])).abs()
    agree = (masks == ref["mask"]).float().mean().item()
    print(f"config0: sigmoid diff max {d.max().item():.4f} mean {d.mean().item():.5f} p99.9 "
          f"{d.flatten().kthvalue(int(d.numel() * 0.999)).values.item():.4f}; mask agreement {agree:.5f}")
    assert d.mean().item() < 5e-3 and d.max().item() < 3e-2
    assert agree >= 0.995


def test_config1_full_size_determinism_and_batch_independence():
    S, B = 518, 64
    pipe, _, _ = _pipe("dinov2", S)
    imgs = synth_batch_u8(0, B, S, S).cuda()
    m1 = pipe(imgs)
    m2 = pipe(imgs)
    assert torch.equal(m1, m2)                                   # bitwise deterministic
    assert 0.0 < m1.float().mean().item() < 1.0                  # not a degenerate mask
    # image 5 alone, and inside a different batch, gives the same logits up to tile-order rounding of the
    # fp32 accumulations (different M tiling), and the same mask except at the sigmoid = 0.5 boundary
    solo = pipe(imgs[5:6])
    other = pipe(torch.cat([imgs[40:47], imgs[5:6]], 0))[-1:]
    assert (solo != m1[5:6]).float().mean().item() < 1e-3
    assert (other != m1[5:6]).float().mean().item() < 1e-3
    lg_full = pipe.logits(imgs)[5]
    lg_solo = pipe.logits(imgs[5:6])[0]
    assert (torch.sigmoid(lg_full) - torch.sigmoid(lg_solo)).abs().max().item() < 2e-3
