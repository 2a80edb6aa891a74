"""Pins oracle/train.py (one first-stage training step: loss, gradients, AdamW, StepLR, EMA) against the reference's
own TrainLoop._process_batch run by tools/make_golden_train.py.  CPU-only."""
from pathlib import Path

import numpy as np
import torch
from safetensors.torch import load_file

from oracle import decoder as odec
from oracle import train as otr

ROOT = Path(__file__).resolve().parents[1]


def train_inputs(seed: int, B: int = 2):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, 768, 37, 37, generator=g)
    pl = (torch.rand(B, 1, 16, 16, generator=g) < 0.35).float()
    return feats, pl


def test_three_training_steps_match_reference():
    gold = np.load(ROOT / "tests" / "golden" / "train.npz")
    sd = {k: v.clone() for k, v in load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")).items()}
    dis_sd = odec.random_discriminator_state_dict(68, seed=31)
    state = otr.new_state(sd)
    gstep = 0
    for step in range(3):
        feats, pl = train_inputs(100 + step)
        out = otr.train_step(sd, dis_sd, state, feats, pl, cur_epoch=3, global_step=gstep,
                             lr=otr.step_lr(2e-4, step))
        gstep += 2  # the reference bumps global_step in _process_batch and again in run_epoch
        np.testing.assert_allclose(float(out["loss"]), float(gold[f"loss_{step}"]), rtol=2e-5)
        if step == 0:
            for k in otr.PARAM_ORDER:
                g = gold["grad0_" + k]
                np.testing.assert_allclose(out["grads"][k].numpy(), g, atol=2e-6 + 1e-3 * np.abs(g).max())
            # F.normalize makes the output independent of |learnable_embedding|: its gradient is rounding noise
            assert np.abs(gold["grad0_learnable_embedding"]).max() < 1e-6
    for k, v in sd.items():
        np.testing.assert_allclose(v.numpy(), gold["final_" + k], atol=2e-6, rtol=1e-4, err_msg=k)


def dis_inputs(seed: int, B: int = 4):
    g = torch.Generator().manual_seed(seed)
    pseudo = (torch.rand(B, 1, 68, 68, generator=g) < 0.4).float()
    student = (torch.rand(B, 1, 68, 68, generator=g) < torch.rand(B, 1, 1, 1, generator=g)).float()
    return pseudo, student


def test_three_discriminator_steps_match_reference():
    gold = np.load(ROOT / "tests" / "golden" / "train.npz")
    dis_sd = odec.random_discriminator_state_dict(68, seed=31)
    state = otr.new_dis_state(dis_sd)
    for step in range(3):
        pseudo, student = dis_inputs(200 + step)
        out = otr.discriminator_step(dis_sd, state, pseudo, student, lr=otr.step_lr(1e-3, step))
        np.testing.assert_allclose(float(out["loss"]), float(gold[f"dis_loss_{step}"]), rtol=2e-5)
        if step == 0:
            for k in otr.DIS_PARAM_ORDER:
                g = gold["dis_grad0_" + k]
                np.testing.assert_allclose(out["grads"][k].numpy(), g, atol=1e-6 + 1e-4 * np.abs(g).max())
    for k in otr.DIS_PARAM_ORDER + ["maskConv.layers.1.running_mean", "convs.1.layers.1.running_var"]:
        np.testing.assert_allclose(dis_sd[k].numpy(), gold["dis_final_" + k], atol=5e-6, rtol=1e-4, err_msg=k)
