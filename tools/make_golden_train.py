"""Golden vectors for the first-stage training step: the REFERENCE's `TrainLoop._process_batch` /
`update_ema_decoder` (engine/runner/loop_UCOD_DPL.py:148-191) run unbound on CPU with the reference's `baseline`,
`Discriminator`, torch AdamW + StepLR, on seeded inputs.  Called by tools/make_golden.py."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


def train_inputs(seed: int, B: int = 2):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, 768, 37, 37, generator=g)
    pl = (torch.rand(B, 1, 16, 16, generator=g) < 0.35).float()
    return feats, pl


def gold_train():
    from safetensors.torch import load_file

    from engine.config.config import CfgNode as RefCfg
    from engine.runner.loop_UCOD_DPL import TrainLoop
    from models.discriminator import Discriminator
    from models.uscod import baseline

    from oracle import decoder as odec
    res = {}
    sd = load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
    model = baseline(RefCfg({"dim": 768}))
    model.load_state_dict(sd, strict=True)
    model.train()
    D = Discriminator(RefCfg({"dis_use_features": False, "dim": 768, "feature_size": 68}))
    D.load_state_dict(odec.random_discriminator_state_dict(68, seed=31), strict=True)
    D.train()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=25, gamma=0.95)
    log = SimpleNamespace(log=lambda *a, **k: None)
    fake = SimpleNamespace(
        runner=SimpleNamespace(model=model, discriminator=D, optimizer=opt, lr_scheduler=sched, logger=log,
                               accelerator=SimpleNamespace(backward=lambda l: l.backward())),
        cfg=SimpleNamespace(model_cfg=SimpleNamespace(feature_size=68)), criterion=nn.BCEWithLogitsLoss(),
        dis_loss=nn.BCELoss(), finetune=False, global_step=0, _cur_epoch=3, _max_epoch=25, _start_finetune=-5,
        ema_alpha=0.99)
    fake.merge_pseudo_label = lambda *a: TrainLoop.merge_pseudo_label(fake, *a)
    fake.update_ema_decoder = lambda: TrainLoop.update_ema_decoder(fake)
    orig_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else orig_to(self, *a, **k)
    try:
        for step in range(3):
            feats, pl = train_inputs(100 + step)
            batch = {"pseudo_labels": pl, "labels": None, "features": feats, "paths": None}
            loss = TrainLoop._process_batch(fake, batch)
            fake.global_step += 1            # run_epoch bumps it a second time (loop_UCOD_DPL.py:143)
            res[f"loss_{step}"] = np.float32(loss.item())
            if step == 0:
                for n, p in model.decoder.named_parameters():
                    res["grad0_" + n] = p.grad.detach().numpy().copy()
    finally:
        torch.Tensor.to = orig_to
    for n, p in model.state_dict().items():
        res["final_" + n] = p.detach().numpy().copy()
    np.savez_compressed(GOLD / "train.npz", **res)


def dis_inputs(seed: int, B: int = 4):
    g = torch.Generator().manual_seed(seed)
    pseudo = (torch.rand(B, 1, 68, 68, generator=g) < 0.4).float()
    student = (torch.rand(B, 1, 68, 68, generator=g) < torch.rand(B, 1, 1, 1, generator=g)).float()
    return pseudo, student


def gold_discriminator_train():
    """The reference's `Discriminator` module + BCELoss + torch AdamW/StepLR driven exactly like
    `TrainLoop.Discriminator_epoch` (loop_UCOD_DPL.py:230-255) for three iterations."""
    from engine.config.config import CfgNode as RefCfg
    from models.discriminator import Discriminator

    from oracle import decoder as odec
    res = dict(np.load(GOLD / "train.npz"))
    D = Discriminator(RefCfg({"dis_use_features": False, "dim": 768, "feature_size": 68}))
    D.load_state_dict(odec.random_discriminator_state_dict(68, seed=31), strict=True)
    D.train()
    for p in D.parameters():
        p.requires_grad = True
    opt = torch.optim.AdamW(D.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=25, gamma=0.95)
    bce = nn.BCELoss()
    for step in range(3):
        pseudo, student = dis_inputs(200 + step)
        opt.zero_grad()
        B = pseudo.shape[0]
        label = torch.cat((torch.zeros(B), torch.ones(B)), dim=-1).unsqueeze(-1)
        probs_pseudo = D(pseudo, None)
        probs_student = D(student, None)
        loss = bce(torch.cat((probs_student, probs_pseudo), dim=0), label)
        loss.backward()
        if step == 0:
            for n, p in D.named_parameters():
                res["dis_grad0_" + n] = p.grad.detach().numpy().copy()
        opt.step()
        sched.step()
        res[f"dis_loss_{step}"] = np.float32(loss.item())
    for n, p in D.state_dict().items():
        res["dis_final_" + n] = p.detach().numpy().copy()
    np.savez_compressed(GOLD / "train.npz", **res)
