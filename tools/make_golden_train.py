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
