import sys; sys.path.insert(0, '.')
from pathlib import Path
from types import SimpleNamespace
import torch
from safetensors.torch import load_file
from oracle import decoder as odec, train as otr
from ucod_dpl_b200 import ops
from ucod_dpl_b200.models.discriminator import Discriminator
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.train import FirstStageTrainer
sd = load_file("weights/UCOD_DPL_dinov2.safetensors")
model = baseline(SimpleNamespace(dim=768)); model.load_state_dict(sd, strict=True)
D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68))
dsd = odec.random_discriminator_state_dict(68, seed=31); D.load_state_dict(dsd, strict=True)
model.cuda().train(); D.cuda().train()
g = torch.Generator().manual_seed(100)
feats = torch.randn(2, 768, 37, 37, generator=g); pl = (torch.rand(2, 1, 16, 16, generator=g) < 0.35).float()
tr = FirstStageTrainer(model, D, lr0=2e-4); tr.cur_epoch = 3
loss = tr.process_batch(ops.features_to_tokens_bf16(feats.cuda()), (37, 37), pl.cuda())
print("gpu loss", float(loss), "bce", tr.last["bce"].tolist(), "ortho", float(tr.last["ortho"]), "dis", float(tr.last["dis_loss"]))
sd2 = {k: v.clone() for k, v in sd.items()}
st = otr.new_state(sd2)
out = otr.train_step(sd2, dsd, st, feats, pl, cur_epoch=3, global_step=0, lr=2e-4)
print("ref loss", float(out["loss"]), "ortho", float(out["ortho"]), "dis", float(out["dis_loss"]), "w", out["weight"].flatten().tolist())
from ucod_dpl_b200.models.discriminator import merge_pseudo_label
print("gpu w", merge_pseudo_label.last["weight"].flatten().tolist(), "ps", merge_pseudo_label.last["p_s"].flatten().tolist(), "pp", merge_pseudo_label.last["p_p"].flatten().tolist())
print("merged diff", (tr.last["merged"].cpu() - out["merged"]).abs().max().item())
