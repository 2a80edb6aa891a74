import sys; sys.path.insert(0, '.')
import torch
from oracle import vit as ovit
from ucod_dpl_b200.synth import random_vit_state_dict, synth_batch_u8
from ucod_dpl_b200.vit import VitKeyExtractor, spec_for
for kind, S in (("dinov2", 518), ("dinov1", 296), ("dinov2", 224)):
    sd = random_vit_state_dict(spec_for(kind), seed=0)
    u8 = synth_batch_u8(0, 2, S, S)
    ref = ovit.vit_forward(sd, ovit.spec_for(kind), ovit.normalize_u8(u8))["key_tokens"][:, 1:]
    ext = VitKeyExtractor(sd, spec_for(kind))
    k32, _, _ = ext.keys(u8.cuda(), want_f32=True)
    k32 = k32.cpu()
    err = (k32 - ref)
    # variation of the keys across tokens (what the decoder can use) vs error
    tok_std = (ref - ref.mean(1, keepdim=True)).std().item()
    print(kind, S, "key abs max", ref.abs().max().item(), "std over tokens", tok_std, "err rms", err.pow(2).mean().sqrt().item(),
          "err max", err.abs().max().item(), "err rms / token std", err.pow(2).mean().sqrt().item() / tok_std)
