"""Where does the bf16-vs-fp32 gap of the first-stage eval come from?  Compares, against the committed oracle goldens
(tests/golden/configs_eval.npz): the product pipeline, and the product backbone's fp32 keys pushed through an fp32
decoder (torch on the GPU) — i.e. with the last two bf16 roundings (keys -> bf16, 768->128 GEMM in bf16) removed."""
import sys
from pathlib import Path
from types import SimpleNamespace
import numpy as np
import torch
import torch.nn.functional as F
from safetensors.torch import load_file
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import decoder as odec
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.pipeline import FirstStageEval
from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8
from ucod_dpl_b200.vit import spec_for

g = np.load(ROOT / "tests" / "golden" / "configs_eval.npz")
for tag, kind, S in (("c0", "dinov1", 296), ("c1", "dinov2", 518)):
    idx = g[tag + "_images"].tolist()
    vit_sd = random_vit_state_dict(spec_for(kind), seed=0)
    dec_sd = load_file(str(ROOT / "weights" / f"UCOD_DPL_{kind}.safetensors"))
    model = baseline(SimpleNamespace(dim=768)); model.load_state_dict(dec_sd, strict=True)
    pipe = FirstStageEval(vit_sd, spec_for(kind), model, (S, S), 68, device="cuda")
    imgs = torch.stack([synth_image_u8(i, S, S) for i in idx]).cuda()
    ref = torch.from_numpy(g[tag + "_logits"])
    ref_mask = torch.from_numpy(np.unpackbits(g[tag + "_mask"], axis=-1)[..., :S])
    def report(name, fg):
        fg = fg.float().cpu()
        d = (torch.sigmoid(fg) - torch.sigmoid(ref)).abs()
        up = F.interpolate(fg, size=(S, S), mode="bilinear", align_corners=False)
        mask = (torch.sigmoid(up) > 0.5)[:, 0].to(torch.uint8)
        agree = (mask == ref_mask).float().mean().item()
        upr = F.interpolate(ref, size=(S, S), mode="bilinear", align_corners=False)[:, 0]
        bad = mask != ref_mask
        worst = upr[bad].abs().max().item() if bad.any() else 0.0
        print(f"{tag} {name:34s} sigmoid diff max {d.max():.4f} mean {d.mean():.5f} | mask agreement {agree:.5f} "
              f"| max |oracle logit| at a mismatching pixel {worst:.4f}  (logit rms {ref.pow(2).mean().sqrt():.3f})")
    report("product (bf16 keys, bf16 GEMM)", pipe.logits(imgs))
    k32, _, _ = pipe.extractor.keys(imgs, want_f32=True, want_bf16=False)
    p = spec_for(kind).patch
    gh = S // p
    keys = k32.reshape(len(idx), gh, gh, 768).permute(0, 3, 1, 2).contiguous()
    sd_cuda = {k: v.cuda() for k, v in dec_sd.items()}
    feats = odec.upsample_bilinear(keys, (68, 68))
    report("fp32 keys -> fp32 decoder", odec.baseline_forward(sd_cuda, feats, want_ortho=False)[0])
    report("bf16-rounded keys -> fp32 decoder", odec.baseline_forward(sd_cuda, odec.upsample_bilinear(keys.bfloat16().float(), (68, 68)), want_ortho=False)[0])
