#!/usr/bin/env python
"""Golden vectors with the REFERENCE'S OWN LOOP CODE as the caller: `ValLoop_Look_Twice.process_preds` + `look_twice`
(+ the final resize / threshold of `run`, engine/runner/loop_UCOD_DPL.py:297-352) executed here on the CPU, unmodified,
over the reference's own `backbone.forward` (data/utils/feature_extractor.py:49-59) and `baseline` (models/uscod.py) —
the HF DINOv2 model carries this repo's seeded weights, the decoder the shipped checkpoint.
The only shims: packages absent offline (tools/make_golden.py: install_shims), `backbone.__init__` (hub download +
hard .cuda()) replaced by attaching a locally built HF model, and `.to('cuda')` made a no-op.
Writes tests/golden/refcaller_looktwice.npz.   Run on the build box: python tools/make_golden_refcaller.py"""
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F
import transformers  # noqa: F401  (must precede the timm stub)
from PIL import Image
from safetensors.torch import load_file

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tools.make_golden import GOLD, install_shims  # noqa: E402
from ucod_dpl_b200.synth import planted_object_logits, random_vit_state_dict, synth_image_u8  # noqa: E402
from ucod_dpl_b200.vit import spec_for  # noqa: E402

S, FS = 224, 68
CASES = [dict(seed=7, hw=(300, 340), objects=2), dict(seed=8, hw=(448, 290), objects=3)]


def main():
    install_shims()
    from torchvision import transforms
    from transformers import Dinov2Config, Dinov2Model

    from data.utils.feature_extractor import backbone as RefBackbone
    from engine.runner.loop_UCOD_DPL import ValLoop_Look_Twice
    from models.uscod import baseline as RefBaseline

    # .to('cuda') / .cuda() are no-ops on this CPU box
    _to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else _to(self, *a, **k)

    cfg = Dinov2Config(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, patch_size=14, image_size=518,
                       layer_norm_eps=1e-6)
    hf = Dinov2Model(cfg)
    missing, unexpected = hf.load_state_dict(random_vit_state_dict(spec_for("dinov2"), seed=0), strict=False)
    assert not unexpected, unexpected
    hf.eval()
    bb = RefBackbone.__new__(RefBackbone)
    torch.nn.Module.__init__(bb)
    bb.config = SimpleNamespace(backbone="facebook/dinov2-base")
    bb.feature_extractor = hf
    bb.key = None
    hf.encoder.layer[-1].attention.attention.key.register_forward_hook(bb.hook_fn_key)   # feature_extractor.py:42

    model = RefBaseline(SimpleNamespace(dim=768))
    print(model.load_state_dict(load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))))
    model.eval()
    seen = []
    orig_forward = model.forward
    model.forward = lambda x, *a, **k: (lambda out: (seen.append(out[0].detach().clone()), out)[1])(orig_forward(x, *a, **k))

    loop = SimpleNamespace(
        cfg=SimpleNamespace(val_cfg=SimpleNamespace(look_twice_th=0.15, expand_type="dynamic", look_twice=True)),
        img_size=(S, S), feature_extractor=bb, runner=SimpleNamespace(model=model),
        transform_image=transforms.Compose([transforms.Resize((S, S)), transforms.ToTensor(),
                                            transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]),
        to_PIL=transforms.ToPILImage(), to_tensor=transforms.ToTensor())
    for name in ("expand_bbox", "resize_bbox"):
        setattr(loop, name, (lambda n: lambda *a, **k: getattr(ValLoop_Look_Twice, n)(loop, *a, **k))(name))

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for ci, case in enumerate(CASES):
            h, w = case["hw"]
            img = synth_image_u8(case["seed"], h, w).permute(1, 2, 0).contiguous().numpy()
            path = str(Path(tmp) / f"img_{ci}.png")
            Image.fromarray(img).save(path)
            logits = planted_object_logits(900 + ci, FS, case["objects"])[None]          # [1,1,fs,fs]
            seen.clear()
            with torch.no_grad():
                preds_up, bboxes = ValLoop_Look_Twice.process_preds(loop, logits, None)   # :354-384
                new_mask = ValLoop_Look_Twice.look_twice(loop, path, bboxes, preds_up)     # :326-352
                final = F.interpolate(new_mask.unsqueeze(0), size=(h, w), mode="bilinear").squeeze(0) > 0.5   # :315-317
            out[f"c{ci}_image"] = img
            out[f"c{ci}_logits"] = logits.numpy()
            out[f"c{ci}_first"] = np.packbits(preds_up[0].numpy().astype(np.uint8))
            out[f"c{ci}_boxes"] = np.asarray(bboxes, np.int32)
            out[f"c{ci}_second_logits"] = torch.cat(seen).numpy().astype(np.float32)       # [n_boxes,1,16,16]
            out[f"c{ci}_new_mask"] = np.rint(new_mask[0].numpy() * 255).astype(np.uint8)
            out[f"c{ci}_final"] = np.packbits(final[0].numpy().astype(np.uint8))
            print("case", ci, "boxes", bboxes, "second logits", out[f"c{ci}_second_logits"].shape)
    np.savez_compressed(GOLD / "refcaller_looktwice.npz", **out)
    print("written", GOLD / "refcaller_looktwice.npz")


if __name__ == "__main__":
    main()
