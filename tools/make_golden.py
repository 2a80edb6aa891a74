#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by running the REFERENCE's own code (imported read-only from /root/reference)
and the real third-party libraries it calls (HF transformers, cv2, Pillow) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):  python tools/make_golden.py
Import shims follow SURVEY.md §8(c): `transformers` first, then stub modules for packages the reference imports
but that are absent here and never touch hot-path arithmetic (timm, prettytable, pytz, ntplib, accelerate,
matplotlib).
"""
from __future__ import annotations

import json
import sys
import types
import zoneinfo
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import transformers  # noqa: F401  (must be imported before the timm stub exists)

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x), trunc_normal_=lambda *a, **k: None)
    _stub("timm.models.registry", register_model=lambda f: f)
    _stub("timm.models.vision_transformer", _cfg=lambda **k: {})
    _stub("prettytable", PrettyTable=object)
    _stub("pytz", timezone=lambda n: zoneinfo.ZoneInfo(n))
    _stub("ntplib")
    _stub("accelerate", Accelerator=object)
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patches")
    sys.path.insert(0, str(REF))


def versions():
    import cv2
    import PIL
    return {"torch": torch.__version__, "transformers": transformers.__version__, "cv2": cv2.__version__,
            "pillow": PIL.__version__, "numpy": np.__version__}


# ------------------------------------------------------------------------------------------------
def gold_configs():
    from engine.config.config import CfgNode as RefCfg

    def plain(d):
        return {k: (plain(v) if isinstance(v, dict) else (list(v) if isinstance(v, tuple) else v)) for k, v in d.items()}

    out = {}
    for f in sorted((REF / "configs" / "uscod").glob("*.py")):
        out[f.name] = plain(RefCfg.load_with_base(str(f)))
    (GOLD / "configs.json").write_text(json.dumps(out, indent=1, sort_keys=True))


def gold_vit():
    """HF Dinov2Model / ViTModel loaded with the repo's seeded weights; the reference's hook + kwargs
    (data/utils/feature_extractor.py:42,49-54; generate_pseudo_label.py:76-77,111-112)."""
    from transformers import Dinov2Config, Dinov2Model, ViTConfig, ViTModel

    from oracle import vit as ovit
    from ucod_dpl_b200.synth import random_vit_state_dict, synth_batch_u8
    res = {}
    for kind, S in (("dinov2", 224), ("dinov2", 518), ("dinov1", 296)):
        spec = ovit.spec_for(kind)
        sd = random_vit_state_dict(spec, seed=0)
        if kind == "dinov2":
            cfg = Dinov2Config(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, patch_size=14,
                               image_size=518, layer_norm_eps=1e-6)
            cfg._attn_implementation = "eager"
            model = Dinov2Model(cfg)
        else:
            cfg = ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, patch_size=8,
                            image_size=224)
            cfg._attn_implementation = "eager"
            model = ViTModel(cfg, add_pooling_layer=False)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all("pooler" in m for m in missing), (missing, unexpected)
        model.eval()
        key = {}
        model.encoder.layer[-1].attention.attention.key.register_forward_hook(
            lambda m, i, o: key.__setitem__("k", o.detach()))
        x = ovit.normalize_u8(synth_batch_u8(0, 1, S, S))
        with torch.no_grad():
            if kind == "dinov2":
                out = model(x, output_attentions=True)
            else:
                out = model(x, interpolate_pos_encoding=True, output_attentions=True)
        k = key["k"][0]  # [T,768]
        att = out.attentions[-1][0, :, 0, 1:]  # [12,P]
        tag = f"{kind}_{S}"
        res[tag + "_key_rows"] = k[::97].numpy().astype(np.float32)      # every 97th token, all channels
        res[tag + "_key_cols"] = k[:, ::61].numpy().astype(np.float32)   # all tokens, every 61st channel
        res[tag + "_cls_attn"] = att.numpy().astype(np.float32)
    np.savez_compressed(GOLD / "vit_hf.npz", **res)


def planted(B, P=256, nh=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    keys = torch.empty(B, P, nh * 64)
    att = torch.empty(B, nh, P)
    for b in range(B):
        nc = 2 + (b % 2)
        centres = torch.randn(nc, nh * 64, generator=g)
        assign = torch.randint(0, nc, (P,), generator=g)
        keys[b] = centres[assign] + 0.3 * torch.randn(P, nh * 64, generator=g)
        logits = torch.randn(nh, nc, generator=g)[:, assign] * 2 + 0.3 * torch.randn(nh, P, generator=g)
        att[b] = torch.softmax(torch.cat([torch.zeros(nh, 1), logits], 1), dim=1)[:, 1:]
    return att, keys


def gold_pseudo_label():
    from data.utils.found_bkg_mask import compute_img_bkg_seg
    import generate_pseudo_label as gpl
    res = {}
    att, keys = planted(4, seed=11)
    T = 257
    g = torch.Generator().manual_seed(12)
    full_att = torch.rand(4, 12, T, T, generator=g)
    full_att[:, :, 0, 1:] = att
    feats = torch.cat([torch.randn(4, 1, 768, generator=g), keys], 1)
    bk, sm = [], []
    for b in range(4):  # the reference runs B = 1
        m, s = compute_img_bkg_seg(full_att[b:b + 1], feats[b:b + 1], (16, 16), 0.6, dim=64)
        bk.append(m[0].numpy()), sm.append(s[0].numpy())
    res["score_bkg"], res["score_sim"] = np.stack(bk), np.stack(sm)
    mb, sb = compute_img_bkg_seg(full_att, feats, (16, 16), 0.6, dim=64)  # batched call: batch-global sim max
    res["score_bkg_batched"], res["score_sim_batched"] = mb.numpy(), sb.numpy()
    rng = np.random.default_rng(5)
    masks = []
    m = np.zeros((16, 16), np.uint8); masks.append(m.copy())
    m = np.ones((16, 16), np.uint8); masks.append(m.copy())
    m = np.zeros((16, 16), np.uint8); m[0, 0] = m[15, 15] = m[0, 15] = 1; masks.append(m.copy())
    m = np.zeros((16, 16), np.uint8); m[5, 5] = m[6, 6] = m[7, 7] = 1; masks.append(m.copy())
    m = np.zeros((16, 16), np.uint8); m[5, 5] = m[6, 6] = m[7, 7] = m[8, 8] = 1; masks.append(m.copy())
    m = np.ones((16, 16), np.uint8); m[4:7, 4:7] = 0; m[5, 5] = 1; masks.append(m.copy())
    m = np.zeros((16, 16), np.uint8); m[2, 2:4] = 1; m[4, 2] = 1; m[2, 6] = 1; masks.append(m.copy())
    m = np.zeros((16, 16), np.uint8); m[0, 3:5] = 1; m[7, 0] = 1; m[15, 8:11] = 1; masks.append(m.copy())
    for p in (0.05, 0.15, 0.3, 0.5, 0.8, 0.95):
        for _ in range(20):
            masks.append((rng.random((16, 16)) < p).astype(np.uint8))
    masks = np.stack(masks)
    refined = np.stack([gpl.refine_post_process(torch.from_numpy(mm).unsqueeze(0).float())[0].numpy().astype(np.uint8)
                        for mm in masks])
    res["refine_in"], res["refine_out"] = masks, refined
    np.savez_compressed(GOLD / "pseudo_label.npz", **res)


def gold_decoder():
    from safetensors.torch import load_file

    from engine.config.config import CfgNode as RefCfg
    from models.discriminator import Discriminator
    from models.uscod import baseline

    from oracle import decoder as odec
    res = {}
    for kind in ("dinov1", "dinov2"):
        sd = load_file(str(REF / "weights" / f"UCOD_DPL_{kind}.safetensors"))
        m = baseline(RefCfg({"dim": 768}))
        m.load_state_dict(sd, strict=True)
        m.eval()
        for (B, S, seed) in ((2, 68, 21), (1, 37, 22)):
            x = torch.randn(B, 768, S, S, generator=torch.Generator().manual_seed(seed))
            with torch.no_grad():
                fg, bg, ortho = m(x)
                ema = m(x, ema=True)
            t = f"{kind}_{S}"
            res[t + "_fg"], res[t + "_bg"] = fg.numpy(), bg.numpy()
            res[t + "_ortho"], res[t + "_ema"] = np.float32(ortho.item()), ema.numpy()
    # discriminator (random seeded weights through the reference module), BN in train mode as in APM
    dsd = odec.random_discriminator_state_dict(68, seed=31)
    D = Discriminator(RefCfg({"dis_use_features": False, "dim": 768, "feature_size": 68}))
    D.load_state_dict(dsd, strict=True)
    D.train()
    g = torch.Generator().manual_seed(32)
    masks = (torch.rand(8, 1, 68, 68, generator=g) < torch.rand(8, 1, 1, 1, generator=g)).float()
    with torch.no_grad():
        res["disc_train"] = D(masks, None).numpy()
    D.load_state_dict(dsd, strict=True)
    D.eval()
    with torch.no_grad():
        res["disc_eval"] = D(masks, None).numpy()
    # APM through the reference's TrainLoop.merge_pseudo_label (unbound, minimal fake `self`)
    from engine.runner.loop_UCOD_DPL import TrainLoop
    D.load_state_dict(dsd, strict=True)
    D.train()
    pl = torch.rand(8, 1, 68, 68, generator=g)
    teacher = torch.randn(8, 1, 68, 68, generator=g)
    student = torch.randn(8, 1, 68, 68, generator=g) + 0.3
    fake = SimpleNamespace(runner=SimpleNamespace(discriminator=D, logger=SimpleNamespace(log=lambda *a, **k: None)),
                           _cur_epoch=3, _max_epoch=25, _start_finetune=-5, dis_loss=nn.BCELoss())
    orig_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else orig_to(self, *a, **k)
    try:
        with torch.no_grad():
            merged, loss = TrainLoop.merge_pseudo_label(fake, pl, teacher, student, None)
    finally:
        torch.Tensor.to = orig_to
    res["apm_merged"], res["apm_loss"] = merged.numpy(), np.float32(loss.item())
    np.savez_compressed(GOLD / "decoder.npz", **res)


def blob_logits(n, fs=68, amp=4.0, seed=0, size=(0.03, 0.12)):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:fs, 0:fs]
    z = -amp * np.ones((fs, fs), np.float32)
    for _ in range(n):
        cy, cx = g.uniform(5, fs - 5, 2)
        r = g.uniform(*size) * fs
        ax, ay = r * g.uniform(0.6, 1.6), r * g.uniform(0.6, 1.6)
        z = np.maximum(z, amp * (1 - 2 * (((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2)))
    return torch.from_numpy(z.astype(np.float32))[None, None]


def gold_looktwice():
    import cv2
    from PIL import Image

    from engine.runner.loop_UCOD_DPL import ValLoop_Look_Twice
    res = {}
    meta = {}
    for S, th in ((518, 0.15), (296, 0.05)):
        fake = SimpleNamespace(cfg=SimpleNamespace(val_cfg=SimpleNamespace(look_twice_th=th, expand_type="dynamic")),
                               img_size=(S, S))
        fake.expand_bbox = lambda *a, **k: ValLoop_Look_Twice.expand_bbox(fake, *a, **k)
        rng = np.random.default_rng(S)
        logits, outs = [], []
        cases = [blob_logits(int(rng.integers(0, 6)), seed=t, size=(0.03, 0.2) if t % 2 else (0.02, 0.08))
                 for t in range(16)]
        cases += [blob_logits(0, seed=99), torch.full((1, 1, 68, 68), 4.0)]
        for lg in cases:
            try:
                up, bb = ValLoop_Look_Twice.process_preds(fake, lg, None)
                outs.append("None" if bb is None else bb)
                res[f"mask_{S}_{len(logits)}"] = np.packbits(up[0].numpy().astype(np.uint8))
            except ValueError:
                outs.append("ValueError")
            logits.append(lg[0, 0].numpy())
        res[f"logits_{S}"] = np.stack(logits)
        meta[str(S)] = outs
    meta["resize_bbox"] = [[b, ValLoop_Look_Twice.resize_bbox(None, b, 518, 518, W0, H0), [W0, H0]]
                           for b, (W0, H0) in (([10, 20, 100, 50], (1036, 777)), ([129, 129, 259, 259], (640, 480)),
                                               ([0, 0, 518, 518], (3000, 2000)), ([511, 3, 7, 500], (519, 517)))]
    # cv2 label order on a mask where pixel-raster and block-raster first-touch order differ
    m = np.zeros((12, 40), np.uint8)
    m[1:5, 10:14] = 255
    m[0:4, 20:24] = 255
    m[6:9, 2:5] = 255
    n, lab = cv2.connectedComponents(m, connectivity=8)
    res["cc_mask"], res["cc_labels"] = m, lab.astype(np.int32)
    rng = np.random.default_rng(7)
    big = (rng.random((64, 80)) < 0.45).astype(np.uint8) * 255
    n2, lab2, st2, _ = cv2.connectedComponentsWithStats(big, connectivity=8)
    res["cc_big_mask"], res["cc_big_labels"], res["cc_big_stats"] = big, lab2.astype(np.int32), st2.astype(np.int32)
    # Pillow resampling
    img = rng.integers(0, 256, (300, 260, 3), dtype=np.uint8)
    res["pil_src"] = img
    pim = Image.fromarray(img)
    res["pil_bilinear_518"] = np.asarray(pim.resize((518, 518), Image.BILINEAR))
    res["pil_bilinear_64x48"] = np.asarray(pim.resize((64, 48), Image.BILINEAR))
    crop = pim.crop((-7, 250, 120, 330))
    res["pil_crop_resize_296"] = np.asarray(crop.resize((296, 296), Image.BILINEAR))
    pred = ((rng.random((37, 37)) < 0.4).astype(np.uint8) * 255)
    res["pil_pred"] = pred
    for (w, h) in ((120, 77), (300, 41), (37, 37), (12, 9), (518, 518)):
        res[f"pil_bicubic_{w}x{h}"] = np.asarray(Image.fromarray(pred).resize((w, h)))  # Pillow default filter
    np.savez_compressed(GOLD / "looktwice.npz", **res)
    (GOLD / "looktwice_meta.json").write_text(json.dumps(meta))


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    install_shims()
    (GOLD / "VERSIONS.json").write_text(json.dumps(versions(), indent=1))
    only = set(sys.argv[1:])
    for name, fn in (("configs", gold_configs), ("vit", gold_vit), ("pseudo_label", gold_pseudo_label),
                     ("decoder", gold_decoder), ("looktwice", gold_looktwice)):
        if only and name not in only:
            continue
        fn()
        print("wrote", name, flush=True)
    try:
        from tools.make_golden_metrics import gold_metrics
        if not only or "metrics" in only:
            gold_metrics()
            print("wrote metrics", flush=True)
    except ImportError:
        pass
    try:
        from tools.make_golden_train import gold_discriminator_train, gold_train
        if not only or "train" in only:
            gold_train()
            gold_discriminator_train()
            print("wrote train", flush=True)
    except ImportError:
        pass
    try:
        from tools.make_golden_coral import gold_coral
        if not only or "coral" in only:
            gold_coral()
            print("wrote coral", flush=True)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
