"""Preprocessing transforms, dataset feature cache and the eval launcher on the GPU.

* `ImageTransforms` vs torchvision's own `Resize -> ToTensor -> Normalize` on PIL images: bit-exact fp32.
* features cache written by `USCODDataset` is the reference layout and equals backbone(transform_image(img)).
* `scripts.eval.main` on a ragged synthetic image-folder: PNGs vs the fp32 CPU oracle of the whole Look-Twice
  flow, and the printed table vs the oracle metric suite applied to the written PNGs.
"""
import os
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from PIL import Image
from safetensors.torch import load_file

from oracle import decoder as odec
from oracle import looktwice as olt
from oracle import metrics as omet
from oracle import pil_resample as opr
from oracle import vit as ovit
from ucod_dpl_b200.data.datasets import ImageTransforms, USCODDataset, pack_padded
from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.vit import VitKeyExtractor
from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8
from ucod_dpl_b200.vit import spec_for

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
SIZES = [(300, 400), (448, 448), (260, 333), (518, 518), (97, 211)]


def _img(seed, h, w):
    return np.ascontiguousarray(synth_image_u8(seed, h, w).permute(1, 2, 0).numpy())


def test_transforms_match_torchvision_bit_exact():
    from torchvision import transforms as T
    norm = T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    imgs = [_img(i, h, w) for i, (h, w) in enumerate(SIZES)]
    for size in [(518, 518), (296, 296), (756, 756)]:
        ref_tf = T.Compose([T.Resize(size), T.ToTensor(), norm])
        mine = ImageTransforms.get_image_transform(size)
        got = mine.batch(imgs).cpu()
        for i, im in enumerate(imgs):
            want = ref_tf(Image.fromarray(im))
            assert torch.equal(got[i], want), (size, i, (got[i] - want).abs().max().item())
        assert torch.equal(mine(Image.fromarray(imgs[2])).cpu(), got[2])       # PIL input, single image
    # feature-extractor and patch transforms
    fe = ImageTransforms.get_feature_extractor_transform((432, 432))
    assert torch.equal(fe(imgs[0]).cpu(), T.Compose([T.Resize((432, 432)), T.ToTensor(), norm])(Image.fromarray(imgs[0])))
    assert torch.equal(ImageTransforms.get_patch_transform()(imgs[4]).cpu(),
                       T.Compose([T.ToTensor(), norm])(Image.fromarray(imgs[4])))
    # labels ('L'): resized and keep_size
    lab = (np.indices((260, 333)).sum(0) % 256).astype(np.uint8)
    want = T.Compose([T.Resize((518, 518)), T.ToTensor()])(Image.fromarray(lab, mode="L"))
    assert torch.equal(ImageTransforms.get_label_transform((518, 518))(lab).cpu(), want)
    assert torch.equal(ImageTransforms.get_label_transform((518, 518), keep_size=True)(lab).cpu(),
                       T.ToTensor()(Image.fromarray(lab, mode="L")))
    # raw transform = the uint8 image torchvision's Resize produces
    raw = ImageTransforms.get_raw_transform((296, 296))(imgs[1]).cpu()
    assert torch.equal(raw, torch.from_numpy(np.array(T.Resize((296, 296))(Image.fromarray(imgs[1])))).permute(2, 0, 1))


def _write_set(root, name, n, seed0):
    os.makedirs(root / name / "im"), os.makedirs(root / name / "gt")
    for i in range(n):
        # two-colour image (ellipse on a flat background, +-25 noise): through the random-init ViT this spreads the
        # first-look outcomes over "no second look", the default box, an empty and a 10-entry box list
        h, w = SIZES[(seed0 + i) % len(SIZES)]
        rng = np.random.default_rng(seed0 + i)
        yy, xx = np.mgrid[0:h, 0:w]
        blob = ((yy - h * 0.45) ** 2 / (0.08 * h * h) + (xx - w * 0.5) ** 2 / (0.05 * w * w)) < 1
        fgc, bgc = rng.integers(0, 255, 3), rng.integers(0, 255, 3)
        img = np.where(blob[..., None], fgc, bgc) + rng.integers(-25, 25, (h, w, 3))
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(root / name / "im" / f"img_{i:02d}.png")
        gt = blob.astype(np.uint8) * 255
        Image.fromarray(gt, mode="L").save(root / name / "gt" / f"img_{i:02d}.png")


def test_feature_cache_is_reference_layout(tmp_path):
    _write_set(tmp_path / "data", "TR-X", 3, 0)
    fe_cfg = SimpleNamespace(type="dinov2", backbone="facebook/dinov2-base", backbone_type="huggingface")
    cfg = SimpleNamespace(DATASET="TR-X", image_size=(224, 224), require_label=False)
    ds = USCODDataset(cfg, fe_cfg, "train", str(tmp_path / "data"), str(tmp_path / "cache"), extract_batch=2)
    base = tmp_path / "cache" / "features_cache" / "dinov2" / "train" / "TR-X"
    assert sorted(os.listdir(base)) == ["data_0.pkl", "data_1.pkl", "data_2.pkl", "index.json"]
    item = ds[2]
    assert item["features"].shape == (768, 16, 16) and item["features"].device.type == "cpu"
    img = ds.transform_image(Image.open(ds.image_paths[2]).convert("RGB"))
    _, key = ds.feature_extractor(img[None])
    # fp32-normalised input vs the fused uint8 path: same bf16 tokens up to the rounding of the folded constants
    assert (key[0].cpu() - item["features"]).abs().max().item() < 5e-2
    assert (key[0].cpu() - item["features"]).abs().mean().item() < 2e-3


def _oracle_stages(vit_sd, dec_sd, S, fs):
    """fp32 CPU pieces of loop_UCOD_DPL.py:297-317 (one image at a time), exposed stage by stage."""
    from torchvision import transforms as T
    spec = ovit.spec_for("dinov2")
    tf = T.Compose([T.Resize((S, S)), T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])

    def seg_keys(x):
        return ovit.keys_to_map(ovit.vit_forward(vit_sd, spec, x)["key_tokens"])

    def first_logits(img_u8):
        feats = F.interpolate(seg_keys(tf(Image.fromarray(img_u8))[None]), size=(fs, fs), mode="bilinear")
        return odec.baseline_forward(dec_sd, feats, want_ortho=False)[0]

    def second_logits(x):  # normalised crop [1,3,S,S] -> logits on the raw grid
        return odec.baseline_forward(dec_sd, seg_keys(x), want_ortho=False)[0]

    return first_logits, second_logits


def _margin_ok(got_mask, want_mask, ref_logit_up, delta):
    """mask pixels may differ from the oracle only where the oracle's own (interpolated) logit is a near-tie"""
    bad = got_mask != want_mask
    return (not bad.any()) or float(np.abs(ref_logit_up[bad]).max()) < delta, float(1.0 - bad.mean())


def test_eval_launcher_end_to_end(tmp_path, monkeypatch):
    """`scripts.eval` on ragged synthetic image folders (both test sets of the sweep), checked as a CHAIN against the
    fp32 CPU oracle, every link with identical inputs on both sides:
      (1) the PNG the launcher wrote == the evaluator called directly on the same decoded image (>= 99.95 %: the
          launcher batches four images, which changes the GEMM tile order and so the last bits of the logits), and the
          result table == the oracle metric suite over the PNGs (1e-9);
      (2) first look, float: |sigmoid diff| <= 1e-2 on the 68x68 logits, mask pixels differ only at oracle near-ties;
      (3) boxes, integer: the CUDA box list == the oracle's component / box logic run on the CUDA first-look mask;
      (4) second look, float: per box, logits on the raw grid within 1e-2 (sigmoid) of the oracle's on the PIL-made
          crop, cells differ only at oracle near-ties; paste, integer: Pillow-exact bicubic pastes of the CUDA-binarised
          second looks reproduce the evaluator's canvas bit for bit; the final bilinear resize + threshold agrees to
          >= 99.99 % (fp32 ties at exactly 0.5).
    No link injects oracle results into the CUDA path; floats are compared with north_star's tolerances and every
    integer stage bit for bit."""
    from ucod_dpl_b200 import ops
    from ucod_dpl_b200.scripts import eval as ev
    S, fs = 224, 68
    data = tmp_path / "data"
    _write_set(data, "SETA", 5, 0)
    _write_set(data, "SETB", 3, 11)
    cfg_text = (ROOT / "configs" / "uscod" / "UCOD-DPL_dinov2.py").read_text().replace("(518, 518)", f"({S}, {S})")
    os.makedirs(tmp_path / "configs" / "uscod"), os.makedirs(tmp_path / "configs" / "__base__")
    (tmp_path / "configs" / "uscod" / "tiny.py").write_text(cfg_text)
    (tmp_path / "configs" / "__base__" / "shared_defaults.py").write_text(
        (ROOT / "configs" / "__base__" / "shared_defaults.py").read_text())
    monkeypatch.chdir(tmp_path)
    ckpt = str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")
    res = ev.main(["--config", "configs/uscod/tiny.py", "--work_dir", "work", "--load_from", ckpt, "--dataset_dir",
                   str(data), "--datasets", "SETA,SETB", "--batch_size", "4", "--exp_name", "t0"])
    run = tmp_path / "work" / "uscod" / "tiny" / "t0"
    assert (run / "config.yaml").exists() and (run / "eval0.log").exists()
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = load_file(ckpt)
    model = baseline(SimpleNamespace(dim=768)).cuda().eval()
    model.load_state_dict(dec_sd)
    looker = LookTwiceEvaluator(VitKeyExtractor(vit_sd, spec_for("dinov2")), model, (S, S), fs, 0.15, "dynamic")
    tf = ImageTransforms.get_image_transform((S, S))
    first_logits, second_logits = _oracle_stages(vit_sd, dec_sd, S, fs)
    n_boxed = n_second = 0
    for name, n in (("SETA", 5), ("SETB", 3)):
        files = sorted(os.listdir(run / "preds" / name))
        assert files == [f"img_{i:02d}.png" for i in range(n)]
        items = []
        for f in files:
            pred = np.asarray(Image.open(run / "preds" / name / f))
            gt = np.asarray(Image.open(data / name / "gt" / f))
            assert pred.shape == gt.shape and set(np.unique(pred)) <= {0, 255}
            items.append(omet.per_image(gt.astype(np.float64) / 255.0, pred > 0))
            img = np.asarray(Image.open(data / name / "im" / f).convert("RGB"))
            # (1) launcher == evaluator on the same image, exactly
            canvas, sizes = pack_padded([img], "cuda")
            r = looker.look_twice_device(tf.batch([img]), canvas, layout="HWC", orig_sizes=sizes)
            r.check()
            direct = ops.upsample_bilinear(r.final[0], gt.shape, binarize=3).cpu().numpy()
            assert float(((direct > 0) == (pred > 0)).mean()) >= 0.9995, (name, f)
            # (2) first look (float)
            lg_ref = first_logits(img)
            lg = looker.first_look(tf.batch([img])).cpu()
            assert (torch.sigmoid(lg) - torch.sigmoid(lg_ref)).abs().max().item() <= 1e-2
            up_ref = F.interpolate(lg_ref, size=(S, S), mode="bilinear", align_corners=False)[0, 0].numpy()
            ok, agree1 = _margin_ok(r.first[0].cpu().numpy(), (up_ref > 0).astype(np.uint8), up_ref, 0.05)
            assert ok, (name, f, agree1)   # (an image whose logits hover around 0 has many such ties: 98.4 % on img_03)
            # (3) boxes (integer) from the CUDA mask
            boxes = olt.boxes_from_mask(r.first[0].cpu().numpy(), (S, S), 0.15, "dynamic")
            assert boxes == r.bboxes[0], (name, f, boxes, r.bboxes[0])
            # (4a) second look, float: per box, the CUDA logits on the raw grid vs the oracle's on the PIL-made crop
            first01 = r.first[0].cpu().numpy()
            canvas_u8 = (first01 * 255).astype(np.uint8)
            if boxes:
                n_boxed += boxes != [olt.DEFAULT_BOX]
                n_second += len(boxes)
                H0, W0 = img.shape[:2]
                cj = torch.tensor([[0] + olt.resize_bbox(bb, S, S, W0, H0) for bb in boxes], dtype=torch.int32).cuda()
                crops = ops.roi_crop_resize(canvas, cj, (S, S), layout="HWC")
                _, k16, _ = looker.extractor.keys(crops, want_f32=False, want_bf16=True)
                g = S // 14
                fg2 = model.decoder.forward_tokens(k16, (g, g), (g, g), want_bg=False)[0].cpu()
                for i, bb in enumerate(boxes):
                    x, _ = olt.crop_resize_normalize(img, cj[i, 1:].tolist(), (S, S))
                    ref2 = second_logits(x[None])
                    assert (torch.sigmoid(fg2[i]) - torch.sigmoid(ref2[0])).abs().max().item() <= 1e-2
                    bad = (fg2[i] > 0) != (ref2[0] > 0)
                    assert (not bad.any()) or ref2[0][bad].abs().max().item() < 0.05
                    # (4b) paste, integer: Pillow-exact bicubic of the CUDA-binarised second look
                    pred_u8 = ((fg2[i, 0] > 0).to(torch.uint8) * 255).numpy()
                    opr.paste_u8(canvas_u8, opr.resize_u8(pred_u8, bb[2], bb[3], "bicubic"), bb[0], bb[1])
            assert np.array_equal(canvas_u8, np.rint(r.final[0].cpu().numpy() * 255).astype(np.uint8)), (name, f)
            up = torch.from_numpy(canvas_u8).float().div(255.0)
            want = (F.interpolate(up.reshape(1, 1, S, S), size=gt.shape, mode="bilinear")[0, 0] > 0.5).numpy()
            agree = float(((direct > 0) == want).mean())
            print(name, f, "boxes", boxes, "first-look agreement", round(agree1, 5), "final agreement", round(agree, 5))
            assert agree >= 0.9999, (name, f, agree)
        # the launcher's table is the oracle metric suite over the PNGs it wrote
        want_tab = omet.aggregate(items)
        for k, v in want_tab.items():
            assert abs(res[name][k] - v) < 1e-9, (name, k, res[name][k], v)
    assert n_boxed >= 1 and n_second >= 3, "the synthetic sets should exercise real box lists"


def test_second_stage_launcher(tmp_path, monkeypatch):
    """`scripts.LTeval` on a small image folder: masks equal `CoralEvaluator` called directly on the decoded images
    (its parity with the CPU oracle is test_coral_gpu's job), grouped launches for equal-size originals included."""
    from safetensors.torch import save_file
    from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator
    from ucod_dpl_b200.models.UDLR import SparseRefiner
    from ucod_dpl_b200.scripts import LTeval
    from ucod_dpl_b200.synth import random_refiner_state_dict
    S = 224
    data = tmp_path / "data"
    os.makedirs(data / "SETC" / "im"), os.makedirs(data / "SETC" / "gt")
    shapes = [(300, 340), (260, 333), (300, 340)]
    for i, (h, w) in enumerate(shapes):
        Image.fromarray(_img(20 + i, h, w)).save(data / "SETC" / "im" / f"c{i}.png")
        Image.fromarray(((np.indices((h, w)).sum(0) % 64) < 20).astype(np.uint8) * 255, mode="L").save(
            data / "SETC" / "gt" / f"c{i}.png")
    cfg_text = (ROOT / "configs" / "uscod" / "CORAL_dinov2.py").read_text().replace("(518, 518)", f"({S}, {S})")
    os.makedirs(tmp_path / "configs" / "uscod"), os.makedirs(tmp_path / "configs" / "__base__")
    (tmp_path / "configs" / "uscod" / "tiny_coral.py").write_text(cfg_text)
    (tmp_path / "configs" / "__base__" / "shared_defaults.py").write_text(
        (ROOT / "configs" / "__base__" / "shared_defaults.py").read_text())
    ref_sd = random_refiner_state_dict(0)
    save_file({k: v.contiguous() for k, v in ref_sd.items()}, str(tmp_path / "refiner.safetensors"))
    monkeypatch.chdir(tmp_path)
    ckpt = str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")
    res = LTeval.main(["--config", "configs/uscod/tiny_coral.py", "--work_dir", "work", "--load_from", ckpt,
                       "--refiner_path", str(tmp_path / "refiner.safetensors"), "--dataset_dir", str(data),
                       "--datasets", "SETC", "--exp_name", "c0"])
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    model = baseline(SimpleNamespace(dim=768)).cuda().eval()
    model.load_state_dict(load_file(ckpt))
    refiner = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015)).cuda().eval()
    refiner.load_state_dict(ref_sd, strict=True)
    ev = CoralEvaluator(VitKeyExtractor(vit_sd, spec_for("dinov2")), model, refiner, (S, S), 3, 56)
    items = []
    for i, (h, w) in enumerate(shapes):
        pred = np.asarray(Image.open(tmp_path / "work" / "uscod" / "tiny_coral" / "c0" / "preds" / "SETC" / f"c{i}.png"))
        img = torch.from_numpy(np.array(Image.open(data / "SETC" / "im" / f"c{i}.png").convert("RGB")))
        want = ev(img[None].cuda(), label_sizes=[(h, w)], layout="HWC")[0].cpu().numpy()
        # same kernels, different batch composition (fp32 accumulation order): boundary pixels may flip
        assert ((pred > 0) == (want > 0)).mean() >= 0.999
        gt = np.asarray(Image.open(data / "SETC" / "gt" / f"c{i}.png"))
        items.append(omet.per_image(gt.astype(np.float64) / 255.0, pred > 0))
    for k, v in omet.aggregate(items).items():
        assert abs(res["SETC"][k] - v) < 1e-9, (k, res["SETC"][k], v)


def test_train_launcher_end_to_end(tmp_path, monkeypatch):
    """`scripts.train` on a tiny image folder: builds both caches in the reference layout, runs the schedule
    (discriminator epoch, decoder epochs, finetune switch, checkpoint, Look-Twice validation), and the checkpoint it
    writes is found and evaluated by `scripts.eval` without --load_from."""
    from ucod_dpl_b200.engine.utils.fileio import MetaListPickleIO
    from ucod_dpl_b200.scripts import eval as ev
    from ucod_dpl_b200.scripts import train as tr
    S = 224
    data = tmp_path / "data"
    _write_set(data, "TR-X", 12, 0)
    _write_set(data, "TE-X", 3, 30)
    os.makedirs(tmp_path / "configs" / "uscod"), os.makedirs(tmp_path / "configs" / "__base__")
    (tmp_path / "configs" / "uscod" / "tiny.py").write_text(
        (ROOT / "configs" / "uscod" / "UCOD-DPL_dinov2.py").read_text().replace("(518, 518)", f"({S}, {S})"))
    (tmp_path / "configs" / "__base__" / "shared_defaults.py").write_text(
        (ROOT / "configs" / "__base__" / "shared_defaults.py").read_text()
        .replace("TR-CAMO+TR-COD10K", "TR-X").replace("TE-CAMO", "TE-X"))
    monkeypatch.chdir(tmp_path)
    out = tr.main(["--config", "configs/uscod/tiny.py", "--work_dir", "work", "--dataset_dir", str(data),
                   "--cache_dir", str(tmp_path / "cache"), "--max_epoch", "6", "--exp_name", "run0", "--no_save"])
    # caches: reference layout, one item per image, reference item shapes
    feats = MetaListPickleIO(base_path=tmp_path / "cache" / "features_cache" / "dinov2" / "train" / "TR-X")
    pls = MetaListPickleIO(base_path=tmp_path / "cache" / "pseudo_label_cache" / "TR-X")
    assert feats.mode == "r" and pls.mode == "r" and feats.len() == 12 and pls.len() == 12
    assert tuple(feats.read_file(3).shape) == (768, 16, 16) and tuple(pls.read_file(3).shape) == (1, 16, 16)
    assert set(pls.read_file(3).unique().tolist()) <= {0.0, 1.0}
    # schedule: epoch 0 logs its losses; checkpoint and validation after epoch 5
    assert len(out["losses"]) == 1 and np.isfinite(out["losses"][0])
    run = tmp_path / "work" / "uscod" / "tiny" / "run0"
    ck = run / "ckp" / "epoch5.pth" / "model.safetensors"
    assert ck.exists() and sorted(os.listdir(run / "ckp")) == ["epoch5.pth"]
    sd = load_file(str(ck))
    ref = load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
    assert set(sd.keys()) == set(ref.keys()) and all(sd[k].shape == ref[k].shape for k in ref)
    assert all(torch.isfinite(v).all() for v in sd.values())
    assert not torch.equal(sd["decoder.decoupling.weight"], sd["decoder_ema.decoupling.weight"])
    assert out["best"] is not None and 0.0 <= out["best"]["MAE"] <= 1.0
    # the eval launcher discovers that checkpoint (no --load_from) and reproduces the validation result
    res = ev.main(["--config", "configs/uscod/tiny.py", "--work_dir", "work", "--dataset_dir", str(data),
                   "--datasets", "TE-X", "--exp_name", "eval0", "--no_save"])
    for k, v in out["best"].items():
        assert abs(res["TE-X"][k] - v) < 1e-9, (k, res["TE-X"][k], v)


def test_lr_dataset_patch_caches(tmp_path):
    """`LRDataset` (second-stage dataset): window / m-patch caches in the reference layout and shapes; a window's keys
    equal the backbone run on that window of the 3x-resized image."""
    from ucod_dpl_b200.data.datasets import LRDataset
    from ucod_dpl_b200.engine.utils.fileio import MetaListPickleIO
    S = 224
    data = tmp_path / "data"
    os.makedirs(data / "TR-L" / "im")
    shapes = [(300, 340), (260, 333), (300, 340)]
    for i, (h, w) in enumerate(shapes):
        Image.fromarray(_img(50 + i, h, w)).save(data / "TR-L" / "im" / f"l{i}.png")
    fe_cfg = SimpleNamespace(type="dinov2", backbone="facebook/dinov2-base", backbone_type="huggingface")
    cfg = SimpleNamespace(DATASET="TR-L", image_size=(S, S), require_label=False, require_m_patches=False, use_cache=True)
    ds = LRDataset(cfg, fe_cfg, "train", str(data), str(tmp_path / "cache"), window_size=3)
    patch = MetaListPickleIO(base_path=tmp_path / "cache" / "patch_cache" / "dinov2" / "train" / "TR-L")
    mpatch = MetaListPickleIO(base_path=tmp_path / "cache" / "m_patch_cache" / "dinov2" / "train" / "TR-L")
    assert patch.mode == "r" and patch.len() == 3 and mpatch.mode == "r" and mpatch.len() == 3   # train mode: m-patches
    item = ds[1]
    assert list(item.keys()) == ["pseudo_label", "label_tensor", "features", "img_path", "m_inputs", "h_inputs", "index"]
    assert tuple(item["h_inputs"].shape) == (9, 768, 16, 16) and tuple(item["m_inputs"].shape) == (4, 768, 36, 36)
    assert tuple(item["features"].shape) == (768, 16, 16) and item["index"] == [1]
    # window (row 1, col 2) of image 1: resize to 3S x 3S (Pillow-exact), cut, backbone
    img = np.asarray(Image.open(ds.image_paths[1]).convert("RGB"))
    big = ImageTransforms.get_raw_transform((3 * S, 3 * S))(img)
    win = big[:, S:2 * S, 2 * S:3 * S].contiguous()
    _, key = ds.feature_extractor(win[None])
    d = (key[0].cpu() - item["h_inputs"][1 * 3 + 2]).abs()
    assert d.max().item() < 5e-2 and d.mean().item() < 2e-3, (d.max().item(), d.mean().item())
    # reference-compatible per-image entry
    patches, m = ds.get_features(str(ds.image_paths[1]))
    assert len(patches) == 9 and tuple(patches[0].shape) == (768, 16, 16) and tuple(m.shape) == (1, 4, 768, 36, 36)
    key_c, h_c, _ = ds.get_features(str(ds.image_paths[1]), crop_center=True)
    assert tuple(key_c.shape) == (1, 768, 16, 16) and tuple(h_c.shape) == (1, 9, 768, 16, 16)


def test_pseudo_label_command_line(tmp_path):
    """`python -m ucod_dpl_b200.generate_pseudo_label` arguments of the reference's `main()`: cache in the reference
    layout, one `[1,16,16]` {0,1} CPU float tensor per image in sorted path order, equal to the batched generator."""
    from ucod_dpl_b200 import generate_pseudo_label as gpl
    from ucod_dpl_b200.engine.utils.fileio import MetaListPickleIO
    data = tmp_path / "RefCOD"
    _write_set(data, "TR-A", 5, 0)
    _write_set(data, "TR-B", 3, 7)
    n = gpl.main(["--dataset", "TR-A+TR-B", "--image_path", str(data / "{}" / "im"), "--cache_path",
                  str(tmp_path / "plc"), "--backbone_weights", str(tmp_path / "none")])
    assert n == 8
    io = MetaListPickleIO(base_path=tmp_path / "plc" / "TR-A+TR-B")
    assert io.mode == "r" and io.len() == 8
    paths = sorted([str(p) for d in ("TR-A", "TR-B") for p in (data / d / "im").iterdir()])
    gen = gpl.PseudoLabelGenerator(random_vit_state_dict(spec_for("dinov2"), seed=0), "dinov2")
    tf = ImageTransforms.get_raw_transform((224, 224))
    want = gen(tf.batch([np.asarray(Image.open(p).convert("RGB")) for p in paths])).cpu()
    for i in range(8):
        item = io.read_file(i)
        assert tuple(item.shape) == (1, 16, 16) and item.dtype == torch.float32 and item.device.type == "cpu"
        assert torch.equal(item[0], want[i].float())
