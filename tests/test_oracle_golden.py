"""Pins the oracle: every CPU restatement under oracle/ is checked against golden vectors that
tools/make_golden.py produced by running the REFERENCE's own modules (imported from /root/reference) and the real
third-party libraries it calls (HF transformers, cv2, Pillow) — versions in tests/golden/VERSIONS.json.

CPU-only; nothing here touches the CUDA library.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import cc as occ
from oracle import decoder as odec
from oracle import looktwice as olt
from oracle import pil_resample as opr
from oracle import pseudo_label as opl
from oracle import vit as ovit
from ucod_dpl_b200.synth import random_vit_state_dict, synth_batch_u8

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def g_vit():
    return np.load(GOLD / "vit_hf.npz")


@pytest.fixture(scope="module")
def g_pl():
    return np.load(GOLD / "pseudo_label.npz")


@pytest.fixture(scope="module")
def g_dec():
    return np.load(GOLD / "decoder.npz")


@pytest.fixture(scope="module")
def g_lt():
    return np.load(GOLD / "looktwice.npz")


# ---- a1/a2: ViT key extraction + CLS attention row vs HF transformers --------------------------------------
@pytest.mark.parametrize("kind,S", [("dinov2", 224), ("dinov2", 518), ("dinov1", 296)])
def test_vit_matches_hf(g_vit, kind, S):
    spec = ovit.spec_for(kind)
    sd = random_vit_state_dict(spec, seed=0)
    x = ovit.normalize_u8(synth_batch_u8(0, 1, S, S))
    out = ovit.vit_forward(sd, spec, x, want_attn=True)
    k = out["key_tokens"][0].numpy()
    tag = f"{kind}_{S}"
    # fp32 vs fp32, different op order only (tolerance 2e-4 absolute on O(1) activations)
    np.testing.assert_allclose(k[::97], g_vit[tag + "_key_rows"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(k[:, ::61], g_vit[tag + "_key_cols"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(out["cls_attn"][0].numpy(), g_vit[tag + "_cls_attn"], atol=1e-6, rtol=1e-3)


# ---- a3: compute_img_bkg_seg ------------------------------------------------------------------------------
def _planted(B, P=256, nh=12, seed=0):
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


def test_pseudo_label_scoring(g_pl):
    att, keys = _planted(4, seed=11)
    for b in range(4):  # reference runs B = 1 (generate_pseudo_label.py:141-143)
        bkg, sim, _, _ = opl.compute_img_bkg_seg(att[b:b + 1], keys[b:b + 1], (16, 16), 0.6)
        assert np.array_equal(bkg[0].numpy(), g_pl["score_bkg"][b])
        np.testing.assert_allclose(sim[0].numpy(), g_pl["score_sim"][b], atol=1e-6)
    bkg, sim, _, _ = opl.compute_img_bkg_seg(att, keys, (16, 16), 0.6)  # batch-global sim max (:81-85)
    assert np.array_equal(bkg.numpy(), g_pl["score_bkg_batched"])
    np.testing.assert_allclose(sim.numpy(), g_pl["score_sim_batched"], atol=1e-6)
    assert 0 < g_pl["score_bkg"].mean() < 1  # the planted inputs exercise both labels


def test_pseudo_label_scoring_optional_arguments():
    """`up_size` != grid and `apply_weights=False` (found_bkg_mask.py:9-12) against outputs of the reference itself."""
    from tools.make_golden_bkgseg_options import CASES, TH_BKG, planted_inputs
    gold = np.load(GOLD / "bkgseg_options.npz")
    att, feats = planted_inputs()
    for name, kw in CASES:
        bkg, sim, _, _ = opl.compute_img_bkg_seg(att, feats, (16, 16), TH_BKG, **kw)
        assert bkg.shape == gold[f"{name}_bkg"].shape
        assert np.array_equal(bkg.numpy().astype(np.uint8), gold[f"{name}_bkg"]), name
        np.testing.assert_allclose(sim.numpy(), gold[f"{name}_sim"], atol=2e-6)
        assert 0 < gold[f"{name}_bkg"].mean() < 1


# ---- a4: refine_post_process (cv2.connectedComponentsWithStats semantics) -----------------------------------
def test_refine_post_process(g_pl):
    changed = 0
    for m, want in zip(g_pl["refine_in"], g_pl["refine_out"]):
        got = opl.refine_post_process(m)
        assert np.array_equal(got, want)
        changed += int((want != m).any())
    assert changed > 10  # the fixture really contains components that get flipped


# ---- a6-a8: decoder with the shipped checkpoints (the known-answer test) ----------------------------------
@pytest.mark.parametrize("kind", ["dinov1", "dinov2"])
def test_decoder_shipped_weights(g_dec, kind):
    from safetensors.torch import load_file
    sd = load_file(str(ROOT / "weights" / f"UCOD_DPL_{kind}.safetensors"))
    for (B, S, seed) in ((2, 68, 21), (1, 37, 22)):
        x = torch.randn(B, 768, S, S, generator=torch.Generator().manual_seed(seed))
        fg, bg, ortho = odec.baseline_forward(sd, x)
        ema = odec.baseline_forward(sd, x, ema=True)
        t = f"{kind}_{S}"
        np.testing.assert_allclose(fg.numpy(), g_dec[t + "_fg"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(bg.numpy(), g_dec[t + "_bg"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(ema.numpy(), g_dec[t + "_ema"], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(float(ortho), float(g_dec[t + "_ortho"]), rtol=1e-3)


# ---- a9/a10: discriminator + APM --------------------------------------------------------------------------
def test_discriminator_and_apm(g_dec):
    dsd = odec.random_discriminator_state_dict(68, seed=31)
    g = torch.Generator().manual_seed(32)
    masks = (torch.rand(8, 1, 68, 68, generator=g) < torch.rand(8, 1, 1, 1, generator=g)).float()
    np.testing.assert_allclose(odec.discriminator_forward(dsd, masks, True).numpy(), g_dec["disc_train"], atol=1e-5)
    np.testing.assert_allclose(odec.discriminator_forward(dsd, masks, False).numpy(), g_dec["disc_eval"], atol=1e-5)
    pl = torch.rand(8, 1, 68, 68, generator=g)
    teacher = torch.randn(8, 1, 68, 68, generator=g)
    student = torch.randn(8, 1, 68, 68, generator=g) + 0.3
    merged, loss, w, _, _ = odec.apm_merge(dsd, pl, teacher, student, cur_epoch=3)
    np.testing.assert_allclose(merged.numpy(), g_dec["apm_merged"], atol=1e-5)
    np.testing.assert_allclose(float(loss), float(g_dec["apm_loss"]), rtol=1e-5)
    assert 0.0 < float(w.min()) and float(w.max()) <= 1.0


# ---- a12/a13: process_preds / expand_bbox / resize_bbox ---------------------------------------------------
@pytest.mark.parametrize("S,th", [(518, 0.15), (296, 0.05)])
def test_process_preds_boxes(g_lt, S, th):
    meta = json.loads((GOLD / "looktwice_meta.json").read_text())[str(S)]
    logits = g_lt[f"logits_{S}"]
    kinds = set()
    for i, want in enumerate(meta):
        lg = torch.from_numpy(logits[i])[None, None]
        if want == "ValueError":
            with pytest.raises(ValueError):
                olt.process_preds(lg, (S, S), th, "dynamic")
            kinds.add("err")
            continue
        up, bb = olt.process_preds(lg, (S, S), th, "dynamic")
        gold_mask = np.unpackbits(g_lt[f"mask_{S}_{i}"])[: S * S].reshape(S, S)
        assert np.array_equal(up[0].numpy().astype(np.uint8), gold_mask)
        if want == "None":
            assert bb is None
            kinds.add("none")
        else:
            assert bb == want, (i, bb, want)
            kinds.add("default" if want == [olt.DEFAULT_BOX] else ("boxes" if want else "empty"))
    assert {"none", "default", "boxes"} <= kinds


def test_sigmoid_half_threshold_is_pinned():
    """The CUDA binarisation compares the interpolated logit with 1.5 * 2^-24 instead of evaluating
    `sigmoid(x) > 0.5` (csrc/decoder.cu UCOD_SIGMOID_HALF_THRESHOLD, loop_UCOD_DPL.py:356-361): in fp32 the two
    predicates are the same function of x (torch CPU, dense scan around the threshold and over the normal range)."""
    thr = float(np.float32(1.5 * 2.0 ** -24))
    around = torch.linspace(-4 * 2.0 ** -24, 4 * 2.0 ** -24, 400001, dtype=torch.float64).float().unique()
    wide = torch.cat([torch.linspace(-30, 30, 200001), torch.tensor([thr, float(np.nextafter(np.float32(thr), np.float32(1))),
                                                                    float(np.nextafter(np.float32(thr), np.float32(0)))])])
    for x in (around, wide):
        assert torch.equal(torch.sigmoid(x) > 0.5, x > thr)


def test_resize_bbox():
    meta = json.loads((GOLD / "looktwice_meta.json").read_text())["resize_bbox"]
    for box, want, (W0, H0) in meta:
        assert olt.resize_bbox(box, 518, 518, W0, H0) == want


# ---- cv2 connected components: label ORDER, labels and stats ----------------------------------------------
def test_connected_components_match_cv2(g_lt):
    n, lab = occ.connected_components_8(g_lt["cc_mask"])
    assert np.array_equal(lab, g_lt["cc_labels"])  # 2-row-block raster numbering, not pixel raster
    n, lab = occ.connected_components_8(g_lt["cc_big_mask"])
    assert np.array_equal(lab, g_lt["cc_big_labels"])
    st = occ.stats(lab, n)
    assert np.array_equal(st[:, :5], g_lt["cc_big_stats"][:, :5])


# ---- Pillow resampling: antialiased bilinear (crop path) and default bicubic (paste path), bit-exact -------
def test_pil_bilinear_bit_exact(g_lt):
    src = g_lt["pil_src"]
    assert np.array_equal(opr.resize_u8(src, 518, 518, "bilinear"), g_lt["pil_bilinear_518"])
    assert np.array_equal(opr.resize_u8(src, 64, 48, "bilinear"), g_lt["pil_bilinear_64x48"])
    crop = opr.crop_u8(src, -7, 250, 120, 330)  # out-of-image area reads as 0 like PIL
    assert np.array_equal(opr.resize_u8(crop, 296, 296, "bilinear"), g_lt["pil_crop_resize_296"])


@pytest.mark.parametrize("wh", [(120, 77), (300, 41), (37, 37), (12, 9), (518, 518)])
def test_pil_bicubic_bit_exact(g_lt, wh):
    w, h = wh
    assert np.array_equal(opr.resize_u8(g_lt["pil_pred"], w, h, "bicubic"), g_lt[f"pil_bicubic_{w}x{h}"])
