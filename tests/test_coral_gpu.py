"""CORAL SparseRefiner (CUDA) vs. the CPU oracle and the reference-generated golden vectors."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import coral as oc
from ucod_dpl_b200 import ops
from ucod_dpl_b200.models.UDLR import SparseRefiner
from ucod_dpl_b200.synth import random_refiner_state_dict, synth_coral_inputs

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parents[1] / "tests" / "golden"


@pytest.fixture(scope="module")
def refiner():
    r = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015))
    r.load_state_dict(random_refiner_state_dict(0), strict=True)
    return r.cuda().eval()


@pytest.mark.parametrize("P", [56, 102])
def test_entropy_select_matches_oracle(P):
    g = torch.Generator().manual_seed(P)
    preds = torch.randn(3, 1, P, P, generator=g) * 6
    preds[1] = 14.0
    for x in (preds, preds.sigmoid()):   # logits and probabilities take different branches (ASR.py:42-45)
        ent, scores, mask = ops.coral_entropy_select(x.cuda(), 0.0015, 3)
        _, _, rmask, _, rent, rscores = oc.entropy_select(torch.zeros(3, 1, 1, 1), torch.zeros(3, 9, 1, 1, 1), x,
                                                          0.0015, 3)
        np.testing.assert_allclose(ent.cpu().numpy(), rent.numpy(), atol=2e-6)
        np.testing.assert_allclose(scores.cpu().numpy(), rscores.numpy(), rtol=1e-4, atol=1e-7)
        far = (rscores - 0.0015).abs() > 1e-5     # selection is exact wherever the score is not within rounding of the threshold
        assert torch.equal(mask.cpu()[far], rmask[far])


def test_gated_ensemble_and_scatter_match_oracle(refiner):
    sd = random_refiner_state_dict(0)
    g = torch.Generator().manual_seed(3)
    preds = torch.randn(2, 1, 56, 56, generator=g) * 3
    wins = torch.randn(3, 1, 56, 56, generator=g)
    mask = torch.zeros(2, 1, 3, 3, dtype=torch.bool)
    mask[0, 0, 0, 2] = mask[0, 0, 2, 1] = mask[1, 0, 1, 1] = True
    coords = torch.tensor([[0, 2], [2, 1], [1, 1]])
    ref_h = oc.concate_windows(wins, coords, mask, 3)
    h = refiner.HRE.concate_windows(wins.cuda(), coords.cuda(), mask.cuda())
    np.testing.assert_allclose(h.cpu().numpy(), ref_h.numpy(), atol=1e-6)
    assert (h.cpu()[1, 0, :56] == 0).all()                      # unselected cells are exactly zero
    ref_out, ref_w = oc.gated_ensembler(sd, preds, ref_h)
    out, w = refiner.GE(preds.cuda(), h)
    np.testing.assert_allclose(w.cpu().numpy(), ref_w.numpy(), atol=2e-5)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out.numpy(), atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("tag,seed,batch,unc", [("a", 5, 1, ((0, 1), (1, 1), (2, 0))), ("b", 6, 2, ((1, 2),))])
def test_sparse_refiner_matches_reference_golden(refiner, tag, seed, batch, unc):
    gold = np.load(GOLD / "coral.npz")
    l, h, preds = synth_coral_inputs(seed, batch=batch, uncertain=unc)
    if tag == "b":
        preds[1] = -12.0
    out, ex_loss, opt = refiner(l.cuda(), h.cuda(), preds.cuda())
    assert ex_loss == 0
    assert np.array_equal(opt["mask"].cpu().numpy(), gold[tag + "_mask"])            # selection: bit-exact
    assert np.array_equal(opt["coords_list"].cpu().numpy(), gold[tag + "_coords"])
    wp, ref_wp = opt["window_preds"].cpu().numpy(), gold[tag + "_window_preds"]
    assert wp.shape == ref_wp.shape
    # bf16 tensor-core path vs fp32 reference: logits of O(1..10); tolerance 3e-2 * max|ref|
    tol = 3e-2 * np.abs(ref_wp).max()
    assert np.abs(wp - ref_wp).max() < tol, (np.abs(wp - ref_wp).max(), tol)
    assert np.abs(out.cpu().numpy() - gold[tag + "_out"]).max() < 3e-2 * max(1.0, np.abs(gold[tag + "_out"]).max())
    np.testing.assert_allclose(opt["GE_w"].cpu().numpy(), gold[tag + "_ge_w"], atol=2e-5)
    # final masks (sigmoid > 0.5 of the refined logits) agree on >= 99.9 % of pixels
    agree = ((out.cpu().numpy() > 0) == (gold[tag + "_out"] > 0)).mean()
    assert agree >= 0.999, agree


def test_token_entry_matches_nchw_entry(refiner):
    l, h, preds = synth_coral_inputs(11, batch=1, uncertain=((0, 0), (2, 2)))
    out1, _, _ = refiner(l.cuda(), h.cuda(), preds.cuda())
    lt = ops.features_to_tokens_f32(l.cuda())
    ht = ops.features_to_tokens_f32(h.cuda().flatten(0, 1)).reshape(1, 9, 56 * 56, 768)
    out2, _, _ = refiner.forward_tokens(lt, ht, preds.cuda(), 56)
    # identical up to the order of the atomically accumulated global sums in the gated ensemble
    assert (out1 - out2).abs().max().item() < 1e-4


def test_resize_tokens_matches_interpolate():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 768, 37, 37, generator=g)
    ref = torch.nn.functional.interpolate(x, size=(56, 56), mode="bilinear")
    tok = ops.features_to_tokens_f32(x.cuda())
    o32, o16 = ops.resize_tokens_bilinear(tok, (37, 37), (56, 56), want_f32=True, want_bf16=True)
    got = o32.reshape(2, 56, 56, 768).permute(0, 3, 1, 2).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-5)
    assert (o16.float().cpu() - o32.cpu()).abs().max() < 0.04


def test_coral_evaluator_end_to_end():
    """Whole second-stage eval of one image vs the CPU oracle pipeline (random-init ViT-B/14 + shipped decoder +
    seeded refiner), nothing injected.  bf16 backbone + bf16 CSF block vs the fp32 oracle: refined logits within 10 %
    of their range, and the margin rule for the integer output — a final mask pixel may differ from the oracle's only
    where the oracle's own interpolated probability is a near-tie (|p - 0.5| < delta)."""
    from safetensors.torch import load_file
    from oracle import pipeline as opipe
    from oracle import vit as ovit
    from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8
    from ucod_dpl_b200.vit import VitKeyExtractor, spec_for
    root = Path(__file__).resolve().parents[1]
    S = 224                                     # small network size keeps the CPU oracle (10 ViT passes) fast
    img = synth_image_u8(3, 300, 340)           # [3,H0,W0] uint8
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = load_file(str(root / "weights" / "UCOD_DPL_dinov2.safetensors"))
    ref_sd = random_refiner_state_dict(0)
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(dec_sd, strict=True)
    refiner = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015))
    refiner.load_state_dict(ref_sd, strict=True)
    ev = CoralEvaluator(VitKeyExtractor(vit_sd, spec_for("dinov2")), model.cuda().eval(), refiner.cuda().eval(),
                        (S, S), window_size=3, window_length=56)
    results, crop = ev.refine(img.unsqueeze(0).cuda())
    masks = ev(img.unsqueeze(0).cuda(), label_sizes=[(300, 340)])
    ref = opipe.coral_eval(vit_sd, ovit.spec_for("dinov2"), dec_sd, ref_sd, img.permute(1, 2, 0).numpy(), S, (300, 340))
    assert bool(crop[0]) == ref["crop"]
    got = results[0].cpu()
    assert got.shape == ref["refined"].shape
    scale = ref["refined"].abs().max().item()
    assert (got - ref["refined"]).abs().max().item() < 0.1 * max(scale, 1.0)
    import torch.nn.functional as F
    p_ref = F.interpolate(torch.sigmoid(ref["refined"]), size=(300, 340), mode="bilinear")[0, 0]
    bad = masks[0].cpu().float() != ref["mask"][0]
    agree = 1.0 - bad.float().mean().item()
    worst = (p_ref[bad] - 0.5).abs().max().item() if bad.any() else 0.0
    print(f"coral e2e: refined max |diff| {(got - ref['refined']).abs().max().item():.4f} (range {scale:.2f}), mask "
          f"agreement {agree:.5f}, {int(bad.sum())} mismatching pixels, largest |p - 0.5| among them {worst:.4f}")
    assert worst < 0.06, worst
    assert agree >= 0.98, agree


def test_coral_glue_matches_oracle():
    from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator as CE
    gold = np.load(GOLD / "coral.npz")
    g = torch.Generator().manual_seed(9)
    p4 = torch.randn(2, 4, 1, 68, 68, generator=g)
    np.testing.assert_allclose(CE.concate_preds(p4.cuda()).cpu().numpy(), gold["concate_preds_out"], atol=1e-6)
    x = torch.randn(1, 1, 168, 168, generator=g)
    assert np.array_equal(CE._center_pad(x.cuda()).cpu().numpy(), gold["center_pad_out"])
    m = CE.process_preds(x.cuda(), (300, 417)).cpu().numpy().astype(np.float32)
    want = gold["process_preds_out"]
    assert (m != want).mean() < 2e-4        # fp32 sigmoid/interp rounding exactly at the 0.5 boundary only
    mp = CE.process_preds(x.sigmoid().cuda(), (300, 417)).cpu().numpy().astype(np.float32)
    assert (mp != want).mean() < 2e-4


def test_evaluator_is_batch_independent():
    """The reference's CORAL eval loop runs at batch 1 and `GatedEnsembler` normalises by `en_local.max()` over the
    whole tensor, so a batched eval must take that maximum per image: image k inside a batch == image k alone."""
    from safetensors.torch import load_file
    from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8
    from ucod_dpl_b200.vit import VitKeyExtractor, spec_for
    root = Path(__file__).resolve().parents[1]
    S = 224
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    model = baseline(SimpleNamespace(dim=768)).cuda().eval()
    model.load_state_dict(load_file(str(root / "weights" / "UCOD_DPL_dinov2.safetensors")), strict=True)
    refiner = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015)).cuda().eval()
    refiner.load_state_dict(random_refiner_state_dict(0), strict=True)
    ev = CoralEvaluator(VitKeyExtractor(vit_sd, spec_for("dinov2")), model, refiner, (S, S), 3, 56)
    imgs = torch.stack([synth_image_u8(20 + i, 300, 340) for i in range(3)]).cuda()
    batched, crop_b = ev.refine(imgs)
    for k in range(3):
        solo, crop_s = ev.refine(imgs[k:k + 1].contiguous())
        assert bool(crop_b[k]) == bool(crop_s[0])
        assert (batched[k] - solo[0]).abs().max().item() < 1e-4
    # the module-level call keeps the reference's whole-batch maximum unless asked otherwise
    l, h, preds = synth_coral_inputs(6, batch=2, uncertain=((1, 2),))
    preds = preds.clone()
    preds[1] = preds[1].clamp(max=-4.0)  # image 1 never reaches the entropy peak at f = 1/e, image 0 does
    hp = torch.randn(2, 1, 168, 168, device="cuda")
    whole, _ = refiner.GE(preds.cuda(), hp)
    per, _ = refiner.GE(preds.cuda(), hp, max_per_image=True)
    one0, _ = refiner.GE(preds[:1].cuda(), hp[:1])
    one1, _ = refiner.GE(preds[1:].cuda(), hp[1:])
    # (the mean of sigmoid(l1) is an atomic float sum: order-dependent in the last bits)
    assert torch.allclose(per[0], one0[0], atol=1e-4) and torch.allclose(per[1], one1[0], atol=1e-4)
    assert not torch.allclose(whole[1], one1[0], atol=1e-3)


def test_entropy_select_range_test_is_per_image_when_asked():
    """ASR.py:42 `preds if all in [0,1] else sigmoid(preds)`: image 0 holds probabilities, image 1 logits.  The whole-call
    test (reference semantics for one call) applies the sigmoid to both; per_image evaluates each like a batch-1 call."""
    g = torch.Generator().manual_seed(3)
    probs = torch.rand(1, 1, 56, 56, generator=g)
    logits = torch.randn(1, 1, 56, 56, generator=g) * 3
    both = torch.cat([probs, logits]).cuda()
    e_call, _, _ = ops.coral_entropy_select(both, 0.0015, 3)
    e_img, _, m_img = ops.coral_entropy_select(both, 0.0015, 3, per_image=True)
    e0, _, m0 = ops.coral_entropy_select(probs.cuda(), 0.0015, 3)
    e1, _, m1 = ops.coral_entropy_select(logits.cuda(), 0.0015, 3)
    assert torch.equal(e_img[0], e0[0]) and torch.equal(e_img[1], e1[0])
    assert torch.equal(m_img[0], m0[0]) and torch.equal(m_img[1], m1[0])
    assert not torch.allclose(e_call[0], e0[0])          # one flag for the call: image 0 went through the sigmoid
    assert torch.equal(e_call[1], e1[0])
