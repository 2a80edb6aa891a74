"""Look-Twice kernels (CUDA) vs. the oracle: boxes bit-exact, PIL-exact crop/resize and bicubic paste."""
import numpy as np
import pytest
import torch

from oracle import cc as occ
from oracle import looktwice as olt
from oracle import pil_resample as opr
from ucod_dpl_b200 import ops

pytestmark = pytest.mark.gpu


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


def _oracle_boxes(mask01, S, th):
    try:
        return olt.boxes_from_mask(mask01, (S, S), th, "dynamic")
    except ValueError:
        return "ValueError"


@pytest.mark.parametrize("S,th", [(518, 0.15), (296, 0.05)])
def test_boxes_bit_exact(S, th):
    """Integer work given identical inputs: the oracle's component / box logic runs on the very mask the CUDA upsample
    produced, so EVERY case is compared bit for bit; the float step before it (bilinear upsample + threshold) is checked
    against the oracle separately, pixels may differ only where the oracle's interpolated logit is a rounding tie."""
    import torch.nn.functional as F
    rng = np.random.default_rng(S)
    logits = [blob_logits(int(rng.integers(0, 6)), seed=t, size=(0.03, 0.2) if t % 2 else (0.02, 0.08))
              for t in range(24)]
    logits.append(blob_logits(0, seed=99))                          # empty mask -> hard-coded default box
    logits.append(torch.full((1, 1, 68, 68), 4.0))                   # one image-filling component -> None
    lg = torch.cat(logits, 0)
    mask = ops.upsample_bilinear(lg[:, 0].cuda(), (S, S), binarize=True)
    boxes, nbox, status, _ = ops.lt_boxes(mask, th, "dynamic")
    boxes, nbox, mask = boxes.cpu(), nbox.cpu().tolist(), mask.cpu()
    up = F.interpolate(lg, size=(S, S), mode="bilinear", align_corners=False)[:, 0]
    bad = mask != (torch.sigmoid(up) > 0.5).to(torch.uint8)
    assert bad.float().mean().item() < 1e-5 and (not bad.any() or up[bad].abs().max().item() < 1e-5)
    kinds, compared = set(), 0
    for i in range(len(logits)):
        want = _oracle_boxes(mask[i].numpy(), S, th)
        compared += 1
        if want == "ValueError":
            assert nbox[i] == -2
            kinds.add("err")
        elif want is None:
            assert nbox[i] == -1
            kinds.add("none")
        else:
            assert nbox[i] == len(want), (i, nbox[i], want)
            assert boxes[i, :nbox[i]].tolist() == want
            kinds.add("boxes" if want != [olt.DEFAULT_BOX] else "default")
    assert compared == len(logits) == 26
    assert {"none", "boxes", "default"} <= kinds


def test_component_partition_and_ties():
    """random noise masks: component partition equals the oracle's; equal-area ties sort in OpenCV order."""
    rng = np.random.default_rng(0)
    S = 96
    masks = (rng.random((6, S, S)) < 0.45).astype(np.uint8)
    _, _, _, labels = ops.lt_boxes(torch.from_numpy(masks).cuda(), 0.15, "dynamic", want_labels=True)
    labels = labels.cpu().numpy()
    for i in range(6):
        n, ref = occ.connected_components_8(masks[i])
        assert ((labels[i] >= 0) == (ref > 0)).all()
        pairs = set(zip(labels[i][ref > 0].tolist(), ref[ref > 0].tolist()))
        assert len(pairs) == n - 1  # one-to-one between our roots and the oracle's labels
        assert len({p[0] for p in pairs}) == n - 1
    # two identical squares whose pixel-raster order and block-raster order differ
    S = 200
    m = np.zeros((1, S, S), np.uint8)
    m[0, 1:41, 100:140] = 1   # first pixel at row 1
    m[0, 0:40, 150:190] = 1   # first pixel at row 0 (later column) -> OpenCV still labels the left one first
    m[0, 100:140, 20:60] = 1
    boxes, nbox, _, _ = ops.lt_boxes(torch.from_numpy(m).cuda(), 0.15, "const")
    lg = torch.from_numpy(np.where(m[0] > 0, 5.0, -5.0).astype(np.float32))[None, None]
    # oracle on the same mask (feed a logit map that is already at the image size)
    want = []
    n, lab = occ.connected_components_8(m[0] * 255)
    for l in range(1, n):
        binary = (lab == l).astype(np.uint8)
        want.append(olt.expand_bbox(binary, occ.bounding_rect(binary), S, S, expand_type="const"))
    want = sorted(want, key=lambda b: -b[2] * b[3])
    assert boxes[0, :nbox[0].item()].cpu().tolist() == want


def _rand_image(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("H0,W0,S", [(1036, 1036, 518), (700, 933, 518), (296, 296, 296), (400, 300, 296)])
def test_roi_crop_resize_pil_exact(H0, W0, S):
    img = _rand_image(H0, W0, H0 + W0)
    rng = np.random.default_rng(1)
    jobs = [[0, 0, 0, W0, H0], [0, -7, -5, 120, 90], [0, W0 - 50, H0 - 60, 100, 100], [0, 10, 20, S, S],
            [0, 33, 41, 37, 29], [0, 5, 5, 3, 2]]
    for _ in range(6):
        w, h = int(rng.integers(20, W0)), int(rng.integers(20, H0))
        jobs.append([0, int(rng.integers(0, W0 - w + 1)), int(rng.integers(0, H0 - h + 1)), w, h])
    chw = torch.from_numpy(img).permute(2, 0, 1).contiguous()[None].cuda()
    hwc = torch.from_numpy(img)[None].cuda()
    jt = torch.tensor(jobs, dtype=torch.int32).cuda()
    out = ops.roi_crop_resize(chw, jt, (S, S)).cpu().numpy()
    out2 = ops.roi_crop_resize(hwc, jt, (S, S), layout="HWC").cpu().numpy()
    assert np.array_equal(out, out2)
    for j, (_, x, y, w, h) in enumerate(jobs):
        want = opr.resize_u8(opr.crop_u8(img, x, y, x + w, y + h), S, S, "bilinear")
        assert np.array_equal(out[j].transpose(1, 2, 0), want), f"job {j} {jobs[j]}"


def test_paste_bicubic_pil_exact():
    S, g = 518, 37
    rng = np.random.default_rng(3)
    n_img = 3
    base = (rng.random((n_img, S, S)) < 0.3).astype(np.uint8)
    jobs, logit_list = [], []
    geo = [(0, 10, 20, 200, 150), (0, 100, 60, 37, 37), (0, 300, 300, 300, 260), (1, -20, 400, 150, 200),
           (1, 0, 0, 518, 518), (2, 250, 250, 12, 9), (2, 255, 245, 90, 301)]
    ranks = {}
    for (im, x, y, w, h) in geo:
        r = ranks.get(im, 0)
        ranks[im] = r + 1
        jobs.append([im, x, y, w, h, r])
        logit_list.append(rng.normal(0, 1, (g, g)).astype(np.float32))
    canvas = ops.mask_scale_u8(torch.from_numpy(base).cuda(), 255)
    ops.paste_bicubic(torch.from_numpy(np.stack(logit_list)).cuda(), torch.tensor(jobs, dtype=torch.int32).cuda(),
                      canvas)
    got = canvas.cpu().numpy()
    want = base * 255
    for (im, x, y, w, h, r), lg in zip(jobs, logit_list):
        pred = ((torch.sigmoid(torch.from_numpy(lg)) > 0.5).float() * 255).to(torch.uint8).numpy()
        opr.paste_u8(want[im], opr.resize_u8(pred, w, h, "bicubic"), x, y)
    assert np.array_equal(got, want)


def test_look_twice_end_to_end_small():
    """Full second look on one image with planted boxes: crops, ViT, decoder@37, paste — vs. the oracle fed by the
    oracle ViT/decoder.  Pixel agreement >= 99.9 % (bf16 vs fp32 logits near the threshold may differ)."""
    from types import SimpleNamespace
    from oracle import decoder as odec
    from oracle import vit as ovit
    from safetensors.torch import load_file
    from pathlib import Path
    from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.synth import random_vit_state_dict, synth_image_u8
    from ucod_dpl_b200.vit import VitKeyExtractor, spec_for
    S = 224  # small network size keeps the CPU oracle fast; geometry code is size-generic
    spec = ovit.spec_for("dinov2")
    sd = random_vit_state_dict(spec, seed=0)
    dec_sd = load_file(str(Path(__file__).resolve().parents[1] / "weights" / "UCOD_DPL_dinov2.safetensors"))
    model = baseline(SimpleNamespace(dim=768)).cuda().eval()
    model.load_state_dict(dec_sd)
    ev = LookTwiceEvaluator(VitKeyExtractor(sd, spec_for("dinov2")), model, (S, S), 68, 0.15, "dynamic")
    orig = synth_image_u8(7, 448, 448)                       # original-resolution image [3,448,448]
    old = torch.zeros(1, S, S, dtype=torch.uint8)
    old[0, 40:70, 50:90] = 1
    bboxes = [[30, 25, 90, 80], [120, 100, 60, 95]]
    new = ev.look_twice_batch(orig[None].cuda(), [bboxes], old.cuda())

    def seg(x):
        keys = ovit.keys_to_map(ovit.vit_forward(sd, spec, x)["key_tokens"])
        return odec.baseline_forward(dec_sd, keys, want_ortho=False)[0]

    want = olt.look_twice(orig.permute(1, 2, 0).numpy(), bboxes, old.float(), (S, S), seg)
    agree = ((new.cpu() > 0.5) == (want > 0.5)).float().mean().item()
    assert agree >= 0.999, agree
    # outside the boxes nothing may change
    outside = torch.ones(S, S, dtype=torch.bool)
    for (x, y, w, h) in bboxes:
        outside[y:y + h, x:x + w] = False
    assert torch.equal(new.cpu()[0][outside], old[0].float()[outside])


# ---- device-resident control flow (round 2): job tables, device-side counts, order-independent paste ----
def _tiny_evaluator(S=224, **kw):
    from types import SimpleNamespace
    from pathlib import Path
    from safetensors.torch import load_file
    from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
    from ucod_dpl_b200.models.uscod import baseline
    from ucod_dpl_b200.synth import random_vit_state_dict
    from ucod_dpl_b200.vit import VitKeyExtractor, spec_for
    sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = load_file(str(Path(__file__).resolve().parents[1] / "weights" / "UCOD_DPL_dinov2.safetensors"))
    model = baseline(SimpleNamespace(dim=768)).cuda().eval()
    model.load_state_dict(dec_sd)
    return LookTwiceEvaluator(VitKeyExtractor(sd, spec_for("dinov2")), model, (S, S), 68, 0.15, "dynamic", **kw)


def test_build_jobs_matches_host_loop():
    """`ucod_lt_build_jobs` == the reference's host loop (`resize_bbox` in CPython floats, image-major, rank order),
    including ragged original sizes, empty / None images, the capacity clamp and the per-chunk counts."""
    from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import resize_bbox
    S = 518
    rng = np.random.default_rng(5)
    lg = torch.cat([blob_logits(int(rng.integers(0, 5)), seed=100 + t, size=(0.03, 0.1)) for t in range(12)], 0)
    mask = ops.upsample_bilinear(lg[:, 0].cuda(), (S, S), binarize=True)
    boxes, nbox, _, _ = ops.lt_boxes(mask, 0.15, "dynamic")
    sizes = torch.tensor([[int(rng.integers(300, 1400)), int(rng.integers(300, 1400))] for _ in range(12)])
    for cap, chunk in ((64, 16), (7, 4)):
        crop, paste, counts, chunks = ops.lt_build_jobs(boxes, nbox, (S, S), None, sizes, capacity=cap, chunk=chunk)
        want_c, want_p = [], []
        for b, n in enumerate(nbox.cpu().tolist()):
            for r in range(max(n, 0)):
                bb = boxes[b, r].cpu().tolist()
                H0, W0 = sizes[b].tolist()
                want_c.append([b] + resize_bbox(bb, S, S, W0, H0))
                want_p.append([b] + bb + [r])
        kept, status, wanted, _ = counts.cpu().tolist()
        assert wanted == len(want_c) and kept == min(cap, wanted) and bool(status & 2) == (wanted > cap)
        assert crop[:kept].cpu().tolist() == want_c[:kept]
        assert paste[:kept].cpu().tolist() == want_p[:kept]
        assert chunks.cpu().tolist() == [max(0, min(chunk, kept - c * chunk)) for c in range((cap + chunk - 1) // chunk)]
        assert wanted > 7  # the clamp case is exercised


def test_vit_and_decoder_device_count():
    """`count_dev` variants: the first n images equal the plain call bit for bit, for n = 0 .. capacity."""
    from ucod_dpl_b200.synth import synth_batch_u8
    ev = _tiny_evaluator()
    imgs = synth_batch_u8(3, 5, 224, 224).cuda()
    _, k_ref, _ = ev.extractor.keys(imgs, want_f32=False, want_bf16=True)
    fg_ref, _, _ = ev.model.decoder.forward_tokens(k_ref, (16, 16), (16, 16), want_bg=False)
    for n in (0, 1, 3, 5):
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        _, k, _ = ev.extractor.keys(imgs, want_f32=False, want_bf16=True, count_dev=cnt)
        fg, _, _ = ev.model.decoder.forward_tokens(k, (16, 16), (16, 16), want_bg=False, count_dev=cnt)
        assert torch.equal(k[:n], k_ref[:n]), n
        # the decoder's per-channel sum of squares is accumulated with float atomics: equal up to summation order
        assert torch.allclose(fg[:n], fg_ref[:n], rtol=1e-5, atol=1e-5), n


def test_paste_order_independent_overlaps():
    """`paste_bicubic_dyn` (one launch per chunk, later boxes win per pixel) == sequential PIL pastes, with boxes that
    overlap each other, stick out of the mask and straddle a chunk boundary."""
    S, g = 296, 37
    rng = np.random.default_rng(11)
    base = (rng.random((3, S, S)) < 0.3).astype(np.uint8)
    geo = [(0, 10, 20, 200, 150), (0, 100, 60, 137, 137), (0, 150, 100, 100, 160), (1, -20, 200, 150, 120),
           (1, 0, 0, 296, 296), (1, 40, 40, 3, 2), (2, 250, 250, 12, 9), (2, 255, 245, 60, 70), (2, 200, 230, 90, 30)]
    jobs, rank = [], {}
    for (im, x, y, w, h) in geo:
        jobs.append([im, x, y, w, h, rank.get(im, 0)])
        rank[im] = rank.get(im, 0) + 1
    lg = rng.normal(0, 1, (len(geo), g, g)).astype(np.float32)
    want = base * 255
    for (im, x, y, w, h, r), l in zip(jobs, lg):
        pred = ((torch.sigmoid(torch.from_numpy(l)) > 0.5).float() * 255).to(torch.uint8).numpy()
        opr.paste_u8(want[im], opr.resize_u8(pred, w, h, "bicubic"), x, y)
    cap, chunk = 12, 4
    table = torch.zeros(cap, 6, dtype=torch.int32)
    table[:len(jobs)] = torch.tensor(jobs, dtype=torch.int32)
    table = table.cuda()
    n_all = torch.tensor([len(jobs)], dtype=torch.int32, device="cuda")
    canvas = ops.mask_scale_u8(torch.from_numpy(base).cuda(), 255)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for c in reversed(range(cap // chunk)):            # chunk order must not matter
        n_c = torch.tensor([max(0, min(chunk, len(jobs) - c * chunk))], dtype=torch.int32, device="cuda")
        logits = torch.zeros(chunk, g, g)
        logits[:int(n_c)] = torch.from_numpy(lg[c * chunk:c * chunk + int(n_c)])
        ops.paste_bicubic_dyn(logits.cuda(), table, c * chunk, n_c, n_all, canvas, 444, err=err)
    assert int(err.item()) == 0
    assert np.array_equal(canvas.cpu().numpy(), want)


def test_crop_resize_dyn_equals_static_and_large_downscale():
    """device-count crop/resize == the static call on the valid jobs; a > 19x down-scale (beyond the static tap table)
    is PIL-exact on the dynamic path and raises on the static one."""
    from ucod_dpl_b200.synth import synth_batch_u8
    imgs = synth_batch_u8(1, 2, 600, 700).cuda()
    jobs = torch.tensor([[0, 10, 20, 300, 200], [1, -30, 50, 400, 590], [1, 100, 100, 64, 64], [0, 0, 0, 700, 600]],
                        dtype=torch.int32).cuda()
    ref = ops.roi_crop_resize(imgs, jobs, (224, 224))
    table = torch.zeros(6, 5, dtype=torch.int32, device="cuda")
    table[:4] = jobs
    for n in (4, 2, 0):
        got = ops.roi_crop_resize_dyn(imgs, table, torch.tensor([n], dtype=torch.int32, device="cuda"), (224, 224))
        assert torch.equal(got[:n], ref[:n])
    big = synth_batch_u8(2, 1, 1300, 900).cuda()
    job = torch.tensor([[0, 0, 0, 900, 1300]], dtype=torch.int32).cuda()
    got = ops.roi_crop_resize_dyn(big, job, torch.tensor([1], dtype=torch.int32, device="cuda"), (40, 40))
    want = np.stack([opr.resize_u8(big[0, c].cpu().numpy(), 40, 40, "bilinear") for c in range(3)])
    assert np.array_equal(got[0].cpu().numpy(), want)
    with pytest.raises(RuntimeError):
        ops.roi_crop_resize(big, job, (40, 40))


def test_device_pipeline_equals_host_lists():
    """`look_twice_device` (no host synchronisation, chunks with device-side counts) produces bit-identical masks to
    the host-list path (`process_preds` + `look_twice_batch`) on planted objects, ragged originals included."""
    from ucod_dpl_b200.data.datasets.transforms import pack_padded
    from ucod_dpl_b200.synth import planted_object_logits, synth_batch_u8, synth_image_u8
    S = 224
    ev = _tiny_evaluator(S, max_looks_per_image=3)
    imgs = synth_batch_u8(0, 6, S, S).cuda()
    planted = torch.stack([planted_object_logits(70 + i, 68, k) for i, k in enumerate((2, 0, 3, 1, 3, 2))]).cuda()
    originals = [synth_image_u8(20 + i, 300 + 37 * i, 420 - 29 * i).permute(1, 2, 0).contiguous().numpy() for i in range(6)]
    canvas, sizes = pack_padded(originals, "cuda")
    res = ev.look_twice_device(imgs, canvas, layout="HWC", orig_sizes=sizes, first_logits=planted)
    res.check()
    up, boxes = ev.process_preds(planted)
    assert torch.equal(up.to(torch.uint8), res.first) and boxes == res.bboxes
    want = ev.look_twice_batch(canvas, boxes, up.to(torch.uint8), layout="HWC", orig_sizes=sizes)
    assert torch.equal(res.final, want)
    kept, status, wanted, _ = res.counts.cpu().tolist()
    assert status == 0 and kept == wanted == sum(len(b) for b in boxes if b) and kept > 6
    # the number of chunks enqueued ahead of time is only a launch-count guess: with too few, check() completes the batch
    assert ev._recent_chunks == [1]            # 12 second looks fit one 16-job chunk
    ev._recent_chunks = [0]
    res2 = ev.look_twice_device(imgs, canvas, layout="HWC", orig_sizes=sizes, first_logits=planted)
    assert not torch.equal(res2.final, want)          # no second look enqueued ahead of time
    res2.check()
    assert torch.equal(res2.final, want) and ev._recent_chunks == [0, 1]
    # more second looks than the device job table holds: check() falls back to the host-list path, same result
    ev1 = _tiny_evaluator(S, max_looks_per_image=1)
    imgs20 = imgs.repeat(4, 1, 1, 1)[:20]
    canvas20, sizes20 = canvas.repeat(4, 1, 1, 1)[:20], sizes.repeat(4, 1)[:20]
    planted20 = planted.repeat(4, 1, 1, 1)[:20]
    r1 = ev1.look_twice_device(imgs20, canvas20, layout="HWC", orig_sizes=sizes20, first_logits=planted20)
    kept1, status1, wanted1, _ = r1.counts.cpu().tolist()
    assert status1 & 2 and wanted1 > kept1 == 20
    r1.check()
    assert torch.equal(r1.final[:6], want)


def test_shared_memory_labeller_equals_global():
    """The run-based shared-memory labeller (one CTA per mask) and the global-memory union-find produce identical box
    tables, summaries and label images — on blob masks at both pipeline sizes, noise masks (thousands of components),
    stripes / checkerboards (the run-capacity corner) and degenerate masks."""
    rng = np.random.default_rng(21)
    for S in (518, 296, 97):
        masks = []
        for t in range(6):
            lg = blob_logits(int(rng.integers(0, 7)), seed=300 + t, size=(0.02, 0.15))
            masks.append(ops.upsample_bilinear(lg[:, 0].cuda(), (S, S), binarize=True)[0].cpu().numpy())
        masks.append((rng.random((S, S)) < 0.5).astype(np.uint8))             # noise: very many components
        masks.append((rng.random((S, S)) < 0.08).astype(np.uint8))
        m = np.zeros((S, S), np.uint8); m[::2] = 1; masks.append(m)          # horizontal stripes
        m = np.zeros((S, S), np.uint8); m[:, ::2] = 1; masks.append(m)       # vertical stripes: W/2 runs per row
        masks.append(np.zeros((S, S), np.uint8))
        masks.append(np.ones((S, S), np.uint8))
        m = np.zeros((S, S), np.uint8); m[0, 0] = m[-1, -1] = m[0, -1] = 1; masks.append(m)
        mk = torch.from_numpy(np.stack(masks)).cuda()
        a = ops.lt_boxes(mk, 0.15, "dynamic", want_labels=True, algorithm="global")
        b = ops.lt_boxes(mk, 0.15, "dynamic", want_labels=True, algorithm="shared")
        nb_a, nb_b = a[1].cpu().tolist(), b[1].cpu().tolist()
        for i in range(len(masks)):
            if nb_b[i] == -3:                      # capacity: only the dense-noise / stripe masks may hit it
                assert i in (6, 7, 9), (S, i)   # more components than bit-image words, or more runs than slots
                continue
            assert nb_a[i] == nb_b[i], (S, i, nb_a[i], nb_b[i])
            n = max(nb_a[i], 0)
            assert torch.equal(a[0][i, :n], b[0][i, :n]), (S, i)
            assert torch.equal(a[3][i], b[3][i]), (S, i)
        assert sum(1 for v in nb_b if v != -3) >= len(masks) - 3
