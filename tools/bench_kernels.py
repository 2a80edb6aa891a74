#!/usr/bin/env python
"""Stand-alone timing of the memory-bound kernels of the path against the HBM roofline (MEASURED_PEAKS.json).
Algorithmic bytes per unit follow SURVEY.md §8(d) / DESIGN.md §4.  One JSON line per kernel."""
from __future__ import annotations

import json
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ucod_dpl_b200 import ops  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()            # > L2 write between iterations
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def report(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 2),
                      "achieved_GBs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3), "note": note}), flush=True)


def main():
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    # pseudo-label scoring: 1024 images, bf16 keys [B,256,768] + attn [B,12,256]
    B = 1024
    keys = torch.randn(B, 256, 768, device=dev, generator=g).to(torch.bfloat16)
    att = torch.softmax(torch.randn(B, 12, 257, device=dev, generator=g), -1)[..., 1:].contiguous()
    ms = timeit(lambda: ops.pseudo_label_score(att, keys, 0.6), flush=flush)
    report("pseudo_label_score (bf16 keys)", ms, B * (256 * 768 * 2 + 12 * 256 * 4 + 256 * 5 + 4), f"{B} images @16x16")
    keys32 = keys.float()
    ms = timeit(lambda: ops.pseudo_label_score(att, keys32, 0.6), flush=flush)
    report("pseudo_label_score (fp32 keys)", ms, B * (256 * 768 * 4 + 12 * 256 * 4 + 256 * 5 + 4), f"{B} images @16x16")
    m = (torch.rand(B * 16, 16, 16, device=dev, generator=g) < 0.5).to(torch.uint8)
    ms = timeit(lambda: ops.refine_small_components(m), flush=flush)
    report("refine_small_components", ms, m.numel() * 2, f"{m.shape[0]} masks 16x16")
    # logits -> binarised 518^2 masks
    B = 256
    lg = torch.randn(B, 68, 68, device=dev, generator=g)
    ms = timeit(lambda: ops.upsample_bilinear(lg, (518, 518), binarize=True), flush=flush)
    report("upsample_bilinear+binarise 68->518", ms, B * (68 * 68 * 4 + 518 * 518), f"{B} images")
    # CCL + boxes on blob masks
    yy, xx = np.mgrid[0:518, 0:518]
    rng = np.random.default_rng(0)
    masks = np.zeros((B, 518, 518), np.uint8)
    for b in range(B):
        for _ in range(3):
            cy, cx, r = rng.uniform(60, 450), rng.uniform(60, 450), rng.uniform(20, 50)
            masks[b] |= (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).astype(np.uint8)
    mk = torch.from_numpy(masks).to(dev)
    ms = timeit(lambda: ops.lt_boxes(mk, 0.15, "dynamic"), flush=flush)
    report("lt_boxes (CCL + area + boxes)", ms, B * (518 * 518 * (1 + 8)), f"{B} masks 518^2; credited: mask read + one label write/read")
    # ROI crop + antialiased resize to 518^2 (2 boxes per image, ~100 px boxes scaled 2x in the original)
    imgs = torch.randint(0, 256, (64, 3, 1036, 1036), device=dev, dtype=torch.uint8, generator=g)
    jobs = torch.tensor([[n, 100 + 7 * n, 200 + 3 * n, 260, 220] for n in range(64)] +
                        [[n, 500, 300 + 2 * n, 180, 300] for n in range(64)], dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.roi_crop_resize(imgs, jobs, (518, 518)), flush=flush)
    nb = sum(int(j[3]) * int(j[4]) * 3 for j in jobs.cpu().tolist()) + len(jobs) * 3 * 518 * 518
    report("roi_crop_resize -> 518^2", ms, nb, f"{len(jobs)} boxes")
    # bicubic paste of 37^2 predictions
    canvas = torch.zeros(64, 518, 518, device=dev, dtype=torch.uint8)
    logits = torch.randn(128, 37, 37, device=dev, generator=g)
    pj = torch.tensor([[n % 64, 50 + n, 60 + n, 130, 110, n // 64] for n in range(128)], dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.paste_bicubic(logits, pj, canvas), flush=flush)
    report("paste_bicubic 37^2 -> box", ms, 128 * (37 * 37 * 4 + 130 * 110), "128 boxes of 130x110")
    # decoder heads on 64 images (keys bf16 [64,1369,768] -> fg @68^2), GEMM included
    from safetensors.torch import load_file
    from ucod_dpl_b200.models.uscod import baseline
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors")), strict=True)
    model = model.cuda().eval()
    k16 = torch.randn(64, 1369, 768, device=dev, generator=g).to(torch.bfloat16)
    ms = timeit(lambda: model.decoder.forward_tokens(k16, (37, 37), (68, 68), want_bg=False), flush=flush)
    report("decoder forward (1x1 conv GEMM + norm + gate + heads)", ms, 64 * (1369 * 768 * 2 + 4624 * 4), "64 images")
    # APM merge
    pl = torch.rand(256, 1, 68, 68, device=dev, generator=g)
    tm = (torch.rand(256, 1, 68, 68, device=dev, generator=g) > 0.5).float()
    ps, pp = torch.rand(256, 1, device=dev, generator=g), torch.rand(256, 1, device=dev, generator=g)
    ms = timeit(lambda: ops.apm_merge(pl, tm, ps, pp, 0.15), flush=flush)
    report("apm_merge", ms, 256 * 68 * 68 * 12, "256 images")
    # LayerNorm rows
    x = torch.randn(64 * 1370, 768, device=dev, generator=g)
    w, bb = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    ms = timeit(lambda: ops.layernorm_bf16(x, w, bb, 1e-6), flush=flush)
    report("layernorm fp32 -> bf16", ms, x.numel() * 6, "87680 rows x 768")


if __name__ == "__main__":
    main()
