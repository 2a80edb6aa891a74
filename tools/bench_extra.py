#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one B200 (the headline bench.py line is configs[1]):
   pseudo  — configs[2]: APM pseudo-label generation, 256 images @224 per launch
   looktwice — first-stage eval + Look-Twice with 2 planted boxes per image (SURVEY.md §8d "LT" row)
   train   — configs[4]: first-stage training step, 16 cached feature maps per step
   coral   — configs[3]: CORAL second-stage eval, 8 originals @1036^2 per launch (80 ViT passes)
Prints one JSON line per workload with the per-kernel-class breakdown (CUDA events inside the library)."""
from __future__ import annotations

import argparse
import ctypes
import json
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from safetensors.torch import load_file  # noqa: E402

from ucod_dpl_b200 import _lib, ops  # noqa: E402
from ucod_dpl_b200.models.uscod import baseline  # noqa: E402
from ucod_dpl_b200.synth import random_refiner_state_dict, random_vit_state_dict, synth_batch_u8  # noqa: E402
from ucod_dpl_b200.vit import VitKeyExtractor, spec_for  # noqa: E402

NAMES = ["gemm", "attention", "layernorm", "embed", "decoder", "resample", "pseudo_label", "ccl", "other"]


def timed(fn, steps, warmup):
    lib = _lib.load()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    lib.ucod_prof_collect(None, None, None)
    lib.ucod_prof_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    lib.ucod_prof_enable(0)
    ms_c, work_c, n_c = (ctypes.c_double * 9)(), (ctypes.c_double * 9)(), (ctypes.c_longlong * 9)()
    lib.ucod_prof_collect(ms_c, work_c, n_c)
    ms = e0.elapsed_time(e1) / steps
    kern = {}
    for i, n in enumerate(NAMES):
        if n_c[i]:
            k = {"launches_per_step": n_c[i] / steps, "ms_per_step": ms_c[i] / steps}
            rate = work_c[i] / (ms_c[i] * 1e-3) if ms_c[i] else 0.0
            k["tflops" if i < 2 else "gbs"] = rate / (1e12 if i < 2 else 1e9)
            kern[n] = k
    return ms, kern


def blob_logits(B, fs=68, seed=0):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:fs, 0:fs]
    out = np.full((B, 1, fs, fs), -4.0, np.float32)
    for b in range(B):
        for _ in range(2):
            cy, cx = g.uniform(12, fs - 12, 2)
            r = g.uniform(0.07, 0.1) * fs
            out[b, 0] = np.maximum(out[b, 0], 4.0 * (1 - 2 * (((yy - cy) / r) ** 2 + ((xx - cx) / r) ** 2)))
    return torch.from_numpy(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="pseudo,looktwice,train,coral")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    vit_sd = random_vit_state_dict(spec_for("dinov2"), seed=0)
    dec_sd = load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(dec_sd, strict=True)
    model = model.cuda().eval()
    ext = VitKeyExtractor(vit_sd, spec_for("dinov2"), device=dev)
    for wl in args.workloads.split(","):
        if wl == "pseudo":
            from ucod_dpl_b200.generate_pseudo_label import PseudoLabelGenerator
            gen = PseudoLabelGenerator(vit_sd, "dinov2")
            imgs = synth_batch_u8(0, 256, 224, 224).cuda()
            ms, kern = timed(lambda: gen(imgs), args.steps, args.warmup)
            n = 256
        elif wl == "looktwice":
            from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
            B = 32
            ev = LookTwiceEvaluator(ext, model, (518, 518), 68, 0.15, "dynamic")
            imgs = synth_batch_u8(0, B, 518, 518).cuda()
            planted = blob_logits(B).cuda()

            def step():
                ev.first_look(imgs)                       # real first pass (ViT + decoder)
                up, boxes = ev.process_preds(planted)     # planted logits force 2 boxes per image
                return ev.look_twice_batch(imgs, boxes, up.to(torch.uint8))
            ms, kern = timed(step, args.steps, args.warmup)
            n = B
        elif wl == "train":
            from oracle import decoder as odec  # seeded discriminator weights only (test infrastructure data)
            from ucod_dpl_b200.models.discriminator import Discriminator
            from ucod_dpl_b200.train import FirstStageTrainer
            D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68))
            D.load_state_dict(odec.random_discriminator_state_dict(68, seed=31), strict=True)
            m2 = baseline(SimpleNamespace(dim=768))
            m2.load_state_dict(dec_sd, strict=True)
            tr = FirstStageTrainer(m2.cuda().train(), D.cuda().train(), lr0=2e-4)
            tr.cur_epoch = 3
            g = torch.Generator().manual_seed(1)
            tok = torch.randn(16, 1369, 768, generator=g).to(torch.bfloat16).cuda()
            pl = (torch.rand(16, 1, 16, 16, generator=g) < 0.35).float().cuda()
            ms, kern = timed(lambda: tr.process_batch(tok, (37, 37), pl), max(args.steps, 20), max(args.warmup, 5))
            n = 16
        elif wl == "coral":
            from ucod_dpl_b200.engine.runner.loop_CORAL import CoralEvaluator
            from ucod_dpl_b200.models.UDLR import SparseRefiner
            ref = SparseRefiner.from_config(SimpleNamespace(window_size=3, threshold=0.0015))
            ref.load_state_dict(random_refiner_state_dict(0), strict=True)
            ev = CoralEvaluator(ext, model, ref.cuda().eval(), (518, 518), 3, 56)
            imgs = synth_batch_u8(0, 8, 1036, 1036).cuda()
            ms, kern = timed(lambda: ev(imgs), max(2, args.steps // 2), 2)
            n = 8
        else:
            continue
        print(json.dumps({"workload": wl, "images_per_step": n, "ms_per_step": ms, "images_per_s": n / ms * 1e3,
                          "kernels": kern}), flush=True)


if __name__ == "__main__":
    main()
