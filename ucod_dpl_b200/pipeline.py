"""End-to-end device pipelines assembled from the C-ABI kernels.

`FirstStageEval` is BASELINE.json config 2 (UCOD-DPL first-stage eval): frozen ViT key extraction (a1) ->
DBA decoder on the 68^2 grid (a7, with the 37->68 feature upsample folded in) -> bilinear upsample to the
image size + `sigmoid > 0.5` (a12, first half).  It follows engine/runner/loop_UCOD_DPL.py:297-311,354-361
but batched: every stage is per-image independent, so B images run in one launch sequence.
"""
from __future__ import annotations

import torch

from . import ops
from .vit import VitKeyExtractor, VitSpec


class FirstStageEval:
    def __init__(self, vit_state_dict: dict, vit_spec: VitSpec, model, image_size, feature_size: int = 68,
                 device="cuda"):
        self.device = torch.device(device)
        self.extractor = VitKeyExtractor(vit_state_dict, vit_spec, device=self.device)
        self.model = model.to(self.device).eval()
        self.image_size = tuple(image_size)
        self.feature_size = int(feature_size)
        self.patch = vit_spec.patch

    @torch.no_grad()
    def logits(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,3,S,S] uint8 (raw RGB) or fp32 (normalised), CUDA -> student fg logits [B,1,fs,fs]."""
        _, k16, _ = self.extractor.keys(images, want_f32=False, want_bf16=True)
        gh, gw = images.shape[-2] // self.patch, images.shape[-1] // self.patch
        fs = self.feature_size
        fg, _, _ = self.model.decoder.forward_tokens(k16, (gh, gw), (fs, fs), want_bg=False, want_ortho=False)
        return fg

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        """-> uint8 {0,1} masks [B,S,S] at the config's image size (process_preds' `preds_up`)."""
        fg = self.logits(images)
        return ops.upsample_bilinear(fg[:, 0], self.image_size, binarize=True)
