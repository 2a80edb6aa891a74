"""`backbone` with the reference's constructor / forward contract (data/utils/feature_extractor.py:31-59):
frozen DINO / DINOv2 ViT-B, returns the last layer's key tokens as a [B,768,h,w] map.

Differences that are deliberate and documented: the forward runs in `csrc/vit.cu` (no HF module is executed);
`outputs` (the HF model output the reference returns first and every caller discards) is None; when neither
`config.backbone_weights` nor `config.backbone_weight_base` holds a matching checkpoint the constructor raises instead
of attempting a download; seeded random weights need an explicit opt-in (see `load_vit_state_dict`)."""
from __future__ import annotations

import logging
import math
from pathlib import Path

import torch
from torch import nn

from ...engine.registry import BACKBONE_REGISTRY
from ...synth import random_vit_state_dict
from ...vit import VitKeyExtractor, spec_for

logger = logging.getLogger("ucod_dpl_b200")


ALLOW_RANDOM_ENV = "UCOD_B200_ALLOW_RANDOM_BACKBONE"


def _matches(sd: dict, spec) -> bool:
    """is this HF state_dict the requested backbone?  (patch size and LayerScale tell DINO ViT-B/8 from DINOv2-B/14:
    the dinov1 configs need both checkpoints under ./weights, and a generic `**/model.safetensors` glob finds both)"""
    w = sd.get("embeddings.patch_embeddings.projection.weight")
    if w is None or "encoder.layer.0.attention.attention.key.weight" not in sd or "embeddings.cls_token" not in sd:
        return False
    has_ls = "encoder.layer.0.layer_scale1.lambda1" in sd
    return tuple(w.shape) == (spec.hidden, 3, spec.patch, spec.patch) and has_ls == spec.layerscale


def load_vit_state_dict(config, allow_random_init: bool | None = None) -> dict:
    """HF-format state_dict for `config.backbone` from local files only (`config.backbone_weights`, then
    `config.backbone_weight_base`), matched to the requested architecture.
    No checkpoint found: raises, unless random weights were asked for explicitly — `allow_random_init=True`,
    `config.allow_random_init`, or the environment variable UCOD_B200_ALLOW_RANDOM_BACKBONE=1 (tests, benchmarks,
    offline boxes).  A silently random backbone would fill the feature / pseudo-label caches with garbage that later
    runs with real weights would reuse."""
    import os
    spec = spec_for("dinov2" if "dinov2" in str(getattr(config, "backbone", config.type)) else "dinov1")
    candidates = []
    for attr in ("backbone_weights", "backbone_weight_base"):
        p = getattr(config, attr, None)
        if p:
            candidates.append(Path(str(p)).expanduser())
    for base in candidates:
        for f in ([base] if base.is_file() else sorted(base.glob("**/model.safetensors")) if base.is_dir() else []):
            try:
                from safetensors.torch import load_file
                sd = load_file(str(f))
            except Exception:  # not a safetensors file
                continue
            if _matches(sd, spec):
                logger.info("loaded ViT weights from %s", f)
                return sd
    if allow_random_init is None:
        allow_random_init = bool(getattr(config, "allow_random_init", False)) or os.environ.get(ALLOW_RANDOM_ENV) == "1"
    if not allow_random_init:
        raise FileNotFoundError(
            f"no local {getattr(config, 'backbone', config.type)} checkpoint (HF safetensors, patch {spec.patch}) under "
            f"{[str(c) for c in candidates]}; set backbone_weights, or opt into seeded random weights with "
            f"allow_random_init / {ALLOW_RANDOM_ENV}=1")
    logger.warning("no local %s checkpoint found; using seeded random-init ViT-B weights (explicit opt-in)",
                   getattr(config, "backbone", config.type))
    return random_vit_state_dict(spec, seed=0)


@BACKBONE_REGISTRY.register()
class backbone(nn.Module):
    def __init__(self, config, state_dict: dict | None = None, device="cuda") -> None:
        super().__init__()
        assert config.backbone_type == "huggingface"
        if "dino" not in config.type:
            raise ValueError(f"Unsupported model type: {config.type}")
        self.config = config
        self.spec = spec_for("dinov2" if "dinov2" in config.backbone else "dinov1")
        self.feature_extractor = VitKeyExtractor(state_dict or load_vit_state_dict(config), self.spec, device=device)
        self.key = None

    def forward(self, input, reshape_keys: bool = True):
        with torch.no_grad():
            k32, _, _ = self.feature_extractor.keys(input.to(self.feature_extractor.device), want_f32=True,
                                                    keep_cls=not reshape_keys)
        self.key = k32
        if reshape_keys:
            B, L, C = k32.shape
            H = W = int(math.isqrt(L))
            self.key = k32.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return None, self.key
