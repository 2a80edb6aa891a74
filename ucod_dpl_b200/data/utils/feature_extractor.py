"""`backbone` with the reference's constructor / forward contract (data/utils/feature_extractor.py:31-59):
frozen DINO / DINOv2 ViT-B, returns the last layer's key tokens as a [B,768,h,w] map.

Differences that are deliberate and documented: the forward runs in `csrc/vit.cu` (no HF module is executed);
`outputs` (the HF model output the reference returns first and every caller discards) is None; when neither
`config.backbone_weights` nor the HF cache holds a checkpoint (the offline case) the weights are seeded random
init and a warning is logged, instead of a download attempt."""
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


def load_vit_state_dict(config) -> dict:
    """HF-format state_dict for `config.backbone` from local files only; falls back to seeded random init."""
    spec = spec_for(config.type)
    candidates = []
    for attr in ("backbone_weights", "backbone_weight_base"):
        p = getattr(config, attr, None)
        if p:
            candidates.append(Path(str(p)).expanduser())
    for base in candidates:
        for f in ([base] if base.is_file() else list(base.glob("**/model.safetensors")) if base.is_dir() else []):
            try:
                from safetensors.torch import load_file
                sd = load_file(str(f))
                if "embeddings.cls_token" in sd and "encoder.layer.0.attention.attention.key.weight" in sd:
                    logger.info("loaded ViT weights from %s", f)
                    return sd
            except Exception:  # not a ViT checkpoint
                continue
    logger.warning("no local %s checkpoint found; using seeded random-init ViT-B weights", config.backbone)
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
