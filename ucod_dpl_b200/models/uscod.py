"""`baseline` — student + EMA `RevDecoder` pair (reference models/uscod.py:9-22), same attribute names
(`decoder`, `decoder_ema`) and therefore the same 14 `state_dict` keys as weights/UCOD_DPL_dinov{1,2}.safetensors."""
from __future__ import annotations

import torch
from torch import nn

from ..engine.registry import MODULE_REGISTRY
from .modules.DBA import RevDecoder


@MODULE_REGISTRY.register()
class baseline(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.decoder = RevDecoder(cfg)
        self.decoder_ema = RevDecoder(cfg, ema=True)

    def forward(self, batched_inputs, ema: bool = False):
        if ema:
            with torch.no_grad():
                return self.decoder_ema(batched_inputs)
        return self.decoder(batched_inputs)
