"""`baseline` — the first-stage model: a student `RevDecoder` and its EMA teacher.

Drop-in for the reference class of the same name (models/uscod.py:9-22): the constructor takes `cfg.model_cfg`,
the two sub-modules are called `decoder` and `decoder_ema` (hence the same 14 `state_dict` keys as
weights/UCOD_DPL_dinov{1,2}.safetensors, loadable with strict=True), and `forward(x, ema=False)` returns the
student's `(fg, bg, extra_loss)` or, with `ema=True`, the teacher's `fg` without autograd.

Beyond that contract the device pipelines use the token-major entry points below, which skip the NCHW round trip:
backbone keys stay `[B, tokens, 768]` bf16 from the ViT kernels into `ucod_decoder_fwd`.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from ..engine.registry import MODULE_REGISTRY
from .modules.DBA import RevDecoder


@MODULE_REGISTRY.register()
class baseline(nn.Module):
    _BRANCHES = {False: "decoder", True: "decoder_ema"}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        for ema, name in self._BRANCHES.items():
            self.add_module(name, RevDecoder(cfg, ema=ema))

    def branch(self, ema: bool = False) -> RevDecoder:
        """the student (`ema=False`) or the teacher decoder."""
        return getattr(self, self._BRANCHES[bool(ema)])

    def forward(self, batched_inputs, ema: bool = False):
        net = self.branch(ema)
        if not ema:
            return net(batched_inputs)
        with torch.no_grad():
            return net(batched_inputs)

    @torch.no_grad()
    def forward_tokens(self, key_tokens_bf16: torch.Tensor, grid_in: Tuple[int, int], grid_out: Tuple[int, int],
                       ema: bool = False, **kw):
        """inference on token-major keys `[B, gh*gw, dim]` bf16 (see `RevDecoder.forward_tokens`)."""
        return self.branch(ema).forward_tokens(key_tokens_bf16, grid_in, grid_out, **kw)

    @torch.no_grad()
    def sync_teacher(self) -> None:
        """teacher <- student (what a fresh reference model has after its first EMA step with alpha = 0)."""
        for pt, ps in zip(self.decoder_ema.parameters(), self.decoder.parameters()):
            pt.copy_(ps)
