"""`RevDecoder` — Dual-Branch Adversarial decoder with the reference's constructor, parameters and
`state_dict` keys (models/modules/DBA.py:5-59); the forward runs in `csrc/decoder.cu`.

    decoupling.{weight[128,dim,1,1],bias[128]}  learnable_embedding[2,64]
    conv_out_fg.{weight[1,64,1,1],bias[1]}      conv_out_bg.{weight[1,64,1,1],bias[1]}
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...engine.registry import MODULE_REGISTRY


@MODULE_REGISTRY.register()
class RevDecoder(nn.Module):
    def __init__(self, cfg, ema: bool = False):
        super().__init__()
        feature_dim = cfg.dim
        self.embed_dim = 64
        self.decoupling = nn.Conv2d(feature_dim, 2 * self.embed_dim, kernel_size=(1, 1))
        self.learnable_embedding = nn.Parameter(torch.randn(2, self.embed_dim))
        self.conv_out_fg = nn.Conv2d(self.embed_dim, 1, kernel_size=(1, 1))
        self.conv_out_bg = nn.Conv2d(self.embed_dim, 1, kernel_size=(1, 1))
        self.ema = ema
        if ema:
            for p in self.parameters():
                p.detach_()
        self._packed = None  # (versions, bf16 decoupling weight)

    # bf16 copy of the 1x1 conv weight, refreshed whenever the parameter is modified in place
    def _w_dec_bf16(self) -> torch.Tensor:
        w = self.decoupling.weight
        tag = (w._version, w.data_ptr(), w.device)
        if self._packed is None or self._packed[0] != tag:
            self._packed = (tag, w.detach().reshape(w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous())
        return self._packed[1]

    def forward_tokens(self, keys_bf16: torch.Tensor, grid_in, grid_out, want_bg=True, want_ortho=False,
                       count_dev=None):
        """Fused path: token-major bf16 keys on `grid_in`, logits on `grid_out` (the bilinear feature upsample of
        loop_UCOD_DPL.py:153,305 is folded into the decoder)."""
        return ops.decoder_forward(
            keys_bf16, grid_in, grid_out, self._w_dec_bf16(), self.decoupling.bias.detach().float(),
            self.learnable_embedding.detach().float().contiguous(), self.conv_out_fg.weight.detach().reshape(-1),
            self.conv_out_fg.bias.detach(), self.conv_out_bg.weight.detach().reshape(-1),
            self.conv_out_bg.bias.detach(), want_bg=want_bg, want_ortho=want_ortho, count_dev=count_dev)

    def calc_orthogonal_loss(self, feature_1, feature_2, weight=1.0):
        raise NotImplementedError("the orthogonality loss is fused into forward() (Gram identity, csrc/decoder.cu)")

    def forward(self, x, get_bg_mask: bool = False):
        if type(x) is list:
            x = x[-1]
        B, _, H, W = x.shape
        tokens = ops.features_to_tokens_bf16(x)
        if not self.ema and self.training and torch.is_grad_enabled():
            from ...train import decoder_forward_autograd  # training path (fwd + hand-written bwd kernels)
            self._packed = None  # the optimiser may have stepped since the last call
            return decoder_forward_autograd(self, tokens, (H, W))
        fg, bg, ortho = self.forward_tokens(tokens, (H, W), (H, W), want_bg=(not self.ema) or get_bg_mask,
                                            want_ortho=not self.ema)
        if not self.ema:
            return fg, bg, ortho
        if get_bg_mask:
            return fg, bg
        return fg
