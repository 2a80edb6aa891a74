"""`SparseRefiner` — CORAL second-stage refiner with the reference's constructor, `from_config`, forward
signature / return structure and `state_dict` keys (models/UDLR.py:9-86).  Eval path only: the reference ships no
CORAL training loop (`LocalRefineTrainLoop` is `pass`, engine/runner/loop_CORAL.py:38-39) and `cal_ex_loss`
returns 0 outside training (:54-55)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..engine.registry import MODULE_REGISTRY
from .modules.refiner import HRE, EntropySelector, GatedEnsembler


@MODULE_REGISTRY.register()
class SparseRefiner(nn.Module):
    def __init__(self, config, window_size: int, threshold: float, dim: int = 768) -> None:
        super().__init__()
        self.config = config
        self.selector = EntropySelector(threshold, window_size)
        self.HRE = HRE(window_size, dim)
        self.GE = GatedEnsembler(1)
        self.window_size = window_size
        self.threshold = threshold

    @classmethod
    def from_config(cls, config):
        return cls(config, config.window_size, config.threshold)

    def cal_ex_loss(self, opt):
        if self.training:
            raise NotImplementedError("CORAL training is not part of the reference release (loop_CORAL.py:38-39)")
        return 0, opt

    @torch.no_grad()
    def forward_tokens(self, l_tokens, h_tokens_all, preds, grid: int, per_image: bool = False):
        """Device-pipeline entry: l_tokens fp32 [B,g*g,C]; h_tokens_all fp32 [B, w*w, g*g, C] (all windows,
        token-major); preds [B,1,P,P].  Same outputs as `forward`.  per_image: every image is treated as its own
        batch-1 call (the gated ensemble's entropy maximum and the selector's probabilities-or-logits test are per image)
        — what the reference's eval loop computes."""
        mask, entropy, win_img, coords, flat = self.selector.select(preds, per_image=per_image)
        dev = preds.device
        B = preds.shape[0]
        N = len(win_img)
        if N:
            idx = torch.nonzero(flat.flatten()).flatten().to(dev)
            h_sel = h_tokens_all.flatten(0, 1).index_select(0, idx)
            window_preds = self.HRE.CSF.forward_tokens(l_tokens, h_sel, torch.tensor(win_img, dtype=torch.int32,
                                                                                     device=dev), grid)
        else:
            window_preds = torch.zeros(0, 1, grid, grid, device=dev)
        h_preds = self.HRE.concate_windows(window_preds, coords, mask)
        outputs, ge_w = self.GE(preds, h_preds, max_per_image=per_image)
        opt = {"mask": mask, "entropy": entropy, "h_preds": h_preds, "window_preds": window_preds, "GE_w": ge_w,
               "preds": preds, "coords_list": coords.to(dev), "h_targets": None}
        return outputs, 0, opt

    @torch.no_grad()
    def forward(self, input_features, h_inputs, preds, h_targets=None):
        """input_features [B,C,g,g], h_inputs [B,w*w,C,g,g], preds [B,1,P,P] -> (outputs [B,1,w*g,w*g], ex_loss, opt)."""
        B, nw, C, g, _ = h_inputs.shape
        l_tokens = ops.features_to_tokens_f32(input_features)
        # only the selected windows are transposed to token-major (<= 9 x 9.6 MB each)
        mask, entropy, win_img, coords, flat = self.selector.select(preds)
        dev = preds.device
        if win_img:
            idx = torch.nonzero(flat.flatten()).flatten().to(dev)
            h_sel = ops.features_to_tokens_f32(h_inputs.flatten(0, 1).index_select(0, idx))
            window_preds = self.HRE.CSF.forward_tokens(l_tokens, h_sel, torch.tensor(win_img, dtype=torch.int32,
                                                                                     device=dev), g)
        else:
            window_preds = torch.zeros(0, 1, g, g, device=dev)
        h_preds = self.HRE.concate_windows(window_preds, coords, mask)
        outputs, ge_w = self.GE(preds, h_preds)
        opt = {"mask": mask, "entropy": entropy, "h_preds": h_preds, "window_preds": window_preds, "GE_w": ge_w,
               "preds": preds, "coords_list": coords.to(dev), "h_targets": h_targets}
        ex_loss, opt = self.cal_ex_loss(opt)
        return outputs, ex_loss, opt
