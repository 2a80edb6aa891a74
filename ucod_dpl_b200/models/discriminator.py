"""`Discriminator` (APM scorer) with the reference's constructor, parameter names and `state_dict` layout
(models/discriminator.py:15-95): maskConv / convs.{0,1} are `ConvBlock`s = Sequential(Conv2d(bias=False),
BatchNorm2d, LeakyReLU(0.1)); `linear`; all parameters `requires_grad=False` at construction.
The forward runs in csrc/discriminator.cu.  `dis_use_features=True` (never used by a shipped config) is rejected."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..engine.registry import MODULE_REGISTRY


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, leaky_relu_slope=0.1, bias=False,
                 zero_init=False):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias),
            nn.BatchNorm2d(out_channels),
            nn.LeakyReLU(leaky_relu_slope, inplace=True))
        if zero_init:
            nn.init.constant_(self.layers[0].weight, 0)


@MODULE_REGISTRY.register()
class Discriminator(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.use_features = config.dis_use_features
        if self.use_features:
            raise NotImplementedError("dis_use_features=True is not used by any shipped config and is not implemented")
        self.maskConv = ConvBlock(1, 32, 3, 1, 1)
        indim = 32
        outdim = indim // 2
        self.convs = nn.ModuleList([ConvBlock(indim // (2 ** i), outdim // (2 ** i), kernel_size=3, stride=2,
                                              padding=1) for i in range(2)])
        self.linear = nn.Linear(outdim // 2 * ((config.feature_size + 3) // 4) ** 2, 1)
        for p in self.parameters():
            p.requires_grad = False

    def _tensors(self) -> dict:
        b1, b2, b3 = self.maskConv.layers, self.convs[0].layers, self.convs[1].layers
        t = {}
        for i, blk in enumerate((b1, b2, b3), start=1):
            t[f"conv{i}"] = blk[0].weight.data
            t[f"bn{i}_w"], t[f"bn{i}_b"] = blk[1].weight.data, blk[1].bias.data
            t[f"bn{i}_mean"], t[f"bn{i}_var"] = blk[1].running_mean, blk[1].running_var
        t["lin_w"], t["lin_b"] = self.linear.weight.data, self.linear.bias.data
        return t

    def forward(self, mask, feature=None):
        train = self.training
        out = ops.discriminator_forward(mask, self._tensors(), bn_train=train, update_running=train)
        if train:
            for blk in (self.maskConv.layers, self.convs[0].layers, self.convs[1].layers):
                blk[1].num_batches_tracked += 1
        return out


    def forward_calls(self, masks, calls: int):
        """`calls` consecutive `forward`s (masks [calls*B,1,fs,fs], call-major) in one set of launches; same results,
        BatchNorm statistics and buffer updates as calling `forward` on each slice in turn."""
        train = self.training
        out = ops.discriminator_forward_calls(masks, calls, self._tensors(), bn_train=train, update_running=train)
        if train:  # one fused launch for the three BatchNorm counters
            torch._foreach_add_([blk[1].num_batches_tracked for blk in
                                 (self.maskConv.layers, self.convs[0].layers, self.convs[1].layers)], calls)
        return out


def merge_pseudo_label(discriminator: Discriminator, pseudo_labels, p_teachers, p_students, features=None, *,
                       cur_epoch: int, max_epoch: int = 25, start_finetune: int = -5):
    """APM (`TrainLoop.merge_pseudo_label`, engine/runner/loop_UCOD_DPL.py:257-272).
    Returns (merged pseudo labels, dis_loss) like the reference; `.weight`/p_s/p_p are attached for logging."""
    s_mask, t_mask, p_mask = ops.apm_binarize(p_students, p_teachers, pseudo_labels)
    pair = getattr(s_mask, "pair", None)
    if pair is not None and hasattr(discriminator, "forward_calls"):
        p_both = discriminator.forward_calls(pair, 2)                            # student call, then pseudo-label call
        p_s, p_p = p_both[:s_mask.shape[0]], p_both[s_mask.shape[0]:]
    else:
        p_s = discriminator(s_mask, features)
        p_p = discriminator(p_mask, features)
    merged, weight, loss = ops.apm_merge(pseudo_labels, t_mask, p_s, p_p, cur_epoch / (max_epoch + start_finetune))
    merge_pseudo_label.last = {"weight": weight, "p_s": p_s, "p_p": p_p}
    return merged, loss
