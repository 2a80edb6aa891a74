"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatements of the reference's decoder-side modules, written functionally over a state_dict with
the reference's key names (weights/UCOD_DPL_dinov{1,2}.safetensors load unchanged):

* `rev_decoder_forward`   — models/modules/DBA.py:31-59 (`RevDecoder.forward`) and :25-29 (`calc_orthogonal_loss`)
* `baseline_forward`      — models/uscod.py:16-22
* `discriminator_forward` — models/discriminator.py:15-70,86-95 (BatchNorm in train OR eval mode, LeakyReLU 0.1)
* `apm_merge`             — engine/runner/loop_UCOD_DPL.py:257-272 (`TrainLoop.merge_pseudo_label`)
* `upsample_bilinear`     — F.interpolate(mode='bilinear', align_corners=False) (loop_UCOD_DPL.py:153,305,356)

Parity pin: tools/make_golden.py runs the reference's own modules (imported from /root/reference) on seeded
inputs with the shipped checkpoints and stores the outputs in tests/golden/decoder_*.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def upsample_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)


@torch.no_grad()
def rev_decoder_forward(sd: dict, prefix: str, x: torch.Tensor, want_ortho: bool = True):
    """x: [B,dim,H,W] fp32.  Returns (fg [B,1,H,W], bg [B,1,H,W], ortho scalar or None)."""
    if isinstance(x, list):
        x = x[-1]
    B, _, H, W = x.shape
    E = 64
    d = F.conv2d(x.float(), sd[prefix + "decoupling.weight"].float(), sd[prefix + "decoupling.bias"].float())
    d1, d2 = torch.chunk(d, 2, dim=1)
    emb = sd[prefix + "learnable_embedding"].float()
    f1 = d1.reshape(B, E, -1).permute(0, 2, 1)          # [B,HW,64]
    f2 = d2.reshape(B, E, -1).permute(0, 2, 1)
    f1 = F.normalize(f1 * emb[0], p=2, dim=1)            # normalised over HW (dim=1), per channel  (DBA.py:40)
    f2 = F.normalize(f2 * emb[1], p=2, dim=1)
    ortho = None
    if want_ortho:
        dot = torch.bmm(f1, f2.transpose(1, 2))          # [B,HW,HW]  (DBA.py:26)
        eye = torch.eye(f1.size(1))
        ortho = (dot * (1 - eye)).pow(2).mean()
    f1 = f1.reshape(B, H, W, E).permute(0, 3, 1, 2)
    f2 = f2.reshape(B, H, W, E).permute(0, 3, 1, 2)
    a1 = torch.sigmoid(f1 * d1) + d1
    a2 = torch.sigmoid(f2 * d2) + d2
    fg = F.conv2d(a1, sd[prefix + "conv_out_fg.weight"].float(), sd[prefix + "conv_out_fg.bias"].float())
    bg = F.conv2d(a2, sd[prefix + "conv_out_bg.weight"].float(), sd[prefix + "conv_out_bg.bias"].float())
    return fg, bg, ortho


def baseline_forward(sd: dict, x: torch.Tensor, ema: bool = False, want_ortho: bool = True):
    """models/uscod.py:16-22 — student returns (fg, bg, extra_loss); EMA returns fg only."""
    if ema:
        fg, _, _ = rev_decoder_forward(sd, "decoder_ema.", x, want_ortho=False)
        return fg
    return rev_decoder_forward(sd, "decoder.", x, want_ortho=want_ortho)


def _conv_block(x, sd, prefix, stride, train_bn, eps=1e-5):
    """ConvBlock: Conv2d(k=3,pad=1,bias=False) -> BatchNorm2d -> LeakyReLU(0.1)  (discriminator.py:28-46)."""
    y = F.conv2d(x, sd[prefix + "layers.0.weight"].float(), None, stride=stride, padding=1)
    w, b = sd[prefix + "layers.1.weight"].float(), sd[prefix + "layers.1.bias"].float()
    if train_bn:
        y = F.batch_norm(y, None, None, w, b, training=True, eps=eps)
    else:
        y = F.batch_norm(y, sd[prefix + "layers.1.running_mean"].float(), sd[prefix + "layers.1.running_var"].float(),
                         w, b, training=False, eps=eps)
    return F.leaky_relu(y, 0.1)


@torch.no_grad()
def discriminator_forward(sd: dict, mask: torch.Tensor, train_bn: bool = True) -> torch.Tensor:
    """mask [B,1,fs,fs] -> [B,1] in (0,1); dis_use_features=False in every shipped config, so `feature` is unused.
    The reference never calls .eval() on the discriminator, so APM runs BatchNorm with batch statistics."""
    y = _conv_block(mask.float(), sd, "maskConv.", 1, train_bn)
    y = _conv_block(y, sd, "convs.0.", 2, train_bn)
    y = _conv_block(y, sd, "convs.1.", 2, train_bn)
    y = torch.flatten(y, 1)
    return torch.sigmoid(F.linear(y, sd["linear.weight"].float(), sd["linear.bias"].float()))


def random_discriminator_state_dict(feature_size: int = 68, seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(seed)

    def rn(*s, std=0.2):
        return torch.randn(*s, generator=g) * std

    sd = {}
    for name, cin, cout in (("maskConv.", 1, 32), ("convs.0.", 32, 16), ("convs.1.", 16, 8)):
        sd[name + "layers.0.weight"] = rn(cout, cin, 3, 3)
        sd[name + "layers.1.weight"] = 1.0 + rn(cout, std=0.1)
        sd[name + "layers.1.bias"] = rn(cout, std=0.1)
        sd[name + "layers.1.running_mean"] = rn(cout, std=0.1)
        sd[name + "layers.1.running_var"] = 1.0 + rn(cout, std=0.1).abs()
        sd[name + "layers.1.num_batches_tracked"] = torch.tensor(0)
    n = 8 * ((feature_size + 3) // 4) ** 2
    sd["linear.weight"] = rn(1, n, std=0.05)
    sd["linear.bias"] = rn(1, std=0.1)
    return sd


@torch.no_grad()
def apm_merge(dis_sd: dict, pseudo_labels: torch.Tensor, teacher_logits: torch.Tensor, student_logits: torch.Tensor,
              cur_epoch: int, max_epoch: int = 25, start_finetune: int = -5, train_bn: bool = True):
    """loop_UCOD_DPL.py:257-272.  Returns (merged [B,1,H,W], dis_loss scalar, weight [B,1], p_s, p_p)."""
    t = (torch.sigmoid(teacher_logits) > 0.5).float()
    s = (torch.sigmoid(student_logits) > 0.5).float()
    p_s = discriminator_forward(dis_sd, s, train_bn)
    p_p = discriminator_forward(dis_sd, (pseudo_labels > 0.5).float(), train_bn)
    w = 0.5 * (1 + torch.cos(torch.abs(p_s - p_p) * math.pi)) + cur_epoch / (max_epoch + start_finetune)
    w = torch.clamp(w, 0, 1)
    w4 = w.unsqueeze(-1).unsqueeze(-1)
    loss = F.binary_cross_entropy(p_s, torch.zeros_like(p_s))
    return pseudo_labels * (1 - w4) + t * w4, loss, w, p_s, p_p
