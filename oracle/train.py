"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of one first-stage training step (engine/runner/loop_UCOD_DPL.py:148-191 `_process_batch` +
`update_ema_decoder`, optimiser from engine/runner/runner.py:276-305: AdamW(lr0) with torch defaults, StepLR stepped
every iteration), using torch autograd for the gradients.

Parity pin: tools/make_golden_train.py drives the reference's own `TrainLoop._process_batch` (unbound, CPU) on
seeded inputs; tests/golden/train.npz holds the loss, the gradients and the updated student / EMA parameters.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import decoder as odec

PARAM_ORDER = ["learnable_embedding", "decoupling.weight", "decoupling.bias", "conv_out_fg.weight",
               "conv_out_fg.bias", "conv_out_bg.weight", "conv_out_bg.bias"]  # nn.Module.parameters() order


def student_forward(p: dict, x: torch.Tensor):
    """Differentiable RevDecoder.forward (models/modules/DBA.py:31-59) on a dict of student parameters."""
    B, _, H, W = x.shape
    d = F.conv2d(x, p["decoupling.weight"], p["decoupling.bias"])
    d1, d2 = torch.chunk(d, 2, dim=1)
    emb = p["learnable_embedding"]
    f1 = F.normalize(d1.reshape(B, 64, -1).permute(0, 2, 1) * emb[0], p=2, dim=1)
    f2 = F.normalize(d2.reshape(B, 64, -1).permute(0, 2, 1) * emb[1], p=2, dim=1)
    dot = torch.bmm(f1, f2.transpose(1, 2))
    ortho = (dot * (1 - torch.eye(f1.size(1)))).pow(2).mean()
    f1 = f1.reshape(B, H, W, 64).permute(0, 3, 1, 2)
    f2 = f2.reshape(B, H, W, 64).permute(0, 3, 1, 2)
    fg = F.conv2d(torch.sigmoid(f1 * d1) + d1, p["conv_out_fg.weight"], p["conv_out_fg.bias"])
    bg = F.conv2d(torch.sigmoid(f2 * d2) + d2, p["conv_out_bg.weight"], p["conv_out_bg.bias"])
    return fg, bg, ortho


def train_step(sd: dict, dis_sd: dict, state: dict, features: torch.Tensor, pseudo_labels: torch.Tensor, *,
               cur_epoch: int, global_step: int, lr: float, feature_size: int = 68, finetune: bool = False,
               ema_weight: float = 0.99, max_epoch: int = 25, start_finetune: int = -5, betas=(0.9, 0.999),
               eps: float = 1e-8, weight_decay: float = 0.01, merged_override=None, dis_loss_override=None):
    """One `_process_batch`.  sd: full baseline state_dict (student + EMA), updated in place; state: dict with
    'm', 'v' (dicts of tensors) and 't' (AdamW step count), updated in place.
    Returns dict(loss, grads, merged, dis_loss, ortho)."""
    fs = (feature_size, feature_size)
    x = F.interpolate(features.float(), size=fs, mode="bilinear")
    pl = F.interpolate(pseudo_labels.float(), size=fs, mode="bilinear")
    with torch.no_grad():
        teacher = odec.baseline_forward(sd, x, ema=True)
    p = {k: sd["decoder." + k].clone().float().requires_grad_(True) for k in PARAM_ORDER}
    fg, bg, ortho = student_forward(p, x)
    merged, dis_loss, w, _, _ = odec.apm_merge(dis_sd, pl, teacher, fg.detach(), cur_epoch, max_epoch, start_finetune)
    if merged_override is not None:  # stage isolation for the GPU parity test: APM outputs injected
        merged, dis_loss = merged_override.float(), dis_loss_override.float()
    flat_t = merged.permute(0, 2, 3, 1).reshape(-1, 1)
    loss = F.binary_cross_entropy_with_logits(fg.permute(0, 2, 3, 1).reshape(-1, 1), flat_t)
    if not finetune:
        loss = loss - dis_loss
    loss = loss + F.binary_cross_entropy_with_logits(bg.permute(0, 2, 3, 1).reshape(-1, 1), 1 - flat_t)
    loss = loss + ortho
    loss.backward()
    grads = {k: p[k].grad.detach().clone() for k in PARAM_ORDER}
    # AdamW (torch defaults) -----------------------------------------------------------------------------
    state["t"] += 1
    t = state["t"]
    b1, b2 = betas
    with torch.no_grad():
        for k in PARAM_ORDER:
            w_ = sd["decoder." + k]
            g = grads[k]
            w_.mul_(1 - lr * weight_decay)
            state["m"][k].mul_(b1).add_(g, alpha=1 - b1)
            state["v"][k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (state["v"][k].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            w_.addcdiv_(state["m"][k], denom, value=-lr / (1 - b1 ** t))
        # EMA (loop_UCOD_DPL.py:186-191)
        alpha = min(1 - 1 / (global_step + 1), ema_weight)
        for k in PARAM_ORDER:
            sd["decoder_ema." + k].mul_(alpha).add_(sd["decoder." + k], alpha=1 - alpha)
    return {"loss": loss.detach(), "grads": grads, "merged": merged, "dis_loss": dis_loss, "ortho": ortho.detach(),
            "weight": w}


def new_state(sd: dict) -> dict:
    return {"t": 0, "m": {k: torch.zeros_like(sd["decoder." + k], dtype=torch.float32) for k in PARAM_ORDER},
            "v": {k: torch.zeros_like(sd["decoder." + k], dtype=torch.float32) for k in PARAM_ORDER}}


def step_lr(lr0: float, optimizer_steps: int, step_size: int = 25, gamma: float = 0.95) -> float:
    """StepLR stepped once per iteration (loop_UCOD_DPL.py:178): lr used for optimiser step number `optimizer_steps`
    (0-based)."""
    return lr0 * gamma ** (optimizer_steps // step_size)


# ---- discriminator epoch step (engine/runner/loop_UCOD_DPL.py:230-255) --------------------------------------
DIS_PARAM_ORDER = ["maskConv.layers.0.weight", "maskConv.layers.1.weight", "maskConv.layers.1.bias",
                   "convs.0.layers.0.weight", "convs.0.layers.1.weight", "convs.0.layers.1.bias",
                   "convs.1.layers.0.weight", "convs.1.layers.1.weight", "convs.1.layers.1.bias",
                   "linear.weight", "linear.bias"]


def _disc_forward_train(p: dict, sd: dict, mask: torch.Tensor, momentum: float = 0.1):
    """Differentiable Discriminator.forward in train mode; updates the running statistics in `sd` in place."""
    y = mask
    for prefix, stride in (("maskConv.", 1), ("convs.0.", 2), ("convs.1.", 2)):
        y = F.conv2d(y, p[prefix + "layers.0.weight"], None, stride=stride, padding=1)
        y = F.batch_norm(y, sd[prefix + "layers.1.running_mean"], sd[prefix + "layers.1.running_var"],
                         p[prefix + "layers.1.weight"], p[prefix + "layers.1.bias"], training=True, momentum=momentum,
                         eps=1e-5)
        y = F.leaky_relu(y, 0.1)
    return torch.sigmoid(F.linear(torch.flatten(y, 1), p["linear.weight"], p["linear.bias"]))


def discriminator_step(dis_sd: dict, state: dict, pseudo_mask: torch.Tensor, student_mask: torch.Tensor, lr: float,
                       betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01):
    """One iteration of Discriminator_epoch given the two binarised inputs: loss = BCE(cat(D(student), D(pseudo)),
    cat(0, 1)); AdamW(dis_lr0) step.  dis_sd (incl. BatchNorm running stats) and state are updated in place."""
    p = {k: dis_sd[k].clone().float().requires_grad_(True) for k in DIS_PARAM_ORDER}
    probs_pseudo = _disc_forward_train(p, dis_sd, pseudo_mask.float())
    probs_student = _disc_forward_train(p, dis_sd, student_mask.float())
    B = pseudo_mask.shape[0]
    label = torch.cat((torch.zeros(B), torch.ones(B))).unsqueeze(-1)
    loss = F.binary_cross_entropy(torch.cat((probs_student, probs_pseudo), dim=0), label)
    loss.backward()
    grads = {k: p[k].grad.detach().clone() for k in DIS_PARAM_ORDER}
    state["t"] += 1
    t = state["t"]
    b1, b2 = betas
    with torch.no_grad():
        for k in DIS_PARAM_ORDER:
            w_, g = dis_sd[k], grads[k]
            w_.mul_(1 - lr * weight_decay)
            state["m"][k].mul_(b1).add_(g, alpha=1 - b1)
            state["v"][k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (state["v"][k].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
            w_.addcdiv_(state["m"][k], denom, value=-lr / (1 - b1 ** t))
    return {"loss": loss.detach(), "grads": grads, "probs_pseudo": probs_pseudo.detach(),
            "probs_student": probs_student.detach()}


def new_dis_state(dis_sd: dict) -> dict:
    return {"t": 0, "m": {k: torch.zeros_like(dis_sd[k], dtype=torch.float32) for k in DIS_PARAM_ORDER},
            "v": {k: torch.zeros_like(dis_sd[k], dtype=torch.float32) for k in DIS_PARAM_ORDER}}
