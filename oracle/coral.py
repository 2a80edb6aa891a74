"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the CORAL second stage (eval):

* `entropy_select`        — models/modules/ASR.py:13-51   (EntropySelector.forward / window_sets / get_position)
* `cross_attention_block` — models/modules/mlp.py:116-148 (CrossAttentionBlock; nn.MultiheadAttention 8 x 96)
* `csf_forward`           — models/modules/CSF.py:38-43
* `concate_windows`       — models/modules/HRE.py:18-39
* `gated_ensembler`       — models/modules/GE_pix_level.py:16-26
* `sparse_refiner_forward`— models/UDLR.py:77-86 (eval: cal_ex_loss returns 0)
* `prepare_validation_features`, `concate_preds`, `should_crop_center`, `center_pad`, `process_preds`
                          — engine/runner/loop_CORAL.py:62-96,168-204,206-258,313-341

Parity pin: tools/make_golden_coral.py runs the reference's own modules on the seeded inputs of
`ucod_dpl_b200.synth.synth_coral_inputs` and stores the outputs in tests/golden/coral.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

P = "HRE.CSF."


@torch.no_grad()
def entropy_select(input_features, h_inputs, preds, threshold: float, window_size: int):
    if torch.all((preds >= 0) & (preds <= 1)):
        probs = preds
    else:
        probs = preds.sigmoid()
    entropy = -probs * torch.log(probs.clamp(1e-5))
    scores = F.adaptive_avg_pool2d(entropy.float(), output_size=(window_size, window_size))
    mask = scores > threshold                                      # [B,1,w,w]
    flat = mask.flatten(1)                                         # [B,w*w]
    num = flat.sum(1)
    h_sel = h_inputs.flatten(0, 1)[mask.flatten()]
    l_rep = torch.repeat_interleave(input_features, num, dim=0)
    ys, xs = torch.meshgrid(torch.arange(window_size), torch.arange(window_size), indexing="ij")
    coords = torch.stack([ys, xs], dim=-1).flatten(0, 1)
    coords_list = torch.cat([coords[flat[b]] for b in range(mask.shape[0])], dim=0)
    return l_rep, h_sel, mask, coords_list, entropy, scores


@torch.no_grad()
def cross_attention_block(sd, query, context, heads: int = 8):
    """query [N,Q,C], context [N,K,C] -> [N,Q,C]; residual uses the un-normalised query (mlp.py:144)."""
    C = query.shape[-1]
    a = P + "attn."
    q = F.layer_norm(query, (C,), sd[a + "norm_q.weight"], sd[a + "norm_q.bias"], 1e-5)
    kv = F.layer_norm(context, (C,), sd[a + "norm_kv.weight"], sd[a + "norm_kv.bias"], 1e-5)
    w, b = sd[a + "attn.in_proj_weight"], sd[a + "attn.in_proj_bias"]
    Q = F.linear(q, w[:C], b[:C])
    K = F.linear(kv, w[C:2 * C], b[C:2 * C])
    V = F.linear(kv, w[2 * C:], b[2 * C:])
    N, Lq, _ = Q.shape
    d = C // heads
    Qh = Q.view(N, Lq, heads, d).transpose(1, 2)
    Kh = K.view(N, -1, heads, d).transpose(1, 2)
    Vh = V.view(N, -1, heads, d).transpose(1, 2)
    out = torch.empty_like(Qh)
    for n in range(N):                                             # per window: keeps the 3136^2 score block small
        att = torch.softmax(Qh[n] @ Kh[n].transpose(-1, -2) * d ** -0.5, dim=-1)
        out[n] = att @ Vh[n]
    ctx = out.transpose(1, 2).reshape(N, Lq, C)
    x = query + F.linear(ctx, sd[a + "attn.out_proj.weight"], sd[a + "attn.out_proj.bias"])
    h = F.layer_norm(x, (C,), sd[a + "norm_mlp.weight"], sd[a + "norm_mlp.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[a + "mlp.0.weight"], sd[a + "mlp.0.bias"])), sd[a + "mlp.2.weight"],
                 sd[a + "mlp.2.bias"])
    return x + h


@torch.no_grad()
def csf_forward(sd, l_inputs, h_inputs):
    """l_inputs, h_inputs [N,C,g,g] -> window logits [N,1,g,g]."""
    N, C, gh, gw = h_inputs.shape
    if N == 0:
        return torch.zeros(0, 1, gh, gw)
    out = cross_attention_block(sd, h_inputs.flatten(2).permute(0, 2, 1), l_inputs.flatten(2).permute(0, 2, 1))
    out = out.reshape(N, gh, gw, C).permute(0, 3, 1, 2)
    out = F.conv2d(out, sd[P + "depthwise_conv.weight"], sd[P + "depthwise_conv.bias"], padding=3, groups=C)
    return F.conv2d(out, sd[P + "mask_dec.weight"], sd[P + "mask_dec.bias"])


@torch.no_grad()
def concate_windows(windows, positions, mask, window_size: int):
    N, C, H, W = windows.shape
    B = mask.shape[0]
    full = torch.zeros(B, C, H * window_size, W * window_size)
    counter = torch.zeros(B, 1, H * window_size, W * window_size)
    num = mask.flatten(1).sum(1).tolist()
    last = 0
    for b in range(B):
        for i in range(last, last + num[b]):
            y, x = int(positions[i, 0]) * H, int(positions[i, 1]) * W
            full[b, :, y:y + H, x:x + W] += windows[i]
            counter[b, :, y:y + H, x:x + W] += 1.0
        last += num[b]
    return full / (counter + 1e-6)


@torch.no_grad()
def gated_ensembler(sd, l1, l2):
    _, _, h, w = l2.shape
    l1 = F.interpolate(l1, size=(h, w), mode="bilinear")
    p = torch.sigmoid(l1)
    fg_g = p.mean(dim=(1, 2, 3), keepdim=True)
    fg_l = F.avg_pool2d(p.float(), 19, padding=9, stride=1)
    en = -fg_l * torch.log(fg_l.clamp(1e-5))
    en = 1 - en / en.max()
    wgt = (en + fg_g) / 2
    y = l1 * wgt + l2 * (1 - wgt)
    y = F.conv2d(F.relu(F.conv2d(y, sd["GE.fuser.0.weight"], sd["GE.fuser.0.bias"])), sd["GE.fuser.2.weight"],
                 sd["GE.fuser.2.bias"])
    return y, wgt


@torch.no_grad()
def sparse_refiner_forward(sd, input_features, h_inputs, preds, threshold: float = 0.0015, window_size: int = 3):
    sd = {k: v.float() for k, v in sd.items()}
    l_rep, h_sel, mask, coords, entropy, scores = entropy_select(input_features, h_inputs, preds, threshold,
                                                                 window_size)
    window_preds = csf_forward(sd, l_rep, h_sel)
    h_preds = concate_windows(window_preds, coords, mask, window_size)
    outputs, ge_w = gated_ensembler(sd, preds, h_preds)
    return outputs, {"mask": mask, "entropy": entropy, "scores": scores, "h_preds": h_preds,
                     "window_preds": window_preds, "GE_w": ge_w, "preds": preds, "coords_list": coords}


# ---- engine/runner/loop_CORAL.py glue -----------------------------------------------------------------------
@torch.no_grad()
def concate_preds(preds):
    """[b,4,c,68,68] -> [b,c,102,102], overlapping 2x2 patches averaged (loop_CORAL.py:62-96)."""
    b, n, c, h, w = preds.shape
    full = torch.zeros(b, c, 102, 102)
    counter = torch.zeros(b, c, 102, 102)
    for i in range(2):
        for j in range(2):
            full[:, :, i * 34:i * 34 + 68, j * 34:j * 34 + 68] += preds[:, i * 2 + j]
            counter[:, :, i * 34:i * 34 + 68, j * 34:j * 34 + 68] += 1.0
    return full / (counter + 1e-6)


def should_crop_center(preds) -> bool:
    return bool(((preds > 0).sum() / (preds.shape[2] * preds.shape[3])) < 0.001)


def center_pad(x, fill_value: float = -10.0):
    b, c, h, w = x.shape
    out = torch.full((b, c, 2 * h, 2 * w), fill_value, dtype=x.dtype)
    out[:, :, h // 2:h // 2 + h, w // 2:w // 2 + w] = x
    return out


@torch.no_grad()
def process_preds(preds, size):
    h, w = size
    probs = preds if torch.all((preds >= 0) & (preds <= 1)) else preds.sigmoid()
    up = F.interpolate(probs, size=(h, w), mode="bilinear", align_corners=False)[..., :h, :w]
    return (up > 0.5).squeeze(0).float()
