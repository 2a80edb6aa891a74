"""Seeded synthetic inputs: no datasets or pretrained DINO weights exist offline.

Images: uint8, i.i.d. U{0..255} noise blended 50/50 with 2-4 random soft-edged ellipses so that features have
spatial structure.  Image i of a job always uses `torch.Generator().manual_seed(1234 + i)`, independent of how
the job is sharded over ranks.
"""
from __future__ import annotations

import math

import torch

from .vit import _layer_keys


def synth_image_u8(index: int, height: int, width: int) -> torch.Tensor:
    """[3,H,W] uint8 (CPU)."""
    g = torch.Generator().manual_seed(1234 + int(index))
    noise = torch.randint(0, 256, (3, height, width), generator=g, dtype=torch.int16).float()
    n_ell = int(torch.randint(2, 5, (1,), generator=g).item())
    yy = torch.arange(height, dtype=torch.float32).view(-1, 1)
    xx = torch.arange(width, dtype=torch.float32).view(1, -1)
    canvas = torch.full((3, height, width), 96.0)
    for _ in range(n_ell):
        r = torch.rand(8, generator=g)
        cy, cx = float(r[0]) * height, float(r[1]) * width
        ay = (0.05 + 0.25 * float(r[2])) * height
        ax = (0.05 + 0.25 * float(r[3])) * width
        th = float(r[4]) * math.pi
        c, s = math.cos(th), math.sin(th)
        u = ((xx - cx) * c + (yy - cy) * s) / ax
        v = (-(xx - cx) * s + (yy - cy) * c) / ay
        soft = torch.sigmoid((1.0 - torch.sqrt(u * u + v * v)) * 8.0)  # soft edge
        colour = (r[5:8] * 255.0).view(3, 1, 1)
        canvas = canvas * (1 - soft) + colour * soft
    img = 0.5 * noise + 0.5 * canvas
    return img.round().clamp_(0, 255).to(torch.uint8)


def synth_batch_u8(start: int, count: int, height: int, width: int) -> torch.Tensor:
    """[count,3,H,W] uint8 (CPU) — images start .. start+count-1."""
    return torch.stack([synth_image_u8(start + i, height, width) for i in range(count)], dim=0)


def random_vit_state_dict(spec, seed: int = 0, layerscale_init: float = 1.0) -> dict:
    """Deterministic random-init weights under HF key names (no pretrained DINO weights exist offline).

    Scales are chosen so activations stay O(1) through 12 layers (std 0.02 linears like HF's init, but
    non-trivial biases / LayerNorm affine / LayerScale so every fused epilogue term is exercised)."""
    g = torch.Generator().manual_seed(seed)
    D, Dm, p = spec.hidden, spec.mlp_dim, spec.patch

    def rn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd = {
        "embeddings.cls_token": rn(1, 1, D, std=1.0),
        "embeddings.position_embeddings": rn(1, spec.native_grid ** 2 + 1, D, std=0.2),
        "embeddings.patch_embeddings.projection.weight": rn(D, 3, p, p, std=0.05),
        "embeddings.patch_embeddings.projection.bias": rn(D, std=0.1),
        "layernorm.weight": 1.0 + rn(D, std=0.1),
        "layernorm.bias": rn(D, std=0.1),
    }
    if spec.kind == "dinov2":
        sd["embeddings.mask_token"] = torch.zeros(1, D)
    for i in range(spec.layers):
        k = _layer_keys(spec, i)
        for name in ("ln1", "ln2"):
            sd[k[name] + ".weight"] = 1.0 + rn(D, std=0.1)
            sd[k[name] + ".bias"] = rn(D, std=0.05)
        for name in ("q", "k", "v", "o"):
            sd[k[name] + ".weight"] = rn(D, D, std=0.04)
            sd[k[name] + ".bias"] = rn(D, std=0.05)
        sd[k["fc1"] + ".weight"] = rn(Dm, D, std=0.03)
        sd[k["fc1"] + ".bias"] = rn(Dm, std=0.05)
        sd[k["fc2"] + ".weight"] = rn(D, Dm, std=0.02)
        sd[k["fc2"] + ".bias"] = rn(D, std=0.05)
        if spec.layerscale:
            sd[k["ls1"]] = layerscale_init * (1.0 + rn(D, std=0.1))
            sd[k["ls2"]] = layerscale_init * (1.0 + rn(D, std=0.1))
    return sd


def random_refiner_state_dict(seed: int = 0, dim: int = 768) -> dict:
    """Seeded random weights for the CORAL `SparseRefiner` under the reference's state_dict names
    (weights/CORAL_dinov{1,2}.safetensors are not shipped — .MISSING_LARGE_BLOBS).  Scales keep activations O(1)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    p = "HRE.CSF."
    sd = {}
    for n in ("norm_q", "norm_kv", "norm_mlp"):
        sd[f"{p}attn.{n}.weight"] = 1.0 + rn(dim, std=0.1)
        sd[f"{p}attn.{n}.bias"] = rn(dim, std=0.05)
    sd[p + "attn.attn.in_proj_weight"] = rn(3 * dim, dim, std=0.04)
    sd[p + "attn.attn.in_proj_bias"] = rn(3 * dim, std=0.05)
    sd[p + "attn.attn.out_proj.weight"] = rn(dim, dim, std=0.04)
    sd[p + "attn.attn.out_proj.bias"] = rn(dim, std=0.05)
    sd[p + "attn.mlp.0.weight"] = rn(4 * dim, dim, std=0.03)
    sd[p + "attn.mlp.0.bias"] = rn(4 * dim, std=0.05)
    sd[p + "attn.mlp.2.weight"] = rn(dim, 4 * dim, std=0.02)
    sd[p + "attn.mlp.2.bias"] = rn(dim, std=0.05)
    sd[p + "depthwise_conv.weight"] = rn(dim, 1, 7, 7, std=0.1)
    sd[p + "depthwise_conv.bias"] = rn(dim, std=0.05)
    sd[p + "mask_dec.weight"] = rn(1, dim, 1, 1, std=0.05)
    sd[p + "mask_dec.bias"] = rn(1, std=0.1)
    sd["GE.alpha"] = torch.tensor(0.5)
    sd["GE.fuser.0.weight"] = rn(64, 1, 1, 1, std=0.5)
    sd["GE.fuser.0.bias"] = rn(64, std=0.3)
    sd["GE.fuser.2.weight"] = rn(1, 64, 1, 1, std=0.3)
    sd["GE.fuser.2.bias"] = rn(1, std=0.1)
    return sd


def synth_coral_inputs(seed: int, batch: int = 1, grid: int = 56, dim: int = 768, windows: int = 3,
                       uncertain=((0, 1), (1, 1), (2, 0))):
    """Seeded CORAL refiner inputs: l features [B,dim,g,g], h features [B,w*w,dim,g,g] and coarse logits [B,1,g,g]
    that are confident (|logit| = 12) everywhere except inside the listed (row, col) windows, where they are
    uncertain (|logit| < 1.5) — so the entropy selector picks exactly those windows."""
    g = torch.Generator().manual_seed(seed)
    l = torch.randn(batch, dim, grid, grid, generator=g)
    h = torch.randn(batch, windows * windows, dim, grid, grid, generator=g)
    preds = torch.full((batch, 1, grid, grid), -12.0)
    preds[:, :, : grid // 2, : grid // 3] = 12.0
    edges = [(i * grid) // windows for i in range(windows + 1)]
    for (r, c) in uncertain:
        y0, y1, x0, x1 = edges[r] + 1, edges[r + 1] - 1, edges[c] + 1, edges[c + 1] - 1
        preds[:, :, y0:y1, x0:x1] = 1.5 * torch.randn(batch, 1, y1 - y0, x1 - x0, generator=g).clamp(-1, 1)
    return l, h, preds


def planted_object_logits(index: int, feature_size: int = 68, objects: int = 2) -> torch.Tensor:
    """[1,fs,fs] fp32 first-look logits with `objects` well-separated small blobs (SURVEY.md 8(d) "LT" row: a random-init
    backbone almost never segments small objects, so the Look-Twice benchmark plants them at `process_preds`' input).
    Each blob covers 1.5-5 % of the frame after `sigmoid > 0.5` — above the 1 % box gate, far below look_twice_th —
    so `process_preds` returns exactly `objects` boxes for every image."""
    g = torch.Generator().manual_seed(4321 + int(index))
    fs = feature_size
    yy = torch.arange(fs, dtype=torch.float32).view(-1, 1)
    xx = torch.arange(fs, dtype=torch.float32).view(1, -1)
    out = torch.full((fs, fs), -4.0)
    placed = []
    while len(placed) < objects:
        r = torch.rand(3, generator=g)
        rad = (0.10 + 0.07 * float(r[2])) * fs            # logit > 0 inside rad / sqrt(2): area pi rad^2 / 2
        cy = rad + float(r[0]) * (fs - 2 * rad)
        cx = rad + float(r[1]) * (fs - 2 * rad)
        if any((cy - py) ** 2 + (cx - px) ** 2 < (rad + pr + 3.0) ** 2 for py, px, pr in placed):
            continue
        placed.append((cy, cx, rad))
        out = torch.maximum(out, 4.0 * (1.0 - 2.0 * (((yy - cy) / rad) ** 2 + ((xx - cx) / rad) ** 2)))
    return out[None]


def planted_object_batch(start: int, count: int, feature_size: int = 68, objects: int = 2) -> torch.Tensor:
    """[count,1,fs,fs] fp32 (CPU)."""
    return torch.stack([planted_object_logits(start + i, feature_size, objects) for i in range(count)], dim=0)
