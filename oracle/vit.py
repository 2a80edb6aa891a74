"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the frozen ViT-B forward that the reference runs through HuggingFace
`transformers` (un-vendored, unpinned in the reference's requirement.txt:6; 5.5.0 is installed here) and of
the key-hook glue around it:

* `data/utils/feature_extractor.py:31-59`  — `backbone`: hook on `encoder.layer[-1].attention.attention.key`,
  DINOv1 called with `interpolate_pos_encoding=True`, keys reshaped `[B,1+HW,C] -> [B,C,H,W]` dropping CLS.
* `generate_pseudo_label.py:24-27,76-81`   — same hook + `outputs.attentions[-1]` (eager attention).
* transformers `models/dinov2/modeling_dinov2.py:38-117` (embeddings + bicubic pos-emb interpolation),
  `:153-235` (attention, scaling d_h^-0.5), `:348-387` (layer: LN -> attn -> LayerScale -> +res -> LN -> MLP(erf-GELU)
  -> LayerScale -> +res); `models/vit/modeling_vit.py` `ViTLayer` (same without LayerScale, eps 1e-12).

Weights use the HF state_dict key names so real checkpoints load unchanged.  Parity pin: `tools/make_golden.py`
loads `random_vit_state_dict` into the installed HF `Dinov2Model` / `ViTModel` and stores output slices in
`tests/golden/vit_*.npz`; `tests/test_oracle_golden.py::test_vit_matches_hf` checks this restatement against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VitSpec:
    kind: str            # 'dinov2' | 'dinov1'
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    mlp_dim: int = 3072
    patch: int = 14
    native_grid: int = 37    # sqrt(#position embeddings - 1)
    ln_eps: float = 1e-6
    layerscale: bool = True


DINOV2_B14 = VitSpec(kind="dinov2", patch=14, native_grid=37, ln_eps=1e-6, layerscale=True)
DINOV1_B8 = VitSpec(kind="dinov1", patch=8, native_grid=28, ln_eps=1e-12, layerscale=False)


def spec_for(kind: str) -> VitSpec:
    return DINOV2_B14 if "dinov2" in kind else DINOV1_B8


def layer_keys(spec: VitSpec, i: int) -> dict:
    p = f"encoder.layer.{i}."
    if spec.kind == "dinov2":
        return dict(ln1=p + "norm1", q=p + "attention.attention.query", k=p + "attention.attention.key",
                    v=p + "attention.attention.value", o=p + "attention.output.dense",
                    ls1=p + "layer_scale1.lambda1", ln2=p + "norm2", fc1=p + "mlp.fc1", fc2=p + "mlp.fc2",
                    ls2=p + "layer_scale2.lambda1")
    return dict(ln1=p + "layernorm_before", q=p + "attention.attention.query", k=p + "attention.attention.key",
                v=p + "attention.attention.value", o=p + "attention.output.dense", ls1=None,
                ln2=p + "layernorm_after", fc1=p + "intermediate.dense", fc2=p + "output.dense", ls2=None)


def random_vit_state_dict(spec: VitSpec, seed: int = 0, layerscale_init: float = 1.0) -> dict:
    """Seeded random-init weights (shared generator of the product package: it is input data, not algorithm)."""
    from ucod_dpl_b200.synth import random_vit_state_dict as gen
    return gen(spec, seed=seed, layerscale_init=layerscale_init)


def interpolate_pos_embedding(pos: torch.Tensor, native_grid: int, gh: int, gw: int) -> torch.Tensor:
    """modeling_dinov2.py:57-95 / ViTEmbeddings.interpolate_pos_encoding: bicubic, align_corners=False, fp32;
    identity when the grid is the native square grid."""
    pos = pos.reshape(-1, pos.shape[-1]).float()
    if gh == native_grid and gw == native_grid:
        return pos
    cls_pos, patch_pos = pos[:1], pos[1:]
    D = pos.shape[-1]
    pp = patch_pos.reshape(1, native_grid, native_grid, D).permute(0, 3, 1, 2)
    pp = F.interpolate(pp, size=(gh, gw), mode="bicubic", align_corners=False)
    pp = pp.permute(0, 2, 3, 1).reshape(-1, D)
    return torch.cat([cls_pos, pp], dim=0)


def _linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"].float(), sd[name + ".bias"].float())


@torch.no_grad()
def vit_forward(sd: dict, spec: VitSpec, images: torch.Tensor, want_attn: bool = False):
    """images: [B,3,H,W] fp32, already normalised.  Returns dict with
    key_tokens [B,1+HW,768] (output of the last layer's key projection — what the hook captures),
    cls_attn [B,heads,HW] (attentions[-1][:, :, 0, 1:]) if requested, last_hidden [B,1+HW,768] (pre final LN)."""
    B, _, Himg, Wimg = images.shape
    p, D, H = spec.patch, spec.hidden, spec.heads
    gh, gw = Himg // p, Wimg // p
    x = F.conv2d(images.float(), sd["embeddings.patch_embeddings.projection.weight"].float(),
                 sd["embeddings.patch_embeddings.projection.bias"].float(), stride=p)
    x = x.flatten(2).transpose(1, 2)
    cls = sd["embeddings.cls_token"].float().expand(B, -1, -1)
    x = torch.cat([cls, x], dim=1)
    x = x + interpolate_pos_embedding(sd["embeddings.position_embeddings"], spec.native_grid, gh, gw)[None]
    T = x.shape[1]
    key_tokens, cls_attn = None, None
    for i in range(spec.layers):
        k = layer_keys(spec, i)
        h = F.layer_norm(x, (D,), sd[k["ln1"] + ".weight"].float(), sd[k["ln1"] + ".bias"].float(), spec.ln_eps)
        q = _linear(h, sd, k["q"]).view(B, T, H, D // H).transpose(1, 2)
        kk_flat = _linear(h, sd, k["k"])
        kk = kk_flat.view(B, T, H, D // H).transpose(1, 2)
        v = _linear(h, sd, k["v"]).view(B, T, H, D // H).transpose(1, 2)
        if i == spec.layers - 1:
            key_tokens = kk_flat
        att = torch.softmax((q @ kk.transpose(-1, -2)) * (D // H) ** -0.5, dim=-1)
        if i == spec.layers - 1 and want_attn:
            cls_attn = att[:, :, 0, 1:].contiguous()
        ctx = (att @ v).transpose(1, 2).reshape(B, T, D)
        a = _linear(ctx, sd, k["o"])
        if spec.layerscale:
            a = a * sd[k["ls1"]].float()
        x = x + a
        h = F.layer_norm(x, (D,), sd[k["ln2"] + ".weight"].float(), sd[k["ln2"] + ".bias"].float(), spec.ln_eps)
        h = F.gelu(_linear(h, sd, k["fc1"]))
        h = _linear(h, sd, k["fc2"])
        if spec.layerscale:
            h = h * sd[k["ls2"]].float()
        x = x + h
    return {"key_tokens": key_tokens, "cls_attn": cls_attn, "last_hidden": x}


def keys_to_map(key_tokens: torch.Tensor) -> torch.Tensor:
    """feature_extractor.py:55-58: drop CLS, [B,HW,C] -> [B,C,H,W] (square grid assumed)."""
    B, L, C = key_tokens.shape
    g = int(math.isqrt(L - 1))
    return key_tokens[:, 1:, :].reshape(B, g, g, C).permute(0, 3, 1, 2)


def normalize_u8(images_u8: torch.Tensor) -> torch.Tensor:
    """torchvision ToTensor + Normalize (data/datasets/transforms.py:14-18): x/255, (x-mean)/std."""
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (images_u8.float() / 255.0 - mean) / std
