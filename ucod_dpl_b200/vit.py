"""Host side of the frozen ViT-B key extractor: weight packing + ctypes calls into `ucod_vit_*`.

PyTorch is used for device memory and streams only; every kernel is in `csrc/` (vit.cu, gemm.cu, attention.cu).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import _lib


@dataclass(frozen=True)
class VitSpec:
    kind: str
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    mlp_dim: int = 3072
    patch: int = 14
    native_grid: int = 37
    ln_eps: float = 1e-6
    layerscale: bool = True


DINOV2_B14 = VitSpec(kind="dinov2", patch=14, native_grid=37, ln_eps=1e-6, layerscale=True)
DINOV1_B8 = VitSpec(kind="dinov1", patch=8, native_grid=28, ln_eps=1e-12, layerscale=False)


def spec_for(kind: str) -> VitSpec:
    return DINOV2_B14 if "dinov2" in kind else DINOV1_B8


class _Cfg(ctypes.Structure):
    _fields_ = [("hidden", ctypes.c_int), ("layers", ctypes.c_int), ("heads", ctypes.c_int),
                ("mlp_dim", ctypes.c_int), ("patch", ctypes.c_int), ("patch_kpad", ctypes.c_int),
                ("ln_eps", ctypes.c_float)]


class _Layer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "ln1_w", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_w", "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


def _layer_keys(spec: VitSpec, i: int) -> dict:
    p = f"encoder.layer.{i}."
    if spec.kind == "dinov2":
        return dict(ln1=p + "norm1", q=p + "attention.attention.query", k=p + "attention.attention.key",
                    v=p + "attention.attention.value", o=p + "attention.output.dense",
                    ls1=p + "layer_scale1.lambda1", ln2=p + "norm2", fc1=p + "mlp.fc1", fc2=p + "mlp.fc2",
                    ls2=p + "layer_scale2.lambda1")
    return dict(ln1=p + "layernorm_before", q=p + "attention.attention.query", k=p + "attention.attention.key",
                v=p + "attention.attention.value", o=p + "attention.output.dense", ls1=None,
                ln2=p + "layernorm_after", fc1=p + "intermediate.dense", fc2=p + "output.dense", ls2=None)


class VitKeyExtractor:
    """Last-layer key tokens (and optionally the CLS attention row) of a frozen ViT-B.

    `state_dict` uses HuggingFace key names (Dinov2Model / ViTModel).  Weights are converted once:
    matrices to bf16 (Q/K/V concatenated to one [2304,768]), vectors fp32.
    """

    def __init__(self, state_dict: dict, spec: VitSpec, device="cuda"):
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.UcodError("VitKeyExtractor needs a CUDA device; there is no CPU fallback")
        _lib.load()
        sd = state_dict
        dev = self.device
        D, p = spec.hidden, spec.patch
        self._keep = []  # device tensors referenced by the C handle

        def f32(t):
            t = t.detach().to(dev, torch.float32).contiguous()
            self._keep.append(t)
            return t

        def bf16(t):
            t = t.detach().to(dev, torch.float32).to(torch.bfloat16).contiguous()
            self._keep.append(t)
            return t

        kvalid = 3 * p * p
        self.kpad = (kvalid + 63) // 64 * 64
        pw = sd["embeddings.patch_embeddings.projection.weight"].detach().float().reshape(D, kvalid)
        pw = F.pad(pw, (0, self.kpad - kvalid))
        self.patch_w = bf16(pw)
        self.patch_b = f32(sd["embeddings.patch_embeddings.projection.bias"])
        self.cls = f32(sd["embeddings.cls_token"].reshape(D))
        self.pos_native = sd["embeddings.position_embeddings"].detach().float().reshape(-1, D).cpu()
        self._pos_cache: dict = {}

        layers = (_Layer * spec.layers)()
        for i in range(spec.layers):
            k = _layer_keys(spec, i)
            wqkv = torch.cat([sd[k["q"] + ".weight"], sd[k["k"] + ".weight"], sd[k["v"] + ".weight"]], dim=0)
            bqkv = torch.cat([sd[k["q"] + ".bias"], sd[k["k"] + ".bias"], sd[k["v"] + ".bias"]], dim=0)
            L = layers[i]
            L.ln1_w = f32(sd[k["ln1"] + ".weight"]).data_ptr()
            L.ln1_b = f32(sd[k["ln1"] + ".bias"]).data_ptr()
            L.w_qkv = bf16(wqkv).data_ptr()
            L.b_qkv = f32(bqkv).data_ptr()
            # LayerScale (DINOv2) is folded into the projection that precedes it: ls * (W x + b) = (ls W) x + ls b
            ls1 = sd[k["ls1"]].detach().float() if k["ls1"] else None
            ls2 = sd[k["ls2"]].detach().float() if k["ls2"] else None
            w_o, b_o = sd[k["o"] + ".weight"].detach().float(), sd[k["o"] + ".bias"].detach().float()
            w_2, b_2 = sd[k["fc2"] + ".weight"].detach().float(), sd[k["fc2"] + ".bias"].detach().float()
            if ls1 is not None:
                w_o, b_o = w_o * ls1[:, None], b_o * ls1
            if ls2 is not None:
                w_2, b_2 = w_2 * ls2[:, None], b_2 * ls2
            L.w_o = bf16(w_o).data_ptr()
            L.b_o = f32(b_o).data_ptr()
            L.ln2_w = f32(sd[k["ln2"] + ".weight"]).data_ptr()
            L.ln2_b = f32(sd[k["ln2"] + ".bias"]).data_ptr()
            L.w_fc1 = bf16(sd[k["fc1"] + ".weight"]).data_ptr()
            L.b_fc1 = f32(sd[k["fc1"] + ".bias"]).data_ptr()
            L.w_fc2 = bf16(w_2).data_ptr()
            L.b_fc2 = f32(b_2).data_ptr()
        cfg = _Cfg(spec.hidden, spec.layers, spec.heads, spec.mlp_dim, spec.patch, self.kpad, spec.ln_eps)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("ucod_vit_create", ctypes.byref(self._handle), ctypes.byref(cfg), _lib.ptr(self.patch_w),
                      _lib.ptr(self.patch_b), _lib.ptr(self.cls), layers)
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.load().ucod_vit_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # -- position embedding for a token grid (bicubic interpolation is a one-off per resolution) --
    def pos_embedding(self, gh: int, gw: int) -> torch.Tensor:
        key = (gh, gw)
        if key not in self._pos_cache:
            g = self.spec.native_grid
            pos = self.pos_native
            if not (gh == g and gw == g):
                D = pos.shape[-1]
                pp = pos[1:].reshape(1, g, g, D).permute(0, 3, 1, 2)
                pp = F.interpolate(pp, size=(gh, gw), mode="bicubic", align_corners=False)
                pos = torch.cat([pos[:1], pp.permute(0, 2, 3, 1).reshape(-1, D)], dim=0)
            self._pos_cache[key] = pos.to(self.device, torch.float32).contiguous()
        return self._pos_cache[key]

    def workspace(self, B: int, H: int, W: int) -> torch.Tensor:
        need = ctypes.c_uint64(0)
        _lib.call("ucod_vit_workspace_bytes", self._handle, B, H, W, ctypes.byref(need))
        n = int(need.value)
        if self._ws is None or self._ws.numel() < n:
            self._ws = None
            self._ws = torch.empty(n + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def keys(self, images: torch.Tensor, *, want_f32: bool = True, want_bf16: bool = False,
             want_cls_attn: bool = False, keep_cls: bool = False, count_dev: torch.Tensor | None = None):
        """images: [B,3,H,W] fp32 (normalised) or uint8 (raw RGB) CUDA tensor.

        Returns (keys_f32 | None, keys_bf16 | None, cls_attn | None); keys are token-major [B, P(+1), 768].
        count_dev (optional, int32 device scalar): only the first `count_dev` images are processed — the count is read
        by the kernels on the device (`ucod_vit_keys_dyn`), B is the capacity; rows of later images are undefined."""
        _lib.require_cuda(images)
        if images.dim() != 4 or images.shape[1] != 3:
            raise _lib.UcodError(f"expected images [B,3,H,W], got {tuple(images.shape)}")
        if images.dtype == torch.uint8:
            dt = 1
        elif images.dtype == torch.float32:
            dt = 0
        else:
            images, dt = images.float(), 0
        images = images.contiguous()
        B, _, H, W = images.shape
        p = self.spec.patch
        gh, gw = H // p, W // p
        P = gh * gw
        pos = self.pos_embedding(gh, gw)
        ws = self.workspace(B, H, W)
        off = (-ws.data_ptr()) % 1024
        rows = P + (1 if keep_cls else 0)
        D = self.spec.hidden
        k32 = torch.empty(B, rows, D, device=self.device, dtype=torch.float32) if want_f32 else None
        k16 = torch.empty(B, rows, D, device=self.device, dtype=torch.bfloat16) if want_bf16 else None
        att = torch.empty(B, self.spec.heads, P, device=self.device, dtype=torch.float32) if want_cls_attn else None
        with torch.cuda.device(self.device):
            if count_dev is None:
                _lib.call("ucod_vit_keys", self._handle, _lib.ptr(images), dt, B, H, W, _lib.ptr(pos),
                          ctypes.c_void_p(ws.data_ptr() + off), ctypes.c_uint64(ws.numel() - off), _lib.ptr(k32),
                          _lib.ptr(k16), _lib.ptr(att), 1 if keep_cls else 0, _lib.stream_ptr(self.device))
            else:
                _lib.require_cuda(count_dev)
                if count_dev.dtype != torch.int32:
                    raise _lib.UcodError("count_dev must be an int32 device tensor")
                _lib.call("ucod_vit_keys_dyn", self._handle, _lib.ptr(images), dt, B, _lib.ptr(count_dev), H, W,
                          _lib.ptr(pos), ctypes.c_void_p(ws.data_ptr() + off), ctypes.c_uint64(ws.numel() - off),
                          _lib.ptr(k32), _lib.ptr(k16), _lib.ptr(att), 1 if keep_cls else 0,
                          _lib.stream_ptr(self.device))
        return k32, k16, att
