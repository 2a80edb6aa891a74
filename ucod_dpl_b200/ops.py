"""Thin Python wrappers over the C-ABI kernels (one function per `ucod_*` entry point).

All functions take CUDA tensors, allocate outputs with torch (device memory plumbing only) and launch on the
current CUDA stream.  No CPU fallback exists: non-CUDA inputs raise `UcodError`.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import UcodError, c_float, c_int, ptr, stream_ptr

_u64 = ctypes.c_uint64
_i64 = ctypes.c_int64


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)


def _aligned(ws: torch.Tensor):
    off = (-ws.data_ptr()) % 1024
    return ctypes.c_void_p(ws.data_ptr() + off), _u64(ws.numel() - off)


# ------------------------------------------------------------------------------------------------
def features_to_tokens_bf16(features: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 (any strides for C / HW as long as HW is jointly strided) -> [B,H*W,C] bf16."""
    _lib.require_cuda(features)
    if features.dtype != torch.float32:
        features = features.float()
    B, C, H, W = features.shape
    if features.stride(2) != W * features.stride(3):
        features = features.contiguous()
    out = torch.empty(B, H * W, C, device=features.device, dtype=torch.bfloat16)
    with torch.cuda.device(features.device):
        _lib.call("ucod_features_to_tokens_bf16", ptr(features), ptr(out), B, C, H * W, _i64(features.stride(0)),
                  _i64(features.stride(1)), _i64(features.stride(3)), stream_ptr(features.device))
    return out


def decoder_forward(keys_bf16: torch.Tensor, grid_in, grid_out, w_dec_bf16, b_dec, emb, w_fg, b_fg, w_bg, b_bg, *,
                    want_bg: bool = True, want_ortho: bool = False):
    """keys_bf16 [B, gin_h*gin_w, dim] -> (fg [B,1,oh,ow], bg | None, ortho scalar tensor | None)."""
    _lib.require_cuda(keys_bf16)
    if keys_bf16.dtype != torch.bfloat16 or not keys_bf16.is_contiguous():
        raise UcodError("decoder_forward expects contiguous bf16 token-major keys")
    B, P, dim = keys_bf16.shape
    gh, gw = grid_in
    oh, ow = grid_out
    if gh * gw != P:
        raise UcodError(f"decoder_forward: {P} tokens do not form a {gh}x{gw} grid")
    dev = keys_bf16.device
    fg = torch.empty(B, 1, oh, ow, device=dev, dtype=torch.float32)
    bg = torch.empty(B, 1, oh, ow, device=dev, dtype=torch.float32) if want_bg else None
    ortho = torch.empty((), device=dev, dtype=torch.float32) if want_ortho else None
    lib = _lib.load()
    lib.ucod_decoder_workspace_bytes.restype = _u64
    need = lib.ucod_decoder_workspace_bytes(B, gh, gw, oh, ow, 1 if want_ortho else 0)
    ws = _ws(need, dev)
    wp, wn = _aligned(ws)
    with torch.cuda.device(dev):
        _lib.call("ucod_decoder_fwd", ptr(keys_bf16), B, dim, gh, gw, oh, ow, ptr(w_dec_bf16), ptr(b_dec), ptr(emb),
                  ptr(w_fg), ptr(b_fg), ptr(w_bg), ptr(b_bg), ptr(fg), ptr(bg), ptr(ortho), wp, wn, stream_ptr(dev))
    return fg, bg, ortho


def upsample_bilinear(x: torch.Tensor, size, binarize: bool = False) -> torch.Tensor:
    """F.interpolate(x, size, mode='bilinear', align_corners=False) for [B,1,h,w] / [B,h,w] fp32 maps;
    `binarize=True` returns the uint8 mask sigmoid(x) > 0.5 instead."""
    _lib.require_cuda(x)
    shp = x.shape
    x3 = x.reshape(-1, shp[-2], shp[-1]).float().contiguous()
    oh, ow = size
    out = torch.empty(x3.shape[0], oh, ow, device=x.device, dtype=torch.uint8 if binarize else torch.float32)
    with torch.cuda.device(x.device):
        _lib.call("ucod_upsample_bilinear", ptr(x3), ptr(out), x3.shape[0], shp[-2], shp[-1], oh, ow,
                  1 if binarize else 0, stream_ptr(x.device))
    return out.reshape(*shp[:-2], oh, ow)


# ------------------------------------------------------------------------------------------------
def pseudo_label_score(attn_cls: torch.Tensor, keys: torch.Tensor, th_bkg: float, epsilon: float = 1e-10,
                       want_sim: bool = False):
    """attn_cls [B,heads,P] fp32, keys [B,P,heads*64] fp32|bf16 -> (cos [B,P], bkg u8 [B,P], ref_idx [B], sim|None)."""
    _lib.require_cuda(attn_cls, keys)
    attn_cls = attn_cls.float().contiguous()
    keys = keys.contiguous()
    if keys.dtype not in (torch.float32, torch.bfloat16):
        keys = keys.float()
    B, nh, P = attn_cls.shape
    if keys.shape != (B, P, nh * 64):
        raise UcodError(f"pseudo_label_score: keys {tuple(keys.shape)} do not match attention {tuple(attn_cls.shape)}")
    dev = attn_cls.device
    cos = torch.empty(B, P, device=dev, dtype=torch.float32)
    bkg = torch.empty(B, P, device=dev, dtype=torch.uint8)
    ref = torch.empty(B, device=dev, dtype=torch.int32)
    sim = torch.empty(B, P, device=dev, dtype=torch.float32) if want_sim else None
    scratch = torch.empty(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.call("ucod_pseudo_label_score", ptr(attn_cls), ptr(keys), 1 if keys.dtype == torch.bfloat16 else 0, B,
                  nh, P, c_float(th_bkg), c_float(epsilon), ptr(cos), ptr(bkg), ptr(ref), ptr(sim), ptr(scratch),
                  stream_ptr(dev))
    return cos, bkg, ref, sim


def refine_small_components(mask_u8: torch.Tensor, area_threshold: int = 4) -> torch.Tensor:
    """mask uint8 {0,1} [B,h,w] -> refined uint8 [B,h,w] (refine_post_process, batched)."""
    _lib.require_cuda(mask_u8)
    m = mask_u8.to(torch.uint8).contiguous()
    B, h, w = m.shape
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _lib.call("ucod_refine_small_components", ptr(m), ptr(out), B, h, w, int(area_threshold), stream_ptr(m.device))
    return out
