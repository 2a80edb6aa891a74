"""Attention at the pseudo-label shape (256 images, T = 257): target for ncu launch lists."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
B, H, D, T = 256, 12, 64, 257
qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
ld = 3 * H * D
args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H, D, T, T, _lib.c_float(0.125), _lib.stream_ptr())
for _ in range(4):
    _lib.call("ucod_attention", *args)
torch.cuda.synchronize()
