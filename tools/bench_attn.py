"""Attention timing at the two ViT shapes (B=64, T=1370 and B=256, T=257), with the error against fp32 SDPA."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
H, D = 12, 64
for T, B in ((1370, 64), (257, 256)):
    qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
    ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
    ld = 3 * H * D
    args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H, D,
            T, T, _lib.c_float(0.125), _lib.stream_ptr())
    for _ in range(5):
        _lib.call("ucod_attention", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.call("ucod_attention", *args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    q, k, v = [t.reshape(8, T, H, D).permute(0, 2, 1, 3).float() for t in qkv[:8].split(H * D, dim=-1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(8, T, H * D)
    err = (ctx[:8].float() - ref).abs().max().item()
    err_tail = (ctx[:8, -1].float() - ref[:, -1]).abs().max().item()
    print(f"T={T} B={B}: {ms:.4f} ms, {4.0 * B * H * T * T * D / ms / 1e9:.0f} TFLOP/s, max err {err:.2e} (last row {err_tail:.2e})")
