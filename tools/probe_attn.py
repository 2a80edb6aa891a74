"""Attention throughput probe at ViT-B/14@518 shapes."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib

B, H, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 12, 1370
Tpad = (T + 7) // 8 * 8
q = torch.randn(B, H, T, 64, device="cuda").to(torch.bfloat16)
k = torch.randn(B, H, T, 64, device="cuda").to(torch.bfloat16)
vt = torch.zeros(B, H, 64, Tpad, device="cuda", dtype=torch.bfloat16)
vt[..., :T] = torch.randn(B, H, 64, T, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
args = (_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(ctx), B, H, T, Tpad, _lib.c_float(0.125), _lib.stream_ptr())
for _ in range(3):
    _lib.call("ucod_attention_d64", *args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.call("ucod_attention_d64", *args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * B * H * T * T * 64
print(f"attention B={B} T={T}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
v = vt[..., :T].transpose(-1, -2).contiguous()
for _ in range(3):
    torch.nn.functional.scaled_dot_product_attention(q, k, v)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    torch.nn.functional.scaled_dot_product_attention(q, k, v)
e1.record()
torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / 10
print(f"torch sdpa: {ms2:.3f} ms  {fl / ms2 / 1e9:.1f} TFLOP/s")
