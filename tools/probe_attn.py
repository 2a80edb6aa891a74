"""Attention throughput probe at ViT-B/14@518 shapes (fused QKV buffer, read in place)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib

B, H, T, D = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 12, 1370, 64
qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
ld = 3 * H * D
args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H, D,
        T, T, _lib.c_float(0.125), _lib.stream_ptr())
for _ in range(3):
    _lib.call("ucod_attention", *args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.call("ucod_attention", *args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * B * H * T * T * D
print(f"attention B={B} T={T}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
q, k, v = [t.reshape(B, T, H, D).permute(0, 2, 1, 3).contiguous() for t in qkv.split(H * D, dim=-1)]
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, T, H * D)
print("max abs diff vs sdpa:", (ctx.float() - ref.float()).abs().max().item())
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
