"""Attention throughput at the pseudo-label shape (T = 257) and the main shape."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
for B, T in ((256, 256), (256, 257), (256, 384), (256, 128), (64, 1370)):
    H, D = 12, 64
    qkv = torch.randn(B, T, 3 * H * D, device="cuda").to(torch.bfloat16)
    ctx = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
    ld = 3 * H * D
    args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D, B, H,
            D, T, T, _lib.c_float(0.125), _lib.stream_ptr())
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
    print(f"attention B={B} T={T}: {ms:.3f} ms  {4.0 * B * H * T * T * D / ms / 1e9:.1f} TFLOP/s")
