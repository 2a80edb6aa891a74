"""A/B several builds of the library on the ViT-B/14@518 attention shape (alternating subprocess runs).
usage: python tools/ab_attn.py ab/lib_a.so ab/lib_b.so ... [--sdpa]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from pathlib import Path
from ucod_dpl_b200 import _lib
if sys.argv[1] != "sdpa":
    _lib._LIB_PATH = Path(sys.argv[1]).resolve()
B, H, D = 64, 12, 64
res = []
for T in (1370, 257):
    Bt = B if T == 1370 else 256
    qkv = torch.randn(Bt, T, 3 * H * D, device="cuda").to(torch.bfloat16)
    ctx = torch.empty(Bt, T, H * D, device="cuda", dtype=torch.bfloat16)
    ld = 3 * H * D
    q, k, v = [t.reshape(Bt, T, H, D).permute(0, 2, 1, 3).contiguous() for t in qkv.split(H * D, dim=-1)]
    if sys.argv[1] == "sdpa":
        fn = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)
    else:
        args = (_lib.ptr(qkv), ld, _lib.ptr(qkv[..., H * D:]), _lib.ptr(qkv[..., 2 * H * D:]), ld, _lib.ptr(ctx), H * D,
                Bt, H, D, T, T, _lib.c_float(0.125), _lib.stream_ptr())
        fn = lambda: _lib.call("ucod_attention", *args)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    err = 0.0
    if sys.argv[1] != "sdpa":
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(Bt, T, H * D)
        err = (ctx.float() - ref).abs().max().item()
    res.append("T=%%d %%.3f ms %%.0f TFLOP/s err %%.2e" %% (T, ms, 4.0 * Bt * H * T * T * D / ms / 1e9, err))
print(" | ".join(res))
''' % str(ROOT)
libs = [a for a in sys.argv[1:] if a != "--sdpa"] + (["sdpa"] if "--sdpa" in sys.argv else [])
for rep in range(2):
    for p in libs:
        out = subprocess.run([sys.executable, "-c", CHILD, p], capture_output=True, text=True)
        print(f"{p:28s}", out.stdout.strip() if out.returncode == 0 else out.stderr[-400:], flush=True)
