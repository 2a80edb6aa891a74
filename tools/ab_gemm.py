"""A/B two builds of the library on the GEMM shapes of one ViT-B layer (alternating runs in subprocesses).
usage: python tools/ab_gemm.py ab/libA.so ab/libB.so [...]"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from pathlib import Path
from ucod_dpl_b200 import _lib
_lib._LIB_PATH = Path(sys.argv[1]).resolve()
def run(M, N, K, mode, iters=30):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode in (2, 5) else torch.bfloat16)
    args = (_lib.ptr(a), K, _lib.ptr(w), K, M, N, K, mode, _lib.ptr(bias), _lib.ptr(out), N, _lib.stream_ptr())
    for _ in range(5):
        _lib.call("ucod_gemm_bf16", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.call("ucod_gemm_bf16", *args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
M = 64 * 1370
print(" ".join("%%.4f" %% run(M, n, k, m) for n, k, m in ((2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2))))
''' % str(ROOT)
libs = sys.argv[1:]
res = {p: [] for p in libs}
for rep in range(3):
    for p in libs:
        out = subprocess.run([sys.executable, "-c", CHILD, p], capture_output=True, text=True)
        res[p].append([float(v) for v in out.stdout.split()] if out.returncode == 0 else out.stderr[-300:])
for p, rows in res.items():
    print(p)
    for r in rows:
        print("   qkv proj fc1+gelu fc2 [ms]:", r, " sum %.4f" % sum(r) if isinstance(r, list) else "")
