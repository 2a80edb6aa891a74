"""Debug: per-tile wait/busy cycles of one GEMM CTA (library built with UCOD_NVCC_EXTRA=-DUCOD_GEMM_TIMELINE)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib
M, N, K, mode = 64 * 1370, int(sys.argv[1]) if len(sys.argv) > 1 else 2304, int(sys.argv[2]) if len(sys.argv) > 2 else 768, int(sys.argv[3]) if len(sys.argv) > 3 else 0
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode in (2, 5) else torch.bfloat16)
args = (_lib.ptr(a), K, _lib.ptr(w), K, M, N, K, mode, _lib.ptr(bias), _lib.ptr(out), N, _lib.stream_ptr())
for _ in range(3):
    _lib.call("ucod_gemm_bf16", *args)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (2 * 32 * 4))()
_lib.load().ucod_debug_gemm_timeline(buf)
print(f"N={N} K={K} mode={mode}: ideal MMA clks/tile = {128 * (K // 16)}")
print("tile: mma[wait_epilogue, wait_tma, tile_total] | epi[wait_acc, busy]")
for it in range(0, 32, 2):
    m = [buf[(0 * 32 + it) * 4 + s] for s in range(3)]
    e = [buf[(1 * 32 + it) * 4 + s] for s in range(2)]
    print(it, m, e)
