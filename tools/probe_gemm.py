"""Quick GEMM throughput probe (CUDA events), ViT-B shapes at B=64."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ucod_dpl_b200 import _lib


def run(M, N, K, mode, iters=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if mode in (2, 5) else torch.bfloat16)
    args = (_lib.ptr(a), K, _lib.ptr(w), K, M, N, K, mode, _lib.ptr(bias), _lib.ptr(out), N,
            _lib.stream_ptr())
    for _ in range(3):
        _lib.call("ucod_gemm_bf16", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _lib.call("ucod_gemm_bf16", *args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for context
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K} mode={mode}: {ms:.3f} ms {tf:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms "
          f"{2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    M = 64 * 1370
    run(M, 2304, 768, 0)
    run(M, 768, 768, 2)
    run(M, 3072, 768, 1)
    run(M, 3072, 768, 0)
    run(M, 768, 3072, 2)
    run(64 * 1369, 128, 768, 5)
