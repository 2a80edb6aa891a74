"""tcgen05 GEMM building block vs. a plain PyTorch fp32 matmul of the same bf16 operands."""
import ctypes

import pytest
import torch

from ucod_dpl_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(a, w, mode, bias=None, out=None):
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if mode in (2, 5) else torch.bfloat16)
    _lib.call("ucod_gemm_bf16", _lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), M, N, K, mode,
              _lib.ptr(bias), _lib.ptr(out), out.stride(0), _lib.stream_ptr())
    return out


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 64), (256, 256, 128), (300, 768, 768),
                                    (1370, 2304, 768), (1370 * 3 + 5, 768, 3072), (77, 128, 768), (5000, 3072, 768),
                                    (129, 256, 640)])
def test_gemm_bias_f32(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = _gemm(a, w, 5, bias=bias)
    ref = a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max err {err}"


def test_gemm_epilogues():
    M, N, K = 1000, 768, 768
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out0 = _gemm(a, w, 0, bias=bias)
    assert (out0.float() - ref).abs().max().item() < 0.05
    out1 = _gemm(a, w, 1, bias=bias)
    # erf-GELU through the erfc polynomial: error well below the bf16 output resolution
    gref = torch.nn.functional.gelu(ref)
    assert (out1.float() - gref).abs().max().item() < 0.03
    assert ((out1.float() - gref).abs() / gref.abs().clamp_min(1.0)).max().item() < 8e-3
    x = torch.randn(M, N, device="cuda", generator=g)
    x0 = x.clone()
    _gemm(a, w, 2, bias=bias, out=x)
    assert (x - (x0 + ref)).abs().max().item() < 5e-3
    # no-bias variant
    x = x0.clone()
    _gemm(a, w, 2, out=x)
    assert (x - (x0 + ref - bias)).abs().max().item() < 5e-3


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(1370 * 2 + 3, 2304, 768), (77, 768, 3072), (128, 128, 64), (4000, 3072, 768),
                                   (300, 768, 768), (641, 256, 128), (257, 512, 64)])
def test_gemm_tma_epilogue_tails(mode, M, N, K):
    """TMA-store / reduce-add epilogues: ragged M (clipped boxes), many tiles per CTA, strided output rows; M values
    with an odd number of 128-row tiles leave the second CTA of the last pair entirely out of range."""
    g = torch.Generator(device="cuda").manual_seed(M + N + mode)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    pad = 16
    if mode == 2:
        buf = torch.randn(M + 2, N + pad, device="cuda", generator=g)
        guard = buf.clone()
        out = buf[1:M + 1, :N]
        want = out.clone() + ref
    else:
        buf = torch.full((M + 2, N + pad), 7.0, device="cuda", dtype=torch.bfloat16)
        guard = buf.clone()
        out = buf[1:M + 1, :N]
        want = torch.nn.functional.gelu(ref) if mode == 1 else ref
    _gemm(a, w, mode, bias=bias, out=out)
    assert (out.float() - want).abs().max().item() < (5e-3 if mode == 2 else 0.06)
    # nothing outside the [M, N] window was touched (first/last guard rows and the padding columns)
    assert torch.equal(buf[0], guard[0]) and torch.equal(buf[M + 1], guard[M + 1])
    assert torch.equal(buf[:, N:], guard[:, N:])


def test_gemm_strided_rows():
    # A and W taken as column slices of wider buffers (lda/ldw != K), as the QKV weight slices are.
    M, N, K = 512, 256, 128
    g = torch.Generator(device="cuda").manual_seed(2)
    abig = torch.randn(M, 3 * K, device="cuda", generator=g).to(torch.bfloat16)
    wbig = torch.randn(N, 2 * K, device="cuda", generator=g).to(torch.bfloat16)
    a, w = abig[:, K:2 * K], wbig[:, K:]
    out = _gemm(a, w, 5)
    ref = a.float() @ w.float().t()
    assert (out - ref).abs().max().item() < 1e-2


def test_gemm_error_reporting():
    a = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(100, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.UcodError):
        _gemm(a, w, 5)
