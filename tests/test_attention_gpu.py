"""tcgen05 fused attention vs. PyTorch fp32 softmax attention on the same bf16 q/k/v."""
import pytest
import torch

from ucod_dpl_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(B, H, T, seed=0, scale_in=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = (torch.randn(B, H, T, 64, device="cuda", generator=g) * scale_in).to(torch.bfloat16)
    k = (torch.randn(B, H, T, 64, device="cuda", generator=g) * scale_in).to(torch.bfloat16)
    v = torch.randn(B, H, T, 64, device="cuda", generator=g).to(torch.bfloat16)
    Tpad = (T + 7) // 8 * 8
    vt = torch.zeros(B, H, 64, Tpad, device="cuda", dtype=torch.bfloat16)
    vt[..., :T] = v.transpose(-1, -2)
    ctx = torch.full((B, T, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.call("ucod_attention_d64", _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(ctx), B, H, T, Tpad,
              _lib.c_float(0.125), _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = torch.softmax((q.float() @ k.float().transpose(-1, -2)) * 0.125, dim=-1) @ v.float()
    ref = ref.permute(0, 2, 1, 3).reshape(B, T, H * 64)
    return ctx.float(), ref


@pytest.mark.parametrize("B,H,T", [(1, 1, 128), (1, 2, 256), (2, 12, 257), (1, 12, 1370), (2, 3, 90), (1, 4, 2917)])
def test_attention_matches_fp32(B, H, T):
    out, ref = _run(B, H, T, seed=T)
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item()
    assert err < 2e-2, f"max abs err {err}"


def test_attention_peaky_softmax():
    # large logits: exercises the running-max rescale of O
    out, ref = _run(1, 2, 700, seed=3, scale_in=4.0)
    err = (out - ref).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
