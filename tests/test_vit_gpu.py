"""ViT key extractor (CUDA, bf16 tensor cores) vs. the fp32 CPU oracle on the same seeded weights and images."""
import pytest
import torch

from oracle import vit as ovit
from ucod_dpl_b200.vit import VitKeyExtractor, spec_for

pytestmark = pytest.mark.gpu


def _images(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, 3, S, S), generator=g, dtype=torch.uint8)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("kind,S,B", [("dinov2", 224, 2), ("dinov2", 518, 1), ("dinov1", 296, 1), ("dinov2", 252, 3)])
def test_keys_match_oracle(kind, S, B):
    ospec = ovit.spec_for(kind)
    sd = ovit.random_vit_state_dict(ospec, seed=0)
    u8 = _images(B, S, 1234)
    x = ovit.normalize_u8(u8)
    ref = ovit.vit_forward(sd, ospec, x, want_attn=True)
    ext = VitKeyExtractor(sd, spec_for(kind))
    k32, k16, att = ext.keys(x.cuda(), want_f32=True, want_bf16=True, want_cls_attn=True)
    torch.cuda.synchronize()
    ref_k = ref["key_tokens"][:, 1:]
    assert k32.shape == ref_k.shape
    e = _rel(k32.cpu(), ref_k)
    assert e < 3e-2, f"keys rel err {e}"
    assert _rel(k16.float().cpu(), ref_k) < 4e-2
    ea = (att.cpu() - ref["cls_attn"]).abs().max().item() / ref["cls_attn"].abs().max().item()
    assert ea < 5e-2, f"cls attn rel err {ea}"
    # uint8 input path (normalisation fused into the patch loader) must agree with the fp32 path
    k32_u8, _, _ = ext.keys(u8.cuda(), want_f32=True)
    assert _rel(k32_u8.cpu(), k32.cpu()) < 2e-2
    # keep_cls variant returns the CLS row first
    kc, _, _ = ext.keys(x.cuda(), want_f32=True, keep_cls=True)
    assert kc.shape[1] == ref["key_tokens"].shape[1]
    assert _rel(kc.cpu(), ref["key_tokens"]) < 3e-2
