"""DBA decoder (CUDA) vs. the fp32 oracle; shipped checkpoints loaded through the drop-in `baseline`."""
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

from oracle import decoder as odec
from ucod_dpl_b200 import ops
from ucod_dpl_b200.models.uscod import baseline

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _random_decoder_sd(seed=0, dim=768):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for p in ("decoder.", "decoder_ema."):
        sd[p + "decoupling.weight"] = torch.randn(128, dim, 1, 1, generator=g) * 0.05
        sd[p + "decoupling.bias"] = torch.randn(128, generator=g) * 0.1
        sd[p + "learnable_embedding"] = torch.randn(2, 64, generator=g)
        sd[p + "conv_out_fg.weight"] = torch.randn(1, 64, 1, 1, generator=g) * 0.3
        sd[p + "conv_out_fg.bias"] = torch.randn(1, generator=g) * 0.1
        sd[p + "conv_out_bg.weight"] = torch.randn(1, 64, 1, 1, generator=g) * 0.3
        sd[p + "conv_out_bg.bias"] = torch.randn(1, generator=g) * 0.1
    return sd


def _model(sd):
    m = baseline(SimpleNamespace(dim=768)).cuda().eval()
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.parametrize("B,S", [(2, 68), (1, 37), (3, 56)])
def test_dropin_forward_matches_oracle(B, S):
    sd = _random_decoder_sd(1)
    m = _model(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 768, S, S, generator=g)
    fg_r, bg_r, ortho_r = odec.baseline_forward(sd, x)
    ema_r = odec.baseline_forward(sd, x, ema=True)
    with torch.no_grad():
        fg, bg, ortho = m(x.cuda())
        ema = m(x.cuda(), ema=True)
    torch.cuda.synchronize()
    scale = fg_r.abs().max().item()
    assert (fg.cpu() - fg_r).abs().max().item() < 2e-2 * max(scale, 1.0)
    assert (bg.cpu() - bg_r).abs().max().item() < 2e-2 * max(bg_r.abs().max().item(), 1.0)
    assert (ema.cpu() - ema_r).abs().max().item() < 2e-2 * max(ema_r.abs().max().item(), 1.0)
    assert abs(ortho.item() - ortho_r.item()) < 2e-2 * abs(ortho_r.item()) + 1e-9
    # probabilities: the tolerance BASELINE.json states (<= 1e-2 max-abs on sigmoid outputs)
    assert (torch.sigmoid(fg.cpu()) - torch.sigmoid(fg_r)).abs().max().item() < 1e-2


def test_commuted_upsample_path_matches_reference_order():
    """conv-then-upsample (ours) == upsample-then-conv (loop_UCOD_DPL.py:305 + DBA.py:35) within bf16 noise."""
    sd = _random_decoder_sd(2)
    m = _model(sd)
    g = torch.Generator().manual_seed(6)
    k37 = torch.randn(2, 768, 37, 37, generator=g)
    x68 = odec.upsample_bilinear(k37, (68, 68))
    fg_r, bg_r, _ = odec.baseline_forward(sd, x68, want_ortho=False)
    tokens = ops.features_to_tokens_bf16(k37.cuda())
    fg, bg, _ = m.decoder.forward_tokens(tokens, (37, 37), (68, 68))
    torch.cuda.synchronize()
    assert (torch.sigmoid(fg.cpu()) - torch.sigmoid(fg_r)).abs().max().item() < 1e-2
    assert (torch.sigmoid(bg.cpu()) - torch.sigmoid(bg_r)).abs().max().item() < 1e-2
    agree = ((fg.cpu() > 0) == (fg_r > 0)).float().mean().item()
    assert agree >= 0.999


def test_channels_last_view_input():
    # the backbone hands out keys as a permuted view [B,HW,C] -> [B,C,H,W]; no copy must be needed
    sd = _random_decoder_sd(3)
    m = _model(sd)
    g = torch.Generator().manual_seed(7)
    tok = torch.randn(2, 37 * 37, 768, generator=g)
    x = tok.reshape(2, 37, 37, 768).permute(0, 3, 1, 2)
    fg_r, _, _ = odec.baseline_forward(sd, x.contiguous(), want_ortho=False)
    with torch.no_grad():
        fg = m(tok.cuda().reshape(2, 37, 37, 768).permute(0, 3, 1, 2), ema=False)[0]
    assert (torch.sigmoid(fg.cpu()) - torch.sigmoid(fg_r)).abs().max().item() < 1e-2


@pytest.mark.parametrize("h,w,oh,ow", [(68, 68, 518, 518), (68, 68, 296, 296), (16, 16, 68, 68), (37, 37, 37, 37),
                                       (168, 168, 401, 333)])
def test_upsample_bilinear(h, w, oh, ow):
    g = torch.Generator().manual_seed(h + oh)
    x = torch.randn(3, 1, h, w, generator=g)
    ref = odec.upsample_bilinear(x, (oh, ow))
    out = ops.upsample_bilinear(x.cuda(), (oh, ow))
    assert (out.cpu() - ref).abs().max().item() < 1e-5
    mask = ops.upsample_bilinear(x.cuda(), (oh, ow), binarize=True).cpu()
    ref_mask = (torch.sigmoid(ref) > 0.5).to(torch.uint8)
    mism = (mask != ref_mask)
    # mismatches may only sit on the fp32 rounding edge of the threshold
    assert mism.float().mean().item() < 1e-4
    assert (ref[mism].abs() < 1e-5).all()
