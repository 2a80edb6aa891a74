"""Golden vectors for the two optional arguments of `compute_img_bkg_seg` (reference data/utils/found_bkg_mask.py:9-12):
`up_size` != grid (attention and descriptors bilinearly resampled first, :26-27 and :50-55) and `apply_weights=False`
(:41-48, :63-66).  Run in the build container (imports the reference from /root/reference); the GPU box only reads the
committed tests/golden/bkgseg_options.npz and regenerates the seeded inputs with `planted_inputs`."""
from __future__ import annotations

import importlib.util
import pathlib

import numpy as np
import torch

GOLD = pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden"
CASES = (("up24", dict(up_size=24)), ("noweights", dict(apply_weights=False)),
         ("up20_noweights", dict(up_size=20, apply_weights=False)))
TH_BKG = 0.6


def planted_inputs(B=2, grid=16, nh=12, seed=23):
    """Seeded CLS attention [B,nh,T,T] (only row 0 is read) and last-layer keys [B,T,nh*64] with 2-3 planted clusters."""
    P = grid * grid
    g = torch.Generator().manual_seed(seed)
    feats = torch.empty(B, P + 1, nh * 64)
    att = torch.zeros(B, nh, P + 1, P + 1)
    for b in range(B):
        nc = 2 + (b % 2)
        centres = torch.randn(nc, nh * 64, generator=g)
        # blocky assignment so that the bilinear resample has smooth regions as well as edges
        coarse = torch.randint(0, nc, (grid // 4, grid // 4), generator=g)
        assign = coarse.repeat_interleave(4, 0).repeat_interleave(4, 1).reshape(-1)
        feats[b, 1:] = centres[assign] + 0.3 * torch.randn(P, nh * 64, generator=g)
        feats[b, 0] = torch.randn(nh * 64, generator=g)
        logits = torch.randn(nh, nc, generator=g)[:, assign] * 2 + 0.3 * torch.randn(nh, P, generator=g)
        att[b, :, 0, :] = torch.softmax(torch.cat([torch.zeros(nh, 1), logits], 1), dim=1)
    return att, feats


def main():
    spec = importlib.util.spec_from_file_location("ref_found_bkg_mask", "/root/reference/data/utils/found_bkg_mask.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    att, feats = planted_inputs()
    res = {}
    for name, kw in CASES:
        m, s = ref.compute_img_bkg_seg(att, feats, (16, 16), TH_BKG, dim=64, **kw)
        res[f"{name}_bkg"], res[f"{name}_sim"] = m.numpy().astype(np.uint8), s.numpy()
        print(name, m.shape, float(m.mean()))
    np.savez_compressed(GOLD / "bkgseg_options.npz", **res)


if __name__ == "__main__":
    main()
