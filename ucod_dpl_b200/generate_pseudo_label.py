"""Pseudo-label generation (reference: generate_pseudo_label.py), batched on the GPU.

`refine_post_process(mask, area_threshold=4)` and `generate_mask(image_path, th_bkg=0.6)` keep the reference's
signatures; `PseudoLabelGenerator` is the batched device pipeline the reference's serial B=1 CPU loop
(:141-147) becomes: ViT@224 keys + CLS attention row -> scoring -> 1 - bkg -> small-component cleanup."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .vit import VitKeyExtractor, spec_for

_generator = None


def refine_post_process(mask, area_threshold: int = 4):
    """mask: CPU/GPU tensor [1,h,w] (or [h,w]) of {0,1} -> CPU float tensor [1,h,w] like the reference."""
    m = torch.as_tensor(mask)
    m2 = m.reshape(-1, m.shape[-2], m.shape[-1])[:1]
    out = ops.refine_small_components(m2.to("cuda", torch.uint8), area_threshold)
    return out[0].cpu().unsqueeze(0).float()


class PseudoLabelGenerator:
    def __init__(self, vit_state_dict: dict, kind: str = "dinov2", image_size: int = 224, th_bkg: float = 0.6,
                 area_threshold: int = 4, device="cuda"):
        self.extractor = VitKeyExtractor(vit_state_dict, spec_for(kind), device=device)
        self.image_size, self.th_bkg, self.area_threshold = image_size, th_bkg, area_threshold

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,3,S,S] uint8 raw RGB or fp32 normalised (CUDA) -> uint8 masks [B,g,g] (1 = foreground)."""
        k32, _, att = self.extractor.keys(images, want_f32=True, want_cls_attn=True)
        g = images.shape[-1] // self.extractor.spec.patch
        _, bkg, _, _ = ops.pseudo_label_score(att, k32, self.th_bkg)
        fg = (1 - bkg).reshape(-1, g, g)
        return ops.refine_small_components(fg, self.area_threshold)


def set_generator(gen: PseudoLabelGenerator) -> None:
    global _generator
    _generator = gen


def generate_mask(image_path, th_bkg: float = 0.6):
    """Reference-compatible single-image entry (generate_pseudo_label.py:70-94): PIL load, Resize(224),
    normalise, model, scoring, cleanup -> CPU float tensor [1,16,16]."""
    from PIL import Image
    from torchvision import transforms
    if _generator is None:
        raise RuntimeError("call set_generator(PseudoLabelGenerator(...)) first (the reference's main() does the "
                           "equivalent global initialisation)")
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    x = tf(Image.open(image_path).convert("RGB")).unsqueeze(0).cuda()
    _generator.th_bkg = th_bkg
    return _generator(x)[0].cpu().unsqueeze(0).float()
