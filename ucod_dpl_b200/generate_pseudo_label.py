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


def write_pseudo_label_cache(masks_u8: torch.Tensor, cache_dir) -> int:
    """Store masks uint8 [N,g,g] the way the reference's `main()` does (generate_pseudo_label.py:124,150): item i is a
    CPU float tensor [1,g,g], `data_{i}.pkl` + `index.json` (MetaListPickleIO), readable by the reference trainer."""
    from .engine.utils.fileio import MetaListPickleIO
    io = MetaListPickleIO(base_path=cache_dir)
    if io.mode != "w":
        raise RuntimeError(f"{cache_dir} already holds a valid cache")
    m = masks_u8.detach().cpu()
    io.dump_list([m[i].unsqueeze(0).float() for i in range(m.shape[0])])
    return m.shape[0]


@torch.no_grad()
def generate_pseudo_label_cache(generator: PseudoLabelGenerator, images_u8: torch.Tensor, cache_dir,
                                batch: int = 256) -> int:
    """Data-parallel version of the reference's serial loop (:141-150): this rank scores the images
    `dist.shard_indices(N)`, the uint8 masks are gathered on rank 0, which writes the cache."""
    from . import dist as ud
    N = images_u8.shape[0]
    idx = list(ud.shard_indices(N))
    outs = []
    for i in range(0, len(idx), batch):
        sel = torch.as_tensor(idx[i:i + batch], dtype=torch.long)
        outs.append(generator(images_u8[sel].to(generator.extractor.device)))
    g = images_u8.shape[-1] // generator.extractor.spec.patch
    local = torch.cat(outs, 0) if outs else torch.zeros(0, g, g, dtype=torch.uint8, device=generator.extractor.device)
    full = ud.gather_sharded_masks(local, N)
    rank, _ = ud.world()
    if rank == 0:
        return write_pseudo_label_cache(full, cache_dir)
    return 0


@torch.no_grad()
def generate_from_folders(generator: PseudoLabelGenerator, image_paths, cache_dir, batch: int = 256,
                          decode_threads: int = 8) -> int:
    """The reference's `main()` loop (:126-150) over image files: host decode on a thread pool, Pillow-exact
    resize to 224^2 on the device, batched scoring; this rank takes `dist.shard_indices`, rank 0 writes the cache."""
    from concurrent.futures import ThreadPoolExecutor

    from . import dist as ud
    from .data.datasets.base_dataset import read_image
    from .data.datasets.transforms import ImageTransforms
    tf = ImageTransforms.get_raw_transform((generator.image_size, generator.image_size))
    idx = list(ud.shard_indices(len(image_paths)))
    outs = []
    with ThreadPoolExecutor(decode_threads) as pool:
        for s in range(0, len(idx), batch):
            imgs = list(pool.map(lambda i: read_image(image_paths[i], "RGB"), idx[s:s + batch]))
            outs.append(generator(tf.batch(imgs)))
    g = generator.image_size // generator.extractor.spec.patch
    dev = generator.extractor.device
    local = torch.cat(outs, 0) if outs else torch.zeros(0, g, g, dtype=torch.uint8, device=dev)
    full = ud.gather_sharded_masks(local, len(image_paths))
    if ud.world()[0] == 0:
        return write_pseudo_label_cache(full, cache_dir)
    return 0


def main(argv=None) -> int:
    """`python -m ucod_dpl_b200.generate_pseudo_label --dataset TR-CAMO+TR-COD10K` — the reference's command line."""
    import argparse
    import os
    from types import SimpleNamespace

    from .data.utils.feature_extractor import load_vit_state_dict
    p = argparse.ArgumentParser(description="Generate pseudo labels for COD datasets using DINOv2")
    p.add_argument("--dataset", type=str, default="TR-CAMO+TR-COD10K")
    p.add_argument("--image_path", type=str, default="./datasets/RefCOD/{}/im")
    p.add_argument("--cache_path", type=str, default="./datasets/cache/pseudo_label_cache/")
    p.add_argument("--backbone_weights", type=str, default="./weights")
    args = p.parse_args(argv)
    paths = []
    for name in args.dataset.split("+"):
        d = args.image_path.format(name)
        if not os.path.exists(d):
            raise ValueError(f"Image path {d} does not exist!")
        paths += [os.path.join(d, f) for f in os.listdir(d)]
    paths = sorted(paths)
    print(f"Found {len(paths)} images from {args.dataset}.")
    sd = load_vit_state_dict(SimpleNamespace(type="dinov2", backbone="facebook/dinov2-base",
                                             backbone_weights=args.backbone_weights))
    n = generate_from_folders(PseudoLabelGenerator(sd, "dinov2"), paths, os.path.join(args.cache_path, args.dataset))
    print(f"Successfully generated {n} pseudo labels and saved to cache.")
    return n


if __name__ == "__main__":
    main()
