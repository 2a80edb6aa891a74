"""`LRDataset` — the second-stage (CORAL) dataset: per image the low-resolution keys plus the keys of the 3x3 windows of
the 3x-enlarged image, optionally the four overlapping "m" crops of the 54-token-wide map
(reference: data/datasets/lr_dataset.py:15-217).

Same constructor, `get_features(img_path, crop_center=False)`, item keys (`h_inputs`, `m_inputs`, `index` on top of
the first-stage item) and cache directories (`patch_cache`, `m_patch_cache`).  The feature production itself is
`CoralEvaluator.get_features`: Pillow-exact resizes on the device and all windows of a batch of images through the
ViT kernels in one launch sequence, instead of nine B=1 backbone calls per image.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .base_dataset import USCODDataset, _get, read_image


class LRDataset(USCODDataset):
    def __init__(self, config, feature_extractor_cfg, mode: str, dataset_dir: str, cache_dir: Optional[str],
                 logger=None, window_size: int = 3, **kw):
        super().__init__(config=config, feature_extractor_cfg=feature_extractor_cfg, mode=mode,
                         dataset_dir=dataset_dir, cache_dir=cache_dir, logger=logger, **kw)
        self.window_size = int(window_size)
        self.require_m_patches = mode == "train" or bool(_get(config, "require_m_patches", False))
        self.use_cache = bool(_get(config, "use_cache", True))
        self.grid_h, self.grid_w = self.image_size
        self.patch_cache = self.m_patch_cache = None
        self.patches: List[torch.Tensor] = []
        self.m_patches: List[torch.Tensor] = []
        if self.cache_manager is not None:
            self.patch_cache = self.cache_manager.get_patch_cache()
            if self.require_m_patches:
                self.m_patch_cache = self.cache_manager.get_m_patch_cache()
        if self.patch_cache is None or self.patch_cache.mode == "w" or not self.use_cache:
            self._prepare_patch_cache()

    # ---- device feature production ----
    def _producer(self):
        if not hasattr(self, "_coral_features"):
            from ...engine.runner.loop_CORAL import CoralEvaluator
            self.prepare_feature_extractor()
            self._coral_features = CoralEvaluator(self.feature_extractor.feature_extractor, None, None, self.image_size,
                                                  window_size=self.window_size,
                                                  require_m_patches=self.require_m_patches)
        return self._coral_features

    @staticmethod
    def _crop_center(image: np.ndarray) -> np.ndarray:
        """centre half-size crop (lr_dataset.py:122-133), HWC array."""
        h, w = image.shape[:2]
        nh, nw = h // 2, w // 2
        top, left = (h - nh) // 2, (w - nw) // 2
        return image[top:top + nh, left:left + nw]

    @torch.no_grad()
    def features_of(self, images: List[np.ndarray]) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """equal-size HWC uint8 images -> (l [N,C,g,g], h [N,w*w,C,g,g], m [N,4,C,36,36] | None) fp32 on the device."""
        prod = self._producer()
        dev = prod.extractor.device
        batch = torch.from_numpy(np.stack(images)).to(dev)
        l, h, m = prod.get_features(batch, layout="HWC")
        N, P, C = l.shape
        g = int(round(P ** 0.5))
        l = l.reshape(N, g, g, C).permute(0, 3, 1, 2)
        h = h.reshape(N, -1, g, g, C).permute(0, 1, 4, 2, 3)
        if m is not None:
            m = m.reshape(N, 4, 36, 36, C).permute(0, 1, 4, 2, 3)
        return l, h, m

    def get_features(self, img_path: str, crop_center: bool = False):
        """lr_dataset.py:82-120: (list of w*w window key maps [C,g,g] on the CPU, m_patches [1,4,C,36,36] | None);
        with `crop_center` the centre crop is used and (key [1,C,g,g], windows [1,w*w,C,g,g] on the GPU, m) returned."""
        image = read_image(img_path, "RGB")
        if crop_center:
            image = self._crop_center(image)
        l, h, m = self.features_of([image])
        if crop_center:
            return l, h.contiguous(), m
        return [h[0, i].cpu() for i in range(h.shape[1])], m

    def _prepare_patch_cache(self) -> None:
        """lr_dataset.py:170-194, with images of equal size batched into one launch sequence."""
        self.patches, self.m_patches = [None] * len(self), [None] * len(self)
        for batch in self.iter_image_batches(max(1, self.extract_batch // (self.window_size ** 2 + 1))):
            groups: Dict[tuple, List[int]] = {}
            for k, im in enumerate(batch["originals"]):
                groups.setdefault(im.shape, []).append(k)
            for ks in groups.values():
                _, h, m = self.features_of([batch["originals"][k] for k in ks])
                for j, k in enumerate(ks):
                    self.patches[batch["index"][k]] = h[j].contiguous().cpu()
                    if m is not None:
                        self.m_patches[batch["index"][k]] = m[j].contiguous().cpu()
        if self.use_cache and self.patch_cache is not None:
            self.patch_cache.dump_list(self.patches)
            if self.require_m_patches and self.m_patch_cache is not None:
                self.m_patch_cache.dump_list(self.m_patches)

    def __getitem__(self, index: int) -> Dict[str, Any]:
        items = super().__getitem__(index)
        if self.use_cache and self.patch_cache is not None and self.patch_cache.mode == "r":
            h_inputs = self.patch_cache.read_file(index)
            m_inputs = self.m_patch_cache.read_file(index) if self.require_m_patches and self.m_patch_cache else None
        else:
            h_inputs = self.patches[index]
            m_inputs = self.m_patches[index] if self.require_m_patches else None
        items.update({"m_inputs": m_inputs, "h_inputs": h_inputs, "index": [index]})
        return items
