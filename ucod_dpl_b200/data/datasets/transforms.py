"""Device-side image transforms (reference: data/datasets/transforms.py:8-43 `ImageTransforms`, the
torchvision `Resize -> ToTensor -> Normalize` pipelines every dataset and `ValLoop_Look_Twice` build).

Same factory names and argument meaning; the returned callables take a decoded image (PIL.Image, HWC uint8
numpy array or uint8 torch tensor) and run the Pillow-exact antialiased bilinear resize (C-ABI
`ucod_roi_crop_resize`, ROI = whole image) and the bit-exact `ToTensor`/`Normalize` (`ucod_to_tensor_normalize`)
on the GPU.  `batch()` does a list of differently sized images in one launch pair (zero-padded canvas, one ROI per
image).  JPEG/PNG decode stays on the host (PIL) — see DESIGN.md, out of scope.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from ... import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _as_hwc_u8(img) -> torch.Tensor:
    """PIL.Image / numpy / torch -> uint8 torch tensor [H,W,C] (C = 1 or 3), device unchanged."""
    if isinstance(img, torch.Tensor):
        t = img
    elif isinstance(img, np.ndarray):
        t = torch.from_numpy(np.array(img, copy=True) if not img.flags.writeable else np.ascontiguousarray(img))
    else:  # PIL.Image (duck-typed so PIL stays an optional import)
        t = torch.from_numpy(np.asarray(img).copy())
    if t.dtype != torch.uint8:
        raise TypeError("image transforms expect decoded 8-bit images")
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if t.dim() != 3 or t.shape[-1] not in (1, 3):
        raise ValueError(f"expected an HWC image with 1 or 3 channels, got shape {tuple(t.shape)}")
    return t


def pack_padded(images: Sequence, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """list of HWC uint8 images (ragged sizes, equal channel count) -> (zero-padded canvas uint8 [N,Hmax,Wmax,C] on
    `device`, sizes int32 [N,2] = (h, w) on the host).  The host staging buffer is pinned so the copy is async."""
    ts = [_as_hwc_u8(im) for im in images]
    if not ts:
        raise ValueError("pack_padded: empty batch")
    C = ts[0].shape[-1]
    if any(t.shape[-1] != C for t in ts):
        raise ValueError("pack_padded: mixed channel counts in one batch")
    hmax = max(t.shape[0] for t in ts)
    wmax = max(t.shape[1] for t in ts)
    sizes = torch.tensor([[t.shape[0], t.shape[1]] for t in ts], dtype=torch.int32)
    if all(t.is_cuda for t in ts):
        canvas = torch.zeros(len(ts), hmax, wmax, C, dtype=torch.uint8, device=ts[0].device)
        for i, t in enumerate(ts):
            canvas[i, : t.shape[0], : t.shape[1]] = t
        return canvas, sizes
    host = torch.zeros(len(ts), hmax, wmax, C, dtype=torch.uint8)
    if torch.cuda.is_available():
        host = host.pin_memory()
    for i, t in enumerate(ts):
        host[i, : t.shape[0], : t.shape[1]] = t.cpu()
    return host.to(device, non_blocking=True), sizes


class DeviceTransform:
    """`transforms.Compose([Resize(size)?, ToTensor()?, Normalize(mean, std)?])` on the GPU."""

    def __init__(self, size=None, to_tensor: bool = True, mean=None, std=None, device="cuda"):
        self.size = None if size is None else (int(size[0]), int(size[1]))
        self.to_tensor = to_tensor
        self.mean, self.std = mean, std
        self.device = torch.device(device)

    def _resize(self, canvas: torch.Tensor, sizes: torch.Tensor) -> torch.Tensor:
        """canvas uint8 [N,H,W,C] -> planar uint8 [N,C,h,w]."""
        N, _, _, C = canvas.shape
        if self.size is None:
            if N != 1 and not bool((sizes == sizes[0]).all()):
                raise ValueError("a transform without Resize cannot batch images of different sizes")
            h, w = int(sizes[0, 0]), int(sizes[0, 1])
            return canvas[:, :h, :w].permute(0, 3, 1, 2).contiguous()
        src = canvas if C == 3 else canvas.expand(-1, -1, -1, 3)  # 'L': the same plane three times (stride 0)
        jobs = torch.zeros(N, 5, dtype=torch.int32)
        jobs[:, 0] = torch.arange(N, dtype=torch.int32)
        jobs[:, 3] = sizes[:, 1]
        jobs[:, 4] = sizes[:, 0]
        out = ops.roi_crop_resize(src, jobs.to(canvas.device), self.size, layout="HWC")
        return out if C == 3 else out[:, :1].contiguous()

    @torch.no_grad()
    def batch(self, images: Sequence) -> torch.Tensor:
        canvas, sizes = pack_padded(images, self.device)
        planar = self._resize(canvas, sizes)
        if not self.to_tensor:
            return planar
        if self.mean is not None and planar.shape[1] != len(self.mean):
            raise ValueError("Normalize: channel count does not match mean/std")
        return ops.to_tensor_normalize(planar, self.mean, self.std)

    def __call__(self, image) -> torch.Tensor:
        return self.batch([image])[0]


class ImageTransforms:
    """Centralised transform configurations (same four factories as the reference)."""

    @staticmethod
    def get_image_transform(image_size: Tuple[int, int] = (518, 518)) -> DeviceTransform:
        return DeviceTransform(image_size, True, IMAGENET_MEAN, IMAGENET_STD)

    @staticmethod
    def get_label_transform(image_size: Tuple[int, int] = (518, 518), keep_size=False) -> DeviceTransform:
        return DeviceTransform(None if keep_size else image_size, True)

    @staticmethod
    def get_feature_extractor_transform(img_size) -> DeviceTransform:
        return DeviceTransform(img_size, True, IMAGENET_MEAN, IMAGENET_STD)

    @staticmethod
    def get_patch_transform() -> DeviceTransform:
        return DeviceTransform(None, True, IMAGENET_MEAN, IMAGENET_STD)

    @staticmethod
    def get_raw_transform(image_size: Tuple[int, int] = (518, 518)) -> DeviceTransform:
        """Resize only, planar uint8 out: the ViT kernels fuse ToTensor/Normalize into the patch embedding, so the
        eval pipelines feed this instead of the fp32 tensor (4x fewer bytes)."""
        return DeviceTransform(image_size, False)
