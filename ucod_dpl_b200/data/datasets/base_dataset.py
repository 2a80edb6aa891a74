"""COD image-folder datasets (reference: data/datasets/base_dataset.py:21-176 `BaseCODDataset`,
uscod_dataset.py:9-38 `USCODDataset`).

Layout `{dataset_dir}/{DATASET}/im|gt` ("A+B" concatenates datasets), items
`{"pseudo_label", "label_tensor", "features", "img_path"}` and the feature / pseudo-label cache directories are the
reference's.  What differs is where the work happens: the feature cache is filled in batches on the GPU
(PIL decode on host threads -> pinned uint8 -> Pillow-exact device resize -> ViT key kernels) instead of one
image at a time, and `iter_image_batches` feeds the eval pipelines decoded uint8 images directly (no cache at all).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .cache_manager import MultiCacheManager
from .transforms import ImageTransforms

# engine/utils/fileio/backend/filetype/image_type.py:6 (suffix match is case sensitive there as well)
IMAGE_EXTENSIONS = frozenset(
    "jpg jpeg png bmp gif tiff tif webp heif heic bpg jp2 j2k jpf jpx jpm mj2 svg svgz ico icns cur dds tga exr hdr "
    "pic pnm pbm pgm ppm pam pfm sr ras jpe jpge jif jfif jfi avif avifs apng flif".split())


def list_dir_image(path) -> List[Path]:
    """every image file below `path`, sorted (ImageIO.list_dir_image, imageio.py:132-140)."""
    root = Path(path).resolve()
    return [f for f in sorted(root.glob("**/*")) if f.is_file() and f.suffix.split(".")[-1] in IMAGE_EXTENSIONS]


def read_image(path, mode: str = "RGB") -> np.ndarray:
    """decode on the host (ImageIO(backend='PIL').read_image) -> uint8 [H,W,3] ('RGB') or [H,W] ('L')."""
    from PIL import Image

    with Image.open(path) as im:
        return np.asarray(im.convert(mode))


def _get(cfg, name, default=None):
    try:
        return cfg[name] if name in cfg else default
    except TypeError:
        return getattr(cfg, name, default)


class BaseCODDataset(torch.utils.data.Dataset):
    def __init__(self, config, feature_extractor_cfg, dataset_dir: str, cache_dir: Optional[str] = None,
                 mode: str = "train", load_all: bool = False, keep_size: bool = False,
                 image_size: Tuple[int, int] = (518, 518), require_label: bool = False, logger=None,
                 feature_extractor=None, prepare_cache: bool = True, decode_threads: int = 8,
                 extract_batch: int = 64):
        self.config = config
        self.feature_extractor_cfg = feature_extractor_cfg
        self.mode = mode
        self.cache_dir = cache_dir
        self.logger = logger
        self.load_all = load_all
        self.keep_size = keep_size
        self.image_size = tuple(image_size)
        self.require_label = require_label
        self.decode_threads = decode_threads
        self.extract_batch = extract_batch
        if feature_extractor is not None:
            self.feature_extractor = feature_extractor

        self.transform_image = ImageTransforms.get_image_transform(self.image_size)
        self.transform_raw = ImageTransforms.get_raw_transform(self.image_size)
        self.transform_label = ImageTransforms.get_label_transform(self.image_size, self.load_all or self.keep_size)
        self._setup_file_paths(dataset_dir)
        self.cache_manager = None
        if cache_dir is not None:
            self.cache_manager = MultiCacheManager(cache_dir, _get(feature_extractor_cfg, "type"), mode,
                                                   _get(config, "DATASET"), logger)
            if prepare_cache and self.cache_manager.get_features_cache().mode == "w":
                self._prepare_cache()

    # ---- files ---------------------------------------------------------------------------------
    def _setup_file_paths(self, dataset_dir: str) -> None:
        self.image_paths: List[Path] = []
        self.label_paths: List[Path] = []
        for name in str(_get(self.config, "DATASET")).split("+"):
            self.image_paths.extend(list_dir_image(os.path.join(dataset_dir, name, "im")))
            if self.require_label:
                self.label_paths.extend(list_dir_image(os.path.join(dataset_dir, name, "gt")))
        self.image_paths = sorted(self.image_paths)
        self.label_paths = sorted(self.label_paths)
        if self.require_label:
            self._check_file_mapping(self.image_paths, self.label_paths)

    @staticmethod
    def _check_file_mapping(list_a: Sequence[Path], list_b: Sequence[Path]) -> None:
        assert len(list_a) == len(list_b), "Length of two lists should be the same"
        stems = {p.stem for p in list_b}
        for p in list_a:
            assert p.stem in stems, f"File {p.stem} not found in list_b"

    # ---- device feature extraction -----------------------------------------------------------------
    def prepare_feature_extractor(self) -> None:
        if not hasattr(self, "feature_extractor_transform"):
            kind = _get(self.feature_extractor_cfg, "type")
            self.feature_extractor_transform = ImageTransforms.get_feature_extractor_transform(
                (432, 432) if kind == "dinov1" else (756, 756))
        if not hasattr(self, "feature_extractor"):
            from ..utils.feature_extractor import backbone

            self.feature_extractor = backbone(self.feature_extractor_cfg)

    def _get_features(self, img_tensor: torch.Tensor) -> torch.Tensor:
        if img_tensor.dim() == 3:
            img_tensor = img_tensor.unsqueeze(0)
        _, key = self.feature_extractor(img_tensor)
        return key

    def decode_many(self, paths: Sequence[Path], mode: str = "RGB") -> List[np.ndarray]:
        if self.decode_threads <= 1 or len(paths) <= 1:
            return [read_image(p, mode) for p in paths]
        with ThreadPoolExecutor(self.decode_threads) as pool:
            return list(pool.map(lambda p: read_image(p, mode), paths))

    def iter_image_batches(self, batch_size: int, indices: Optional[Sequence[int]] = None,
                           with_labels: bool = False) -> Iterator[Dict[str, Any]]:
        """decoded batches for the device pipelines: {"index", "img_path", "originals" (list of HWC uint8 arrays),
        "images" (uint8 [b,3,S,S] on the GPU, resized like `transform_image` but not yet normalised — the ViT
        kernels fuse that step), "labels" (list of uint8 [H,W] arrays, only `with_labels`)}."""
        idx = list(range(len(self))) if indices is None else list(indices)
        chunks = [idx[s:s + batch_size] for s in range(0, len(idx), batch_size)]

        def decode(chunk):
            originals = self.decode_many([self.image_paths[i] for i in chunk], "RGB")
            labels = self.decode_many([self.label_paths[i] for i in chunk], "L") if with_labels else None
            return originals, labels

        # the next batch is decoded on host threads while the GPU works on the current one
        with ThreadPoolExecutor(1) as ahead:
            pending = ahead.submit(decode, chunks[0]) if chunks else None
            for n, chunk in enumerate(chunks):
                originals, labels = pending.result()
                pending = ahead.submit(decode, chunks[n + 1]) if n + 1 < len(chunks) else None
                item = {"index": chunk, "img_path": [str(self.image_paths[i]) for i in chunk],
                        "originals": originals, "images": self.transform_raw.batch(originals)}
                if with_labels:
                    item["labels"] = labels
                yield item

    def _prepare_cache(self) -> None:
        """fill the features cache (base_dataset.py:118-139), `extract_batch` images per launch sequence."""
        self.prepare_feature_extractor()
        feats: List[torch.Tensor] = []
        for batch in self.iter_image_batches(self.extract_batch):
            keys = self._get_features(batch["images"])
            feats.extend(k.to("cpu") for k in keys)
        self.cache_manager.get_features_cache().dump_list(feats)

    # ---- torch Dataset protocol ----------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.image_paths)

    def __getitem__(self, index: int) -> Dict[str, Any]:
        label_tensor = None
        if self.label_paths:
            label_tensor = self.transform_label(read_image(self.label_paths[index], "L"))
        features = pseudo_label = None
        if self.cache_manager is not None:
            fc = self.cache_manager.get_features_cache()
            if fc.mode == "r":
                features = fc.read_file(index)
            pc = self.cache_manager.get_pseudo_label_cache()
            if pc is not None and pc.mode == "r":
                pseudo_label = pc.read_file(index)
        return {"pseudo_label": pseudo_label, "label_tensor": label_tensor, "features": features,
                "img_path": str(self.image_paths[index])}


class USCODDataset(BaseCODDataset):
    """first-stage dataset (uscod_dataset.py:9-38): sizes / labels / DATASET come from the split's config node."""

    def __init__(self, config, feature_extractor_cfg, mode: str, dataset_dir: str, cache_dir: Optional[str],
                 logger=None, **kw):
        super().__init__(config=config, feature_extractor_cfg=feature_extractor_cfg, dataset_dir=dataset_dir,
                         cache_dir=cache_dir, mode=mode, load_all=mode == "test",
                         keep_size=bool(_get(config, "keep_size", False)), image_size=_get(config, "image_size"),
                         require_label=bool(_get(config, "require_label", False)), logger=logger, **kw)


def collate_fn(batch: List[Dict[str, Any]]) -> Dict[str, Any]:
    """dataloader_utils.py:13-40: stack what stacks, keep lists where an entry is None or ragged."""
    out: Dict[str, Any] = {}
    for key in batch[0].keys():
        values = [item[key] for item in batch]
        if any(v is None for v in values):
            out[key] = values
            continue
        try:
            out[key] = torch.stack(values) if isinstance(values[0], torch.Tensor) else torch.tensor(values)
        except Exception:
            out[key] = values
    return out
