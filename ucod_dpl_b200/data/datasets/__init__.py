from .base_dataset import BaseCODDataset, USCODDataset, collate_fn, list_dir_image, read_image
from .lr_dataset import LRDataset
from .cache_manager import CacheManager, MultiCacheManager
from .transforms import DeviceTransform, ImageTransforms, pack_padded

__all__ = ["BaseCODDataset", "USCODDataset", "collate_fn", "list_dir_image", "read_image", "LRDataset", "CacheManager",
           "MultiCacheManager", "DeviceTransform", "ImageTransforms", "pack_padded"]
