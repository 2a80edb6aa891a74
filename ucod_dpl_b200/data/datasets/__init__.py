from .base_dataset import BaseCODDataset, USCODDataset, collate_fn, list_dir_image, read_image
from .cache_manager import CacheManager, MultiCacheManager
from .transforms import DeviceTransform, ImageTransforms, pack_padded

__all__ = ["BaseCODDataset", "USCODDataset", "collate_fn", "list_dir_image", "read_image", "CacheManager",
           "MultiCacheManager", "DeviceTransform", "ImageTransforms", "pack_padded"]
