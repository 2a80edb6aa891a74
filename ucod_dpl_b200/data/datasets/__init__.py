from .base_dataset import BaseCODDataset, USCODDataset, collate_fn, list_dir_image, read_image
from .dataloader_utils import (DataLoaderFactory, init_testloaders, init_testloaders_LR, init_trainloader,
                               init_trainloader_LR)
from .lr_dataset import LRDataset
from .cache_manager import CacheManager, MultiCacheManager
from .transforms import DeviceTransform, ImageTransforms, pack_padded

__all__ = ["BaseCODDataset", "USCODDataset", "collate_fn", "list_dir_image", "read_image", "LRDataset", "DataLoaderFactory", "init_trainloader", "init_testloaders", "init_trainloader_LR",
           "init_testloaders_LR", "CacheManager",
           "MultiCacheManager", "DeviceTransform", "ImageTransforms", "pack_padded"]
