"""Loader construction for the four dataset/split combinations the reference builds
(data/datasets/dataloader_utils.py:43-204: `DataLoaderFactory` and the `init_*loader*` shorthands).

One table drives all of them: (dataset class, which split's config node, which loader config node).  The loaders
run in the calling process (`num_workers` of the reference configs is 0 as well) because items are produced by
device kernels; `pin_memory` is left off for the same reason (label tensors already live on the GPU).
"""
from __future__ import annotations

from torch.utils.data import DataLoader

from .base_dataset import USCODDataset, collate_fn
from .lr_dataset import LRDataset

# kind -> (dataset class, mode, split config attribute, loader config attribute, shuffle comes from the config?)
_KINDS = {
    "train": (USCODDataset, "train", "trainset_cfg", "trainloader_cfg", True),
    "test": (USCODDataset, "test", "valset_cfg", "val_loader_cfg", False),
    "lr_train": (LRDataset, "train", "trainset_cfg", "trainloader_cfg", True),
    "lr_test": (LRDataset, "test", "valset_cfg", "val_loader_cfg", False),
}


def _build(kind: str, config, logger=None, **dataset_kw) -> DataLoader:
    cls, mode, split_attr, loader_attr, may_shuffle = _KINDS[kind]
    split, lcfg = config[split_attr], config[loader_attr]
    dataset = cls(config=split, feature_extractor_cfg=config.feature_extractor_cfg, mode=mode,
                  dataset_dir=config.dataset_dir, cache_dir=config.cache_dir, logger=logger, **dataset_kw)
    loader = DataLoader(dataset, batch_size=int(lcfg.batch_size), num_workers=0,
                        shuffle=bool(lcfg.shuffle) if may_shuffle else False, collate_fn=collate_fn)
    if logger is not None and hasattr(logger, "log"):
        logger.log(f"{len(loader)} batches of {mode} dataloader {split.DATASET} has been created.")
    return loader


class DataLoaderFactory:
    @staticmethod
    def create_train_loader(config, logger=None) -> DataLoader:
        return _build("train", config, logger)

    @staticmethod
    def create_test_loader(config, logger=None) -> DataLoader:
        return _build("test", config, logger)

    @staticmethod
    def create_lr_train_loader(config, logger=None, window_size: int = 3) -> DataLoader:
        return _build("lr_train", config, logger, window_size=window_size)

    @staticmethod
    def create_lr_test_loader(config, logger=None, window_size: int = 3) -> DataLoader:
        return _build("lr_test", config, logger, window_size=window_size)


def init_trainloader(config, logger=None) -> DataLoader:
    return DataLoaderFactory.create_train_loader(config, logger)


def init_testloaders(config, logger=None) -> DataLoader:
    return DataLoaderFactory.create_test_loader(config, logger)


def init_trainloader_LR(config, logger=None, window_size: int = 3) -> DataLoader:
    return DataLoaderFactory.create_lr_train_loader(config, logger, window_size)


def init_testloaders_LR(config, logger=None, window_size: int = 3) -> DataLoader:
    return DataLoaderFactory.create_lr_test_loader(config, logger, window_size)
