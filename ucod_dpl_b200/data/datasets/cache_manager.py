"""On-disk cache layout of the reference datasets (reference: data/datasets/cache_manager.py:8-97).

    {cache_dir}/features_cache/{extractor type}/{mode}/{DATASET}/   features  (also patch_cache, m_patch_cache)
    {cache_dir}/pseudo_label_cache/{DATASET}/                      pseudo labels (train mode only)

each a `MetaListPickleIO` directory (`index.json` + `data_{i}.pkl`), so caches written here are read by the
unmodified reference datasets and the other way round.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional

from ...engine.utils.fileio import MetaListPickleIO


class CacheManager:
    def __init__(self, base_path: str, logger=None):
        self.base_path = base_path
        self.logger = logger
        self._io: Optional[MetaListPickleIO] = None

    @property
    def io(self) -> MetaListPickleIO:
        if self._io is None:  # opened on first use, like the reference
            self._io = MetaListPickleIO(base_path=self.base_path, logger_in=self.logger)
        return self._io

    @property
    def mode(self) -> str:
        return self.io.mode

    def dump_list(self, data_list: List[Any]) -> None:
        self.io.dump_list(data_list)
        self.io.reload_path()

    def read_file(self, index: int) -> Any:
        return self.io.read_file(index)

    def length(self) -> int:
        return self.io.len()


class MultiCacheManager:
    _PER_EXTRACTOR = ("features", "patch", "m_patch")

    def __init__(self, cache_dir: str, feature_extractor_type: str, mode: str, dataset_name: str, logger=None):
        self.cache_dir = cache_dir
        self.feature_extractor_type = feature_extractor_type
        self.mode = mode
        self.dataset_name = dataset_name
        self.logger = logger
        self._caches: dict = {}

    def cache_path(self, cache_type: str) -> str:
        if cache_type == "pseudo_label":
            return os.path.join(self.cache_dir, "pseudo_label_cache", self.dataset_name)
        return os.path.join(self.cache_dir, f"{cache_type}_cache", self.feature_extractor_type, self.mode,
                            self.dataset_name)

    def get_cache(self, cache_type: str) -> CacheManager:
        if cache_type not in self._caches:
            self._caches[cache_type] = CacheManager(self.cache_path(cache_type), self.logger)
        return self._caches[cache_type]

    def get_features_cache(self) -> CacheManager:
        return self.get_cache("features")

    def get_pseudo_label_cache(self) -> Optional[CacheManager]:
        return self.get_cache("pseudo_label") if self.mode == "train" else None

    def get_patch_cache(self) -> CacheManager:
        return self.get_cache("patch")

    def get_m_patch_cache(self) -> CacheManager:
        return self.get_cache("m_patch")
