"""Where the reference datasets keep their caches, and lazy handles on them
(reference: data/datasets/cache_manager.py:8-97).

    kind            directory under cache_dir
    features        features_cache/{extractor type}/{mode}/{DATASET}
    patch           patch_cache/{extractor type}/{mode}/{DATASET}
    m_patch         m_patch_cache/{extractor type}/{mode}/{DATASET}
    pseudo_label    pseudo_label_cache/{DATASET}                    (train mode only)

Every directory is a `MetaListPickleIO` store (`index.json` + `data_{i}.pkl`), so what this package writes is read
by the unmodified reference datasets and the other way round.  `CacheManager` / `MultiCacheManager` keep the
reference's method names.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, List, Optional

from ...engine.utils.fileio import MetaListPickleIO

_SHARED_ACROSS_EXTRACTORS = {"pseudo_label"}  # stored by data set only


def cache_directory(cache_dir, kind: str, extractor_type: str, mode: str, dataset_name: str) -> str:
    root = Path(cache_dir) / f"{kind}_cache"
    if kind not in _SHARED_ACROSS_EXTRACTORS:
        root = root / extractor_type / mode
    return str(root / dataset_name)


class CacheManager:
    """One cache directory; the `MetaListPickleIO` behind it is opened on first use, like the reference."""

    def __init__(self, base_path: str, logger=None):
        self.base_path, self.logger = base_path, logger
        self._store: Optional[MetaListPickleIO] = None

    @property
    def io(self) -> MetaListPickleIO:
        if self._store is None:
            self._store = MetaListPickleIO(base_path=self.base_path, logger_in=self.logger)
        return self._store

    @property
    def mode(self) -> str:  # 'r' once index.json and every item file exist, 'w' before
        return self.io.mode

    def length(self) -> int:
        return self.io.len()

    def read_file(self, index: int) -> Any:
        return self.io.read_file(index)

    def dump_list(self, data_list: List[Any]) -> None:
        store = self.io
        store.dump_list(data_list)
        store.reload_path()  # flips the handle to read mode


class MultiCacheManager:
    def __init__(self, cache_dir: str, feature_extractor_type: str, mode: str, dataset_name: str, logger=None):
        self.cache_dir, self.feature_extractor_type = cache_dir, feature_extractor_type
        self.mode, self.dataset_name, self.logger = mode, dataset_name, logger
        self._caches: Dict[str, CacheManager] = {}

    def cache_path(self, cache_type: str) -> str:
        return cache_directory(self.cache_dir, cache_type, self.feature_extractor_type, self.mode, self.dataset_name)

    def get_cache(self, cache_type: str) -> CacheManager:
        handle = self._caches.get(cache_type)
        if handle is None:
            handle = self._caches[cache_type] = CacheManager(self.cache_path(cache_type), self.logger)
        return handle

    def get_features_cache(self) -> CacheManager:
        return self.get_cache("features")

    def get_patch_cache(self) -> CacheManager:
        return self.get_cache("patch")

    def get_m_patch_cache(self) -> CacheManager:
        return self.get_cache("m_patch")

    def get_pseudo_label_cache(self) -> Optional[CacheManager]:
        return self.get_cache("pseudo_label") if self.mode == "train" else None
