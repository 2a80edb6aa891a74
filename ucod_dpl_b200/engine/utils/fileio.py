"""On-disk cache format of the reference (engine/utils/fileio/backend/ioctl/pickleio.py:54-142 `MetaListPickleIO`):
one pickle per item (`{prefix}_{index}.pkl`) plus `index.json` mapping str(index) -> file name.  Same constructor,
mode detection (`'r'` when the index and all files exist, `'w'` otherwise) and methods, so pseudo labels / features
written here are consumed by the unmodified reference trainer and vice versa."""
from __future__ import annotations

import json
import os
import pickle
from pathlib import Path
from typing import Optional, Union


class MetaListPickleIO:
    def __init__(self, index_path: Union[Path, str, None] = None, base_path: Union[Path, str, None] = None,
                 file_prefix: str = "data", logger_in=None):
        if index_path is not None:
            self.index_path = Path(index_path)
            self.base_path = self.index_path.parent
        elif base_path is not None:
            self.base_path = Path(base_path)
            self.index_path = self.base_path / "index.json"
        else:
            raise ValueError("Either index_path or base_path must be specified.")
        self.file_prefix = file_prefix
        self.prefix_counter: dict = {}
        self.logger = logger_in
        self.index_map: dict = {}
        self.reload_path()

    @staticmethod
    def check_integrity(index_file_path: Optional[Union[str, Path]]):
        index_file_path = Path(index_file_path)
        if not index_file_path.exists():
            return False, "Index file does not exist."
        with open(index_file_path, "r") as f:
            index_map = json.load(f)
        for index, file in index_map.items():
            if not (index_file_path.parent / file).exists():
                return False, f"File with index {index} does not exist."
        return True, "_"

    def reload_path(self):
        ok, _ = self.check_integrity(self.index_path)
        self.mode = "r" if ok else "w"
        self.index_map = {}
        if ok:
            with open(self.index_path, "r") as f:
                self.index_map = {k: self.base_path / v for k, v in json.load(f).items()}

    def read_file(self, index):
        assert self.mode == "r", "Not working on read mode!"
        with open(self.index_map[str(index)], "rb") as f:
            return pickle.load(f)

    def len(self):
        return len(self.index_map)

    def write_file(self, index, obj, file_name=None):
        assert self.mode == "w", "Not working on write mode!"
        if file_name:
            self.index_map[index] = "{}_{}.pkl".format(file_name, self.prefix_counter.get(file_name, 0))
            self.prefix_counter[file_name] = self.prefix_counter.get(file_name, 0) + 1
        else:
            self.index_map[index] = "{}_{}.pkl".format(self.file_prefix, index)
        path = self.base_path / self.index_map[index]
        os.makedirs(path.parent, exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(obj, f)

    def dump_list(self, obj_list, file_name_list=None):
        for index, obj in enumerate(obj_list):
            self.write_file(index, obj, file_name_list[index] if file_name_list else None)
        os.makedirs(self.base_path, exist_ok=True)
        with open(self.index_path, "w") as f:
            json.dump(self.index_map, f)
