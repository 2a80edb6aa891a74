"""Attribute-access configuration tree with `_BASE_` inheritance.

API-compatible with the part of the reference's `engine.config.CfgNode` (engine/config/config.py:66-191,
244-263,443-458) that the hot path and its launch scripts use: `CfgNode(dict)`, attribute + item access,
`CfgNode.load_with_base(path)` with recursive `_BASE_` lists resolved relative to the including file,
python-source configs exporting `cfg = dict(...)`, YAML configs, `dump()`, `get()`, `clone()`,
`merge_from_file/merge_from_other_cfg/merge_from_list`, `freeze()/defrost()`.
The reference's own config files load unchanged (tests/test_host_api.py::test_configs_resolve_like_the_reference).
"""
from __future__ import annotations

import copy
import os
import runpy
from ast import literal_eval
from typing import Any, Dict, Iterable

import yaml

BASE_KEY = "_BASE_"
_LEAF_TYPES = (tuple, list, str, int, float, bool, type(None))


class CfgNode(dict):
    _FROZEN = "__immutable__"
    _NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict: Dict[str, Any] | None = None, key_list=None, new_allowed: bool = False):
        super().__init__()
        path = list(key_list or [])
        for k, v in (init_dict or {}).items():
            if isinstance(v, dict):
                v = CfgNode(v, key_list=path + [str(k)], new_allowed=new_allowed)
            elif not isinstance(v, _LEAF_TYPES):
                raise TypeError(f"config key {'.'.join(path + [str(k)])} has unsupported type {type(v).__name__}")
            dict.__setitem__(self, k, v)
        object.__setattr__(self, CfgNode._FROZEN, False)
        object.__setattr__(self, CfgNode._NEW_ALLOWED, new_allowed)

    # ---- attribute access ----
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"attempted to set {name} on a frozen CfgNode")
        if name in (CfgNode._FROZEN, CfgNode._NEW_ALLOWED):
            raise AttributeError(f"reserved attribute name {name}")
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value, new_allowed=self.is_new_allowed())
        self[name] = value

    def __deepcopy__(self, memo):
        out = CfgNode(new_allowed=self.is_new_allowed())
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        object.__setattr__(out, CfgNode._FROZEN, self.is_frozen())
        return out

    # ---- state ----
    def is_frozen(self) -> bool:
        return bool(self.__dict__.get(CfgNode._FROZEN, False))

    def is_new_allowed(self) -> bool:
        return bool(self.__dict__.get(CfgNode._NEW_ALLOWED, False))

    def _set_frozen(self, flag: bool) -> None:
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self) -> None:
        self._set_frozen(True)

    def defrost(self) -> None:
        self._set_frozen(False)

    def set_new_allowed(self, flag: bool) -> None:
        object.__setattr__(self, CfgNode._NEW_ALLOWED, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.set_new_allowed(flag)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    # ---- (de)serialisation ----
    def to_dict(self) -> Dict[str, Any]:
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs) -> str:
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, tuple):
                return [plain(x) for x in v]
            if isinstance(v, list):
                return [plain(x) for x in v]
            return v

        return yaml.safe_dump(plain(self.to_dict()), **kwargs)

    def __str__(self) -> str:
        lines = []
        for k, v in sorted(self.items(), key=lambda kv: str(kv[0])):
            if isinstance(v, CfgNode):
                body = "\n".join("  " + ln for ln in str(v).splitlines())
                lines.append(f"{k}:\n{body}")
            else:
                lines.append(f"{k}: {v}")
        return "\n".join(lines)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({dict.__repr__(self)})"

    # ---- loading ----
    @classmethod
    def _read_file(cls, filename: str) -> "CfgNode":
        ext = os.path.splitext(filename)[1]
        if ext == ".py":
            ns = runpy.run_path(filename)
            if "cfg" not in ns or not isinstance(ns["cfg"], dict):
                raise ValueError(f"python config {filename} must define a dict named 'cfg'")
            return cls(ns["cfg"])
        if ext in ("", ".yaml", ".yml"):
            with open(filename, "r") as f:
                return cls(yaml.safe_load(f) or {})
        raise ValueError(f"unsupported config file type: {filename}")

    @classmethod
    def load_cfg(cls, cfg_file_obj_or_str) -> "CfgNode":
        if isinstance(cfg_file_obj_or_str, str):
            return cls(yaml.safe_load(cfg_file_obj_or_str) or {})
        return cls._read_file(cfg_file_obj_or_str.name)

    @classmethod
    def load_with_base(cls, filename: str, new_allowed: bool = True) -> "CfgNode":
        """Load `filename`; every file named in its `_BASE_` (str or list, relative to `filename`) is loaded first
        (recursively), later bases override earlier ones and the file itself overrides all of them."""
        node = cls._read_file(filename)
        node.set_new_allowed(new_allowed)
        if BASE_KEY not in node:
            return node
        bases = node.pop(BASE_KEY)
        if isinstance(bases, str):
            bases = [bases]
        merged = cls(new_allowed=new_allowed)
        for b in bases:
            b = os.path.expanduser(b) if b.startswith("~") else b
            if not b.startswith(("/", "http://", "https://")):
                b = os.path.join(os.path.dirname(filename), b)
            _overlay(cls.load_with_base(b, new_allowed=new_allowed), merged)
        _overlay(node, merged)
        return merged

    def merge_from_file(self, cfg_filename: str) -> None:
        self.merge_from_other_cfg(self._read_file(cfg_filename))

    def merge_from_other_cfg(self, other: Dict[str, Any]) -> None:
        if self.is_frozen():
            raise AttributeError("attempted to merge into a frozen CfgNode")
        _overlay(other, self)

    def merge_from_list(self, cfg_list: Iterable[Any]) -> None:
        items = list(cfg_list)
        if len(items) % 2:
            raise ValueError("override list must have an even number of entries (key value ...)")
        for full_key, raw in zip(items[0::2], items[1::2]):
            node = self
            *parents, leaf = full_key.split(".")
            for p in parents:
                if p not in node:
                    raise KeyError(f"non-existent config key: {full_key}")
                node = node[p]
            if leaf not in node and not node.is_new_allowed():
                raise KeyError(f"non-existent config key: {full_key}")
            node[leaf] = _decode(raw)


def _decode(value):
    if isinstance(value, dict):
        return CfgNode(value)
    if not isinstance(value, str):
        return value
    try:
        return literal_eval(value)
    except (ValueError, SyntaxError):
        return value


def _overlay(src: Dict[str, Any], dst: Dict[str, Any]) -> None:
    """Recursively write `src` over `dst` (dict values merge, everything else replaces)."""
    for k, v in src.items():
        if isinstance(v, dict) and k in dst:
            if not isinstance(dst[k], dict):
                raise AssertionError(f"Cannot inherit key '{k}' from base!")
            _overlay(v, dst[k])
        else:
            dict.__setitem__(dst, k, copy.deepcopy(v) if isinstance(v, dict) else v)
