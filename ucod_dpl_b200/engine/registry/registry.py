"""Name -> object registry (same surface as the reference's engine/registry/registry.py:10-92:
`Registry(name)`, `.register(obj=None)` usable as decorator or call, `.get(name)` raising KeyError,
`in`, iteration over (name, obj) pairs).  Unlike the reference, the hot-path modules of this package
actually register themselves (`baseline`, `RevDecoder`, `Discriminator`, `SparseRefiner`, `backbone`)."""
from __future__ import annotations

from typing import Any, Dict, Iterator, Tuple


class Registry:
    def __init__(self, name: str) -> None:
        self._name = name
        self._table: Dict[str, Any] = {}

    @property
    def name(self) -> str:
        return self._name

    def _add(self, key: str, obj: Any) -> None:
        if key in self._table:
            raise AssertionError(f"An object named '{key}' was already registered in '{self._name}' registry!")
        self._table[key] = obj

    def register(self, obj: Any = None, name: str | None = None) -> Any:
        if obj is None:
            def deco(target: Any) -> Any:
                self._add(name or target.__name__, target)
                return target
            return deco
        self._add(name or obj.__name__, obj)
        return obj

    def get(self, name: str) -> Any:
        try:
            return self._table[name]
        except KeyError:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!") from None

    def build(self, name: str, *args, **kwargs) -> Any:
        """Convenience builder: look `name` up and call it."""
        return self.get(name)(*args, **kwargs)

    def __contains__(self, name: str) -> bool:
        return name in self._table

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        return iter(self._table.items())

    def __len__(self) -> int:
        return len(self._table)

    def __repr__(self) -> str:
        rows = "\n".join(f"  {k}: {v}" for k, v in self._table.items())
        return f"Registry of {self._name}:\n{rows}"

    __str__ = __repr__
