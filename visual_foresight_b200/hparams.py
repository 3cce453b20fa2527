"""Hyper-parameter container with the semantics the reference relies on from
``tensorflow.contrib.training.HParams`` (reference ``policy/policy.py:4,51-63``,
``cem_base_controller.py:60-76``): explicit registration, typed overwrite, dict export, ``in``."""
from __future__ import annotations

import numbers
from typing import Any, Dict, Iterator


class HParams(object):
    __slots__ = ("_store",)

    def __init__(self, **initial: Any):
        object.__setattr__(self, "_store", {})
        for key, val in initial.items():
            self.add_hparam(key, val)

    # registration ----------------------------------------------------------------------------------
    def add_hparam(self, name: str, value: Any) -> None:
        if name in self._store or hasattr(HParams, name):
            raise ValueError("Hyperparameter name is reserved or already defined: %s" % name)
        self._store[name] = value

    def set_hparam(self, name: str, value: Any) -> None:
        """Overwrite an existing value; numeric kinds must stay compatible (TF raised on e.g. str->int)."""
        if name not in self._store:
            raise KeyError("Hyperparameter %r was never registered" % name)
        old = self._store[name]
        if old is not None and value is not None:
            if isinstance(old, bool) != isinstance(value, bool) and isinstance(old, (bool, numbers.Number)) \
                    and isinstance(value, (bool, numbers.Number)) and isinstance(old, bool):
                raise ValueError("Hyperparameter %r expects a bool" % name)
            if isinstance(old, numbers.Number) and not isinstance(old, bool) and isinstance(value, str):
                raise ValueError("Hyperparameter %r expects a number" % name)
        self._store[name] = value

    # access ------------------------------------------------------------------------------------------
    def get(self, key: str, default: Any = None) -> Any:
        return self._store.get(key, default)

    def values(self) -> Dict[str, Any]:
        return dict(self._store)

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def __iter__(self) -> Iterator[str]:
        return iter(self._store)

    def __getattr__(self, name: str) -> Any:
        store = object.__getattribute__(self, "_store")
        if name in store:
            return store[name]
        raise AttributeError("no hyperparameter %r" % name)

    def __setattr__(self, name: str, value: Any) -> None:
        self._store[name] = value

    def __repr__(self) -> str:
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self._store.items(), key=lambda kv: kv[0]))
