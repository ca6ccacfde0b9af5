"""Minimal stand-in for chex (dataclass + type aliases); see ../README.md."""
import dataclasses as _dc
from typing import Any

Array = Any
ArrayNumpy = Any
ArrayTree = Any
Numeric = Any
PRNGKey = Any
Scalar = Any

_REGISTRY = set()


def is_dataclass_node(obj):
    return type(obj) in _REGISTRY


def dataclass(cls=None, **kwargs):
    kwargs.pop("mappable_dataclass", None)
    kwargs.pop("frozen", None)

    def wrap(c):
        # chex dataclasses take keyword arguments only; dataclasses needs kw_only for defaults before non-defaults
        c = _dc.dataclass(c, kw_only=True, **kwargs)
        _REGISTRY.add(c)

        def replace(self, **changes):
            return _dc.replace(self, **changes)

        c.replace = replace
        return c

    if cls is None:
        return wrap
    return wrap(cls)
