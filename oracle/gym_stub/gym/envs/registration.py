"""gym.envs.registration stand-in - test infrastructure only."""
import importlib

_REGISTRY = {}


def register(id, entry_point, **kwargs):  # noqa: A002
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):  # noqa: A002
    entry_point, defaults = _REGISTRY[id]
    module_name, attr = entry_point.split(':')
    cls = getattr(importlib.import_module(module_name), attr)
    return cls(**{**defaults, **kwargs})
