import importlib

_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, spec = _registry[id]
    mod, cls = entry_point.split(":")
    kw = dict(spec.get("kwargs", {}))
    kw.update(kwargs)
    return getattr(importlib.import_module(mod), cls)(**kw)
