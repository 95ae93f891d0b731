"""register()/make() for the stub: id -> (entry_point, kwargs)."""
import importlib

registry = {}


def register(id, entry_point=None, kwargs=None, max_episode_steps=None, **_):
    registry[id] = dict(entry_point=entry_point, kwargs=dict(kwargs or {}),
                        max_episode_steps=max_episode_steps)


def make(id, **kw):
    spec = registry[id]
    mod, name = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod), name)
    args = dict(spec["kwargs"])
    args.update(kw)
    return cls(**args)
