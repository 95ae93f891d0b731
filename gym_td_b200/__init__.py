"""gym_td_b200 -- B200-native batched simulator for the gym-TD board step.

Drop-in surface of the reference package `gym_TD` (gym_TD/__init__.py:1-86):
    make(id, **kwargs), the twelve 'TD-{def,atk,2p}-{small,middle,large,}-v0' ids,
    TDDefense / TDAttack / TDMulti, paramConfig / getConfig / getHyperParameters / hyper_parameters,
plus the batched `TDVecEnv` that the throughput numbers are measured on.
Importing this package does not need a GPU; creating an env does (there is no CPU fallback).
"""
from .params import config, getConfig, getHyperParameters, hyper_parameters, paramConfig  # noqa: F401

__version__ = "0.1.0"

# gym_TD/__init__.py:7-86: id -> (entry point, kwargs), max_episode_steps = hyper_parameters.max_episode_steps
REGISTRY = {}
for _prefix, _cls in (("def", "TDDefense"), ("atk", "TDAttack"), ("2p", "TDMulti")):
    for _name, _size in (("small", 10), ("middle", 20), ("large", 30), (None, None)):
        _id = "TD-%s-%s-v0" % (_prefix, _name) if _name else "TD-%s-v0" % _prefix
        REGISTRY[_id] = (_cls, {"map_size": _size} if _size else {})


def make(id, **kwargs):
    """gym.make() for the TD ids: returns the n = 1 façade env (see envs.py)."""
    from . import envs
    cls_name, kw = REGISTRY[id]
    args = dict(kw)
    args.update(kwargs)
    return getattr(envs, cls_name)(**args)


def make_vec(id, num_envs, **kwargs):
    """Batched counterpart: N instances of the env id on one GPU (see vec_env.TDVecEnv)."""
    from .vec_env import TDVecEnv
    cls_name, kw = REGISTRY[id]
    kind = {"TDDefense": "def", "TDAttack": "atk", "TDMulti": "2p"}[cls_name]
    args = dict(kw)
    args.update(kwargs)
    return TDVecEnv(kind, args.pop("map_size"), num_envs, **args)


def register_with_gym():
    """Register the ids with a real `gym` when one is installed (it is not in the build image)."""
    try:
        from gym.envs.registration import register
    except Exception:
        return False
    for _id, (cls_name, kw) in REGISTRY.items():
        try:
            register(id=_id, entry_point="gym_td_b200.envs:%s" % cls_name, kwargs=kw,
                     max_episode_steps=hyper_parameters.max_episode_steps)
        except Exception:
            pass
    return True


def __getattr__(name):
    if name in ("TDDefense", "TDAttack", "TDMulti"):
        from . import envs
        return getattr(envs, name)
    if name == "TDVecEnv":
        from .vec_env import TDVecEnv
        return TDVecEnv
    raise AttributeError(name)
