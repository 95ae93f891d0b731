"""Locate and import the UNMODIFIED reference package `gym_TD`.

TEST INFRASTRUCTURE ONLY (oracle/): used by tests/, oracle/make_golden.py and the
`--impl reference` / `cpu_baseline` legs of bench.py.  Never imported by the
product package `gym_td_b200`.

Search order: <repo>/baseline/_ref (pip --target install, travels to the GPU box),
then /root/reference (only exists in the build container).  Two shims are needed
(SURVEY.md section 8c): the `gym` stub under oracle/gym_stub, and a fake
`numpy.lib.function_base` module because gym_TD/envs/TDDefense.py:6 imports a
name that NumPy 2 no longer has.
"""
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_CANDIDATES = [os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]

_ref = None


def find_reference_root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "gym_TD", "envs", "TDBoard.py")):
            return c
    return None


def available():
    return find_reference_root() is not None


def load():
    """Return the imported reference module `gym_TD` (cached)."""
    global _ref
    if _ref is not None:
        return _ref
    root = find_reference_root()
    if root is None:
        raise ImportError("reference gym_TD not found in %r" % (_CANDIDATES,))
    try:
        import gym  # noqa: F401  (a real gym, if one is ever installed, wins)
    except ImportError:
        sys.path.insert(0, os.path.join(_HERE, "gym_stub"))
    if "numpy.lib.function_base" not in sys.modules:
        try:
            import numpy.lib.function_base  # noqa: F401
        except Exception:
            fake = types.ModuleType("numpy.lib.function_base")
            fake.diff = None
            sys.modules["numpy.lib.function_base"] = fake
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gym_TD  # noqa: E402
        import gym_TD.envs  # noqa: F401,E402
    # keep the reference's debug logger quiet (it prints at level >= its threshold)
    _ref = gym_TD
    return _ref


def modules():
    """Convenience: (gym_TD, TDBoard module, TDElements module, TDParam module)."""
    g = load()
    from gym_TD.envs import TDBoard, TDElements, TDParam, TDRoadGen, TDGymBasic
    return g, TDBoard, TDElements, TDParam, TDRoadGen, TDGymBasic
