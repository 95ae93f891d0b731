"""Live differential test: the C oracle against the unmodified reference, step by step (build container
and any box that carries baseline/_ref; skipped elsewhere -- the golden fixtures cover those)."""
import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not present on this box")


def test_reference_self_test_passes_under_the_stub():
    """python -m gym_TD.envs.TDBoard: the reference's own golden check (TDBoard.py:674-756)."""
    import runpy
    import warnings
    ref_loader.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        runpy.run_module("gym_TD.envs.TDBoard", run_name="__main__")     # asserts inside, prints 'passed'


def test_oracle_equals_reference_on_fresh_trajectories():
    from oracle import ref_harness as RH
    from oracle import validate_oracle as V
    ref_loader.load()
    np.seterr(all="ignore")
    rs = np.random.RandomState(2024)
    n = 0
    n += V.run_def(10, 7001, rs, 1, True, max_steps=500)
    n += V.run_def(20, 7002, rs, 0, False, max_steps=300)
    n += V.run_atk(10, 7003, rs, 1, max_steps=500)
    n += V.run_atk(30, 7004, rs, 2, max_steps=200)
    n += V.run_multi(20, 7005, rs, max_steps=300)
    RH.set_multiple_actions(True)
    try:
        n += V.run_def(20, 7006, rs, 1, True, max_steps=200, multi=True)
        n += V.run_multi(10, 7007, rs, max_steps=200, multi=True)
    finally:
        RH.set_multiple_actions(False)
    with RH.ref_config_override(base_LP=None, defender_action_interval=3, attacker_action_interval=2):
        n += V.run_def(10, 7008, rs, 1, True, max_steps=400)
        n += V.run_atk(10, 7009, rs, 1, max_steps=400)
    assert n > 1500
