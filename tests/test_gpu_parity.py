"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle, bit for bit, every step."""
import pytest

from tests import parity_util as PU

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [10, 20, 30])
def test_def_discrete_device_opponent(L):
    n = PU.run_parity("def", L, n_envs=48, steps=1200, seed=L, opponent="device", difficulty=1)
    assert n > 10000


@pytest.mark.parametrize("L", [10, 20])
def test_def_discrete_opponent_lv0(L):
    assert PU.run_parity("def", L, n_envs=24, steps=600, seed=L + 1, opponent="device", difficulty=0) > 3000


def test_def_discrete_host_stream():
    assert PU.run_parity("def", 10, n_envs=32, steps=600, seed=3, opponent="stream") > 3000


@pytest.mark.parametrize("L", [10, 20, 30])
def test_atk_device_opponent(L):
    assert PU.run_parity("atk", L, n_envs=32, steps=1200, seed=L + 2, opponent="device", difficulty=1) > 3000


def test_atk_opponent_lv0():
    assert PU.run_parity("atk", 10, n_envs=24, steps=600, seed=5, opponent="device", difficulty=0) > 2000


@pytest.mark.parametrize("L", [10, 20, 30])
def test_atk_opponent_lv2(L):
    assert PU.run_parity("atk", L, n_envs=24, steps=900, seed=L + 6, opponent="device", difficulty=2) > 2000


@pytest.mark.parametrize("L", [10, 20, 30])
def test_multi_discrete(L):
    assert PU.run_parity("2p", L, n_envs=32, steps=1200, seed=L + 3, opponent="none") > 3000


@pytest.mark.parametrize("L", [10, 20])
def test_def_multi_action(L):
    assert PU.run_parity("def", L, n_envs=24, steps=600, seed=L + 4, multi=True, opponent="device") > 2000


def test_def_multi_action_uniform_sample():
    """BASELINE config 3's action distribution: action_space.sample() = uniform {0, 1, 2} on every flag
    (towers are built, upgraded and destructed within the same step almost everywhere)."""
    assert PU.run_parity("def", 20, n_envs=12, steps=300, seed=77, multi=True, opponent="device",
                         multi_mode="uniform") > 1000


def test_multi_multi_action():
    assert PU.run_parity("2p", 20, n_envs=16, steps=600, seed=9, multi=True, opponent="none") > 2000


def test_config_overrides():
    ov = dict(base_LP=None, defender_action_interval=3, attacker_action_interval=2)
    assert PU.run_parity("def", 10, n_envs=16, steps=1200, seed=21, cfg_overrides=ov) > 2000
    assert PU.run_parity("atk", 10, n_envs=16, steps=1200, seed=22, cfg_overrides=ov) > 2000
    ov = dict(defender_init_cost=60, attacker_init_cost=50, defender_cost_rate=.7)
    assert PU.run_parity("2p", 20, n_envs=16, steps=600, seed=23, cfg_overrides=ov, opponent="none") > 2000


def test_odd_map_size_scalar_store_path():
    assert PU.run_parity("def", 15, n_envs=16, steps=400, seed=31) > 1000


def test_dense_boards_exercise_long_lists():
    """Cheap towers and cheap slow enemies: tower lists beyond the 16 staged speculatively (up to ~30 of the 32
    slots), enemy lists into the second 32-lane chunk at L=20, many enemies per cell -- the rarely taken
    multi-pass paths, kept just below the capacities (run_parity asserts that no overflow was flagged)."""
    ov = dict(tower_cost=[[5, 5]] * 4, defender_init_cost=100, enemy_cost=[[4, 4]] * 4, attacker_init_cost=100,
              enemy_speed=[[.05, .05], [.04, .04], [.03, .03], [.03, .03]], base_LP=None)
    assert PU.run_parity("2p", 20, n_envs=12, steps=150, seed=41, opponent="none", cfg_overrides=ov) > 1000
    assert PU.LAST_MAX["towers"] > 16 and PU.LAST_MAX["enemies"] > 32, PU.LAST_MAX
    assert PU.run_parity("atk", 20, n_envs=8, steps=150, seed=43, opponent="device", difficulty=2, cfg_overrides=ov) > 800
    ov10 = dict(ov, enemy_cost=[[7, 7]] * 4)
    assert PU.run_parity("def", 10, n_envs=12, steps=150, seed=42, opponent="device", cfg_overrides=ov10) > 1000
    assert PU.LAST_MAX["towers"] > 16 and PU.LAST_MAX["enemies"] > 16, PU.LAST_MAX


@pytest.mark.parametrize("L,n", [(8, 1), (13, 3), (25, 5), (40, 7), (64, 2)])
def test_unusual_board_sizes_and_batch_sizes(L, n):
    """The run-time-size kernel variant ('TD-def-v0' with a map_size kwarg), batches that do not fill a CTA."""
    assert PU.run_parity("def", L, n_envs=n, steps=250, seed=50 + L, opponent="device") > 100
    assert PU.run_parity("2p", L, n_envs=n, steps=120, seed=60 + L, opponent="none", multi=True) > 50
