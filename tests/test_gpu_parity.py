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


def test_def_multi_action_host_stream():
    """TDDefense(random_agent=False) in allow_multiple_actions mode: Box defender + host-resolved attacker byte."""
    assert PU.run_parity("def", 10, n_envs=16, steps=400, seed=13, multi=True, opponent="stream") > 1000
    assert PU.run_parity("def", 20, n_envs=8, steps=150, seed=14, multi=True, opponent="stream", multi_mode="uniform") > 500


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


@pytest.mark.parametrize("kind,L,multi", [("def", 10, False), ("atk", 10, False), ("2p", 10, False), ("def", 20, True),
                                          ("atk", 20, False), ("2p", 30, False)])
def test_incremental_observation_equals_the_oracle(kind, L, multi):
    """td_step_io.obs_incremental: the observation updated in place (changed planes, old and new tower / enemy
    cells) must be the tensor the oracle builds from scratch, every step."""
    assert PU.run_parity(kind, L, n_envs=16, steps=500, seed=90 + L, multi=multi,
                         opponent="none" if kind == "2p" else "device", incremental=True) > 1500


def test_incremental_observation_with_auto_reset_and_buffer_changes():
    """Two batched envs, one writing full observations and one updating in place, stay bit-identical through
    auto-resets; the library falls back to a full write when it cannot vouch for the buffer."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    for kind, L in (("def", 10), ("atk", 10), ("2p", 20)):
        N = 2048
        full = TDVecEnv(kind, L, N, seed=5, auto_reset=True)
        inc = TDVecEnv(kind, L, N, seed=5, auto_reset=True, incremental_obs=True)
        assert torch.equal(full.reset(), inc.reset())
        g = torch.Generator(device="cuda").manual_seed(3)
        resets = 0
        for k in range(260):
            d = torch.randint(0, 6 * L * L + 1, (N,), dtype=torch.int64, device="cuda", generator=g)
            a = torch.randint(0, 5, (N, 3, 8), dtype=torch.int64, device="cuda", generator=g)
            act = d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}
            o1, r1, d1, _ = full.step(act)
            if k == 100:
                inc.obs.fill_(7.0)                          # someone scribbled over the buffer ...
                inc.engine.observe(inc.obs)                 # ... and asked for the current observation again
            if k == 150:
                torch.cuda.synchronize()
                inc.engine.set_state_raw(inc.engine.get_state_raw())   # td_set_state: the buffer is not vouched for
                inc.obs.zero_()
            o2, r2, d2, _ = inc.step(act)
            assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)), (kind, k)
            assert torch.equal(r1, r2) and torch.equal(d1, d2)
            resets += int(d1.sum().item())
        assert resets > 0                                      # auto-resets happened inside the comparison
        full.close()
        inc.close()
