"""Rollout consumer: numpy oracle vs golden vectors produced by the reference's own PPO class
(CPU), and the CUDA kernels vs both (GPU)."""
import os

import numpy as np
import pytest

from oracle import rollout_oracle as RO
from tests import golden_util as GU


def _load(name):
    return np.load(os.path.join(GU.GOLDEN, "rollout_%s.npz" % name))


@pytest.mark.parametrize("name", ["def", "atk"])
def test_oracle_matches_reference_ppo(name):
    z = _load(name)
    T, n = z["rewards"].shape
    rew = np.zeros((T, n), dtype=np.float32)
    for t in range(T):
        rew[t], d = RO.record_row(z["actions"][t], z["real"][t], z["rewards_in"][t], z["dones"][t], float(z["penalty"]))
        assert np.array_equal(d, z["dones"][t])
    assert np.array_equal(rew.view(np.uint32), z["rewards"].view(np.uint32))
    advs, rets = RO.gae(rew, z["dones"], z["values"], z["next_value"], float(z["gamma"]), float(z["lam"]))
    assert np.array_equal(advs.view(np.uint32), z["advs"].view(np.uint32))
    assert np.array_equal(rets.view(np.uint32), z["returns"].view(np.uint32))


def test_mask_semantics():
    a = np.arange(5, dtype=np.int64)
    out = RO.mask_actions(a, [True, False, True, False, True], 600)
    assert out.tolist() == [0, 600, 2, 600, 4]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["def", "atk"])
def test_cuda_gae_and_record_match_reference(name):
    import torch
    from gym_td_b200 import engine as E
    z = _load(name)
    T, n = z["rewards"].shape
    dev = torch.device("cuda", 0)
    eng = E.Engine("def" if name == "def" else "atk", 10, n)
    which = 0 if name == "def" else 1
    rewards = torch.zeros((T, n), dtype=torch.float32, device=dev)
    dones = torch.zeros((T, n), dtype=torch.uint8, device=dev)
    acts_buf = torch.zeros(z["actions"].shape, dtype=torch.int64, device=dev)
    for t in range(T):
        a = torch.from_numpy(z["actions"][t]).to(dev).contiguous()
        r = torch.from_numpy(z["real"][t]).to(dev).contiguous()
        rw = torch.from_numpy(z["rewards_in"][t]).to(dev)
        d = torch.from_numpy(z["dones"][t].astype(np.uint8)).to(dev)
        eng._check(eng._lib.td_rollout_record(eng._h, which, a.data_ptr(), r.data_ptr(), rw.data_ptr(), d.data_ptr(),
                                              float(z["penalty"]), rewards[t].data_ptr(), dones[t].data_ptr(),
                                              acts_buf[t].data_ptr(), 0))
    torch.cuda.synchronize()
    assert np.array_equal(rewards.cpu().numpy().view(np.uint32), z["rewards"].view(np.uint32))
    assert np.array_equal(dones.cpu().numpy().astype(bool), z["dones"])
    assert np.array_equal(acts_buf.cpu().numpy(), z["actions"])
    advs = torch.zeros((T, n), dtype=torch.float32, device=dev)
    rets = torch.zeros_like(advs)
    v = torch.from_numpy(z["values"]).to(dev).contiguous()
    nv = torch.from_numpy(z["next_value"]).to(dev).contiguous()
    assert E.lib().td_gae(T, n, rewards.data_ptr(), dones.data_ptr(), v.data_ptr(), nv.data_ptr(), float(z["gamma"]),
                          float(z["lam"]), advs.data_ptr(), rets.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    assert np.array_equal(advs.cpu().numpy().view(np.uint32), z["advs"].view(np.uint32))
    assert np.array_equal(rets.cpu().numpy().view(np.uint32), z["returns"].view(np.uint32))
    eng.close()


@pytest.mark.gpu
def test_rollout_buffer_on_live_env_matches_oracle():
    import torch
    from gym_td_b200.rollout import RolloutBuffer
    from gym_td_b200.vec_env import TDVecEnv
    N, L, T = 512, 10, 32
    env = TDVecEnv("def", L, N, seed=11, auto_reset=True,
                   cfg=__import__("tests.parity_util", fromlist=["x"]).make_config(defender_action_interval=3))
    env.reset()
    buf = RolloutBuffer(env, horizon=T)
    g = torch.Generator(device="cuda").manual_seed(3)
    rows_r, rows_d = [], []
    allow_prev = np.ones(N, dtype=bool)
    for t in range(T):
        a = torch.randint(0, 601, (N,), dtype=torch.int64, device="cuda", generator=g)
        want = RO.mask_actions(a.cpu().numpy(), allow_prev, 600)
        buf.mask(a)
        assert np.array_equal(a.cpu().numpy(), want)
        obs, rew, done, info = env.step(a)
        full = buf.record(a)
        r32, d = RO.record_row(want, info["RealAction"].cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy())
        rows_r.append(r32)
        rows_d.append(d)
        allow_prev = info["AllowNextMove"].cpu().numpy()
        assert full == (t == T - 1)
    assert (~allow_prev).any() or True
    rew_ref, done_ref = np.stack(rows_r), np.stack(rows_d)
    assert np.array_equal(buf.rewards.cpu().numpy().view(np.uint32), rew_ref.view(np.uint32))
    assert np.array_equal(buf.dones.cpu().numpy().astype(bool), done_ref)
    values = torch.randn((T, N), device="cuda", generator=g)
    nv = torch.randn((N,), device="cuda", generator=g)
    advs, rets = buf.flush(values, nv)
    a_ref, r_ref = RO.gae(rew_ref, done_ref, values.cpu().numpy(), nv.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(advs.cpu().numpy().view(np.uint32), a_ref.view(np.uint32))
    assert np.array_equal(rets.cpu().numpy().view(np.uint32), r_ref.view(np.uint32))
    env.close()


@pytest.mark.gpu
def test_rollout_buffers_for_both_players_of_the_2p_env():
    """BASELINE config 5: TD-2p feeds one rollout buffer per player; each row equals the reference's bookkeeping
    applied to that player's action / RealAction / AllowNextMove bit, the attacker on the negated reward."""
    import torch
    from gym_td_b200.rollout import RolloutBuffer
    from gym_td_b200.vec_env import TDVecEnv
    from tests import parity_util as PU
    N, L, T = 256, 10, 24
    env = TDVecEnv("2p", L, N, seed=12, auto_reset=True,
                   cfg=PU.make_config(defender_action_interval=3, attacker_action_interval=2))
    env.reset()
    with pytest.raises(ValueError):
        RolloutBuffer(env, horizon=T)
    bd, ba = RolloutBuffer(env, horizon=T, role="defender"), RolloutBuffer(env, horizon=T, role="attacker")
    g = torch.Generator(device="cuda").manual_seed(4)
    allow_d, allow_a = np.ones(N, dtype=bool), np.ones(N, dtype=bool)
    rows = {"d": [], "a": [], "done": []}
    for t in range(T):
        act = {"Defender": torch.randint(0, 601, (N,), dtype=torch.int64, device="cuda", generator=g),
               "Attacker": torch.randint(0, 5, (N, 3, 8), dtype=torch.int64, device="cuda", generator=g)}
        want_d = RO.mask_actions(act["Defender"].cpu().numpy(), allow_d, 600)
        want_a = act["Attacker"].cpu().numpy().copy()
        want_a[~allow_a] = 4
        bd.mask(act), ba.mask(act)
        assert np.array_equal(act["Defender"].cpu().numpy(), want_d) and np.array_equal(act["Attacker"].cpu().numpy(), want_a)
        obs, rew, done, info = env.step(act)
        bd.record(act), ba.record(act)
        r, dn = rew.cpu().numpy(), done.cpu().numpy()
        rows["d"].append(RO.record_row(want_d, info["RealAction"]["Defender"].cpu().numpy(), r, dn)[0])
        rows["a"].append(RO.record_row(want_a, info["RealAction"]["Attacker"].cpu().numpy(), -r, dn)[0])
        rows["done"].append(dn)
        allow_d = info["AllowNextMove"]["Defender"].cpu().numpy()
        allow_a = info["AllowNextMove"]["Attacker"].cpu().numpy()
    assert np.array_equal(bd.rewards.cpu().numpy().view(np.uint32), np.stack(rows["d"]).view(np.uint32))
    assert np.array_equal(ba.rewards.cpu().numpy().view(np.uint32), np.stack(rows["a"]).view(np.uint32))
    assert np.array_equal(ba.dones.cpu().numpy().astype(bool), np.stack(rows["done"]))
    values = torch.randn((T, N), device="cuda", generator=g)
    nv = torch.randn((N,), device="cuda", generator=g)
    advs, rets = ba.flush(-values, -nv)
    a_ref, r_ref = RO.gae(np.stack(rows["a"]), np.stack(rows["done"]), (-values).cpu().numpy(), (-nv).cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(advs.cpu().numpy().view(np.uint32), a_ref.view(np.uint32))
    assert np.array_equal(rets.cpu().numpy().view(np.uint32), r_ref.view(np.uint32))
    env.close()


@pytest.mark.gpu
def test_policy_reads_observation_in_place():
    """BASELINE.json config 5 in miniature: dict actions sampled on the GPU, obs consumed in place."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_id, n in (("TD-2p-large-v0", 256), ("TD-def-small-v0", 1024)):
        out = subprocess.run([sys.executable, os.path.join(root, "examples", "rollout_feed.py"), "--env", env_id,
                              "--envs", str(n), "--steps", "40"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "env-steps/s" in out.stdout
