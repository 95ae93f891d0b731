"""GPU tests of the host layer: drop-in façade envs, batched env semantics, statistics, state access,
capacity flagging and full-size properties (BASELINE.json sizes)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from tests import golden_util as GU
from tests import parity_util as PU

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,L", [("def", 10), ("2p", 10), ("atk", 10), ("def", 20), ("atk", 30)])
def test_facade_reproduces_survey_known_answers(kind, L):
    """The script that produced SURVEY.md 8(c)'s digests from the reference, run on this package instead."""
    import gym_td_b200 as G
    kats = {(k["kind"], k["L"]): k for k in json.load(open(os.path.join(GU.GOLDEN, "survey_kats.json")))}
    k = kats[(kind, L)]
    rs = np.random.RandomState(0)
    if kind == "def":
        env = G.TDDefense(L, seed=1024, random_agent=False)
    elif kind == "2p":
        env = G.TDMulti(L, seed=1024, random_agent=False)
    else:
        random.seed(1024)
        env = G.TDAttack(L, seed=1024, random_agent=True)
    assert env.observation_space.shape == (45, L, L) and env.num_roads in (1, 2, 3)
    h = hashlib.sha256()
    h.update(env._board.get_states().tobytes())
    ret, n = 0.0, 0
    while True:
        if kind == "def":
            a = int(rs.randint(6 * L * L + 1))
        elif kind == "2p":
            atk = rs.randint(0, 5, size=(3, 8))
            a = {"Attacker": atk, "Defender": int(rs.randint(6 * L * L + 1))}
        else:
            a = rs.randint(0, 5, size=(3, 8))
        o, r, d, info = env.step(a)
        assert o.dtype == np.float32 and isinstance(r, float) and isinstance(d, bool)
        assert set(info) == {"RealAction", "Win", "AllowNextMove", "FailCode"}
        h.update(o.tobytes())
        h.update(np.float64(r).tobytes())
        ret += r
        n += 1
        if d:
            break
    assert (n, repr(ret), h.hexdigest()[:16]) == (k["steps"], k["ret"], k["sha"])
    assert info["Win"] is not None
    env.close()


def test_facade_against_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    import gym_td_b200 as G
    from oracle import ref_harness as RH
    ref_loader.load()
    np.seterr(all="ignore")
    for kind, L, seed in (("def", 10, 1024), ("atk", 10, 1024), ("2p", 20, 1024)):
        ref = RH.make_env(kind, L, seed)
        assert ref is not None
        random.seed(seed)
        st0 = random.getstate()
        mine = G.make("TD-%s-%s-v0" % (kind, {10: "small", 20: "middle"}[L]), seed=seed)
        assert np.array_equal(ref._board.get_states(), mine._board.get_states())
        assert mine.num_roads == ref.num_roads
        rs = np.random.RandomState(3)
        st_ref = st_mine = st0
        for t in range(250):
            if kind == "atk":
                a = rs.randint(0, 5, size=(3, 8))
            elif kind == "def":
                a = RH.smart_defender_action(ref._board, rs)
            else:
                a = {"Attacker": rs.randint(0, 5, size=(3, 8)), "Defender": RH.smart_defender_action(ref._board, rs)}
            random.setstate(st_ref)
            o1, r1, d1, i1 = ref.step(a)
            st_ref = random.getstate()
            random.setstate(st_mine)
            o2, r2, d2, i2 = mine.step(a)
            st_mine = random.getstate()
            assert st_ref == st_mine                         # the global `random` stream advanced identically
            assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)) and repr(r1) == repr(r2) and d1 == d2
            assert type(i1["RealAction"]) == type(i2["RealAction"]) or kind != "2p"
            assert repr(i1["Win"]) == repr(i2["Win"]) and i1["AllowNextMove"] == i2["AllowNextMove"]
            if isinstance(i1["RealAction"], dict):
                assert np.array_equal(i1["RealAction"]["Attacker"], i2["RealAction"]["Attacker"])
                assert i1["RealAction"]["Defender"] == i2["RealAction"]["Defender"]
            else:
                assert np.array_equal(np.asarray(i1["RealAction"]), np.asarray(i2["RealAction"]))
            assert json.dumps(i1["FailCode"], default=int) == json.dumps(i2["FailCode"], default=int)
            b1, b2 = ref._board, mine._board
            assert repr(float(b1.cost_def)) == repr(b2.cost_def) and len(b1.towers) == len(b2.towers)
            assert [(e.loc, int(e.type), float(e.LP)) for e in b1.enemies] == [(e.loc, e.type, e.LP) for e in b2.enemies]
            if d1:
                break
        o1 = ref.reset()                                      # second draw from the same np_random stream
        o2 = mine.reset()
        if o1 is not None:
            assert np.array_equal(o1, o2)
        mine.close()


@pytest.mark.parametrize("kind,difficulty", [("def", 0), ("def", 1), ("atk", 0)])
def test_facade_np_random_opponents_against_live_reference(kind, difficulty):
    """random_agent=False: the scripted opponent draws from the env's own np_random (TDGymBasic.py:87-89, 102-103,
    117-119).  The facade resolves the draws on the host from the live RandomState and hands them to the kernel;
    observations, rewards, info and the np_random stream itself must follow the reference step by step."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    import gym_td_b200 as G
    from oracle import ref_harness as RH
    ref_loader.load()
    np.seterr(all="ignore")
    seed, ref = RH.first_valid_seed(kind, 10, 2040 + difficulty, difficulty=difficulty, random_agent=False)
    cls = G.TDDefense if kind == "def" else G.TDAttack
    mine = cls(10, difficulty=difficulty, seed=seed, random_agent=False)
    assert np.array_equal(ref._board.get_states(), mine._board.get_states())
    rs = np.random.RandomState(6)
    for t in range(400):
        a = RH.smart_defender_action(ref._board, rs) if kind == "def" else rs.randint(0, 5, size=(3, 8))
        o1, r1, d1, i1 = ref.step(a)
        o2, r2, d2, i2 = mine.step(a)
        assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)) and repr(r1) == repr(r2) and d1 == d2, t
        assert np.array_equal(np.asarray(i1["RealAction"]), np.asarray(i2["RealAction"]))
        assert json.dumps(i1["FailCode"], default=int) == json.dumps(i2["FailCode"], default=int)
        assert i1["AllowNextMove"] == i2["AllowNextMove"] and repr(i1["Win"]) == repr(i2["Win"])
        s1, s2 = ref.np_random.get_state(), mine.np_random.get_state()
        assert s1[2] == s2[2] and np.array_equal(s1[1], s2[1]), "np_random streams diverged at step %d" % t
        if d1:
            break
    assert t > 50
    mine.close()


def test_facade_refuses_what_the_reference_cannot_run():
    import gym_td_b200 as G
    for d in (1, 2):
        with pytest.raises(NotImplementedError):
            G.TDAttack(10, difficulty=d, seed=3, random_agent=False)
    with pytest.raises(AttributeError):
        G.TDDefense(10, difficulty=2, seed=3)


def test_invalid_action_asserts_like_the_reference():
    import gym_td_b200 as G
    env = G.make("TD-def-small-v0", seed=5)
    with pytest.raises(AssertionError):
        env.step(601)
    with pytest.raises(AssertionError):
        env.step(1.5)
    assert env.empty_action() == 600
    env.close()
    atk = G.make("TD-atk-small-v0", seed=5)
    with pytest.raises(AssertionError):
        atk.step(np.full((3, 8), 5))
    assert atk.empty_action().shape == (3, 8)
    atk.close()


def test_vec_env_auto_reset_statistics_and_host_path():
    """TDVecEnv over several episodes vs per-env oracles restarted on the next map; statistics; step_host."""
    import torch
    from gym_td_b200 import mapgen
    from gym_td_b200.vec_env import TDVecEnv
    from oracle import td_oracle as TO
    N, L = 24, 10
    cfg = PU.make_config(max_episode_steps=150)
    env = TDVecEnv("def", L, N, seed=400, auto_reset=True, cfg=cfg)
    ocfg = PU.oracle_config_from_engine(cfg)
    maps, _, _ = mapgen.generate_batch((np.arange(N) + 400).astype(np.uint32), L)
    oracles, map_id = [], list(range(N))
    for i in range(N):
        o = PU.oracle_env_from_map(maps[i], ocfg)
        o.set_pyrand(random.Random(400 + i).getstate())
        oracles.append(o)
    env.reset()
    rs = np.random.RandomState(9)
    tot = dict(episodes=0, ret=0.0, length=0, wins=0, kills=0, leaks=0)
    ep_ret = [0.0] * N
    acts_pinned = torch.empty(N, dtype=torch.int64).pin_memory()
    for step in range(400):
        a = np.array([PU.smart_defender_action(o, rs) for o in oracles], dtype=np.int64)
        if step % 2 == 0:
            obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
            torch.cuda.synchronize()
            rew_h, done_h = rew.cpu().numpy(), done.cpu().numpy()
        else:                                                  # same step through the host-buffer ABI call
            acts_pinned.copy_(torch.from_numpy(a))
            h = env.step_host(acts_pinned)
            rew_h, done_h = h["reward"].numpy().copy(), h["done"].numpy().astype(bool)
        obs_h = env.obs.cpu().numpy()
        for i, o in enumerate(oracles):
            out = o.def_step(int(a[i]), 1, False)
            assert float(rew_h[i]).hex() == float(out.reward).hex() and bool(done_h[i]) == bool(out.done), (step, i)
            ep_ret[i] += out.reward
            tot["kills"] += o.e.last_kills
            tot["leaks"] += o.e.last_leaks
            if out.done:
                tot["episodes"] += 1
                tot["ret"] += ep_ret[i]
                tot["length"] += o.e.steps
                tot["wins"] += int(out.win)
                ep_ret[i] = 0.0
                map_id[i] = (map_id[i] + 1) % N                # map_stride = 1
                py = o.e.pyrand                                # the opponent stream keeps running across episodes
                oracles[i] = o = PU.oracle_env_from_map(maps[map_id[i]], ocfg)
                o.e.pyrand = py
            assert np.array_equal(obs_h[i].view(np.uint32), o.get_states().view(np.uint32)), (step, i)
    st = env.stats()
    assert st["episodes"] == tot["episodes"] > N and st["length_sum"] == tot["length"] and st["wins"] == tot["wins"]
    assert st["steps"] == 400 * N and st["overflow_envs"] == 0
    assert abs(st["return_sum"] - tot["ret"]) < 1e-9 * max(1.0, abs(tot["ret"]))
    # kills/leaks are accumulated per finished episode on the device
    assert st["kills"] <= tot["kills"] and st["leaks"] <= tot["leaks"]
    assert env.allreduce_stats()["episodes"] == st["episodes"]
    env.close()


def test_observe_and_state_round_trip():
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    env = TDVecEnv("2p", 20, 8, seed=3, auto_reset=False, scripted_opponent=False)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(120):
        a = {"Attacker": torch.randint(0, 5, (8, 3, 8), device="cuda", generator=g),
             "Defender": torch.randint(0, 2401, (8,), device="cuda", generator=g)}
        env.step(a)
    ref_obs = env.obs.clone()
    out = torch.zeros_like(ref_obs)
    env.engine.observe(out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), ref_obs.view(torch.int32))       # kernel (f) alone == fused path
    blob = env.engine.get_state_raw()
    blob2 = blob.copy()
    blob2[1] = blob[0]                                                          # clone env 0 into env 1
    hdr = blob2[1][:64].view(np.uint8)
    env.engine.set_state_raw(blob2)
    env.engine.observe(out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out[1].view(torch.int32), ref_obs[0].view(torch.int32))
    bad = blob.copy()
    bad[0][44] = 200                                                            # n_towers > capacity
    with pytest.raises(Exception):
        env.engine.set_state_raw(bad)
    env.close()


def test_capacity_overflow_is_flagged_not_silent():
    import torch
    from gym_td_b200 import engine as E
    from gym_td_b200.vec_env import TDVecEnv
    cheap = [[1, 1]] * 4
    cfg = PU.make_config(enemy_cost=cheap, attacker_init_cost=100, base_LP=None,
                         enemy_speed=[[.01, .01]] * 4)
    env = TDVecEnv("atk", 10, 4, seed=1, auto_reset=False, scripted_opponent=False, cfg=cfg)
    env.reset()
    a = torch.zeros((4, 3, 8), dtype=torch.int64, device="cuda")
    for _ in range(12):
        env.step(a)
    with pytest.raises(E.TdError):
        env.stats()
    env.close()


def test_capacity_overflow_raises_in_the_facade_too():
    """the n = 1 facades read the sticky overflow flags after every step."""
    import gym_td_b200 as G
    from gym_td_b200 import engine as E
    kw = dict(enemy_cost=[[1, 1]] * 4, attacker_init_cost=100, base_LP=None, enemy_speed=[[.01, .01]] * 4)
    old = {k: getattr(G.config, k) for k in kw}
    G.paramConfig(**kw)
    try:
        env = G.make("TD-atk-small-v0", seed=1)
        with pytest.raises(E.TdError):
            for _ in range(12):
                env.step(np.zeros((3, 8), dtype=np.int64))
        env.close()
    finally:
        G.paramConfig(**old)


def test_rebinding_obs_restarts_the_incremental_update():
    """the in-place observation update trusts a buffer by address; TDVecEnv drops that trust whenever
    `obs` is rebound, also to memory at the very same address."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    N, L = 64, 10
    a = TDVecEnv("def", L, N, seed=3, auto_reset=True, incremental_obs=True)
    b = TDVecEnv("def", L, N, seed=3, auto_reset=True)
    a.reset(), b.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(60):
        act = torch.randint(0, 601, (N,), device="cuda", generator=g)
        if t == 20:
            a.obs.fill_(7.0)                    # somebody scribbled over the buffer ...
            a.obs = a.obs                       # ... and says so by rebinding (same address)
        if t == 40:
            a.obs.zero_()
            a.invalidate_obs()
        o1, _, _, _ = a.step(act)
        o2, _, _, _ = b.step(act)
        assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)), t
    a.close(), b.close()


@pytest.mark.parametrize("kind,L", [("def", 10), ("atk", 10), ("2p", 20), ("def", 30)])
def test_reduced_precision_observation_planes(kind, L):
    """the opt-in bfloat16 / uint8 observation tensors are exactly the float32 observation rounded
    (bf16: nearest-even) resp. quantised (u8: rint(min(255 v, 255))), every step, through resets and td_observe."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    N = 96
    envs = {f: TDVecEnv(kind, L, N, seed=8, auto_reset=True, obs_format=f) for f in ("f32", "bf16", "u8")}
    for e in envs.values():
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(2)

    def check(tag):
        ref = envs["f32"].obs
        assert torch.equal(envs["bf16"].obs.view(torch.int16), ref.to(torch.bfloat16).view(torch.int16)), tag
        want = torch.clamp(torch.round(ref * 255.0), 0, 255).to(torch.uint8)
        assert torch.equal(envs["u8"].obs, want), tag
        assert envs["bf16"].obs.dtype == torch.bfloat16 and envs["u8"].obs.element_size() == 1

    check("reset")
    for t in range(150):
        d = torch.randint(0, 6 * L * L + 1, (N,), dtype=torch.int64, device="cuda", generator=g)
        a = torch.randint(0, 5, (N, 3, 8), dtype=torch.int64, device="cuda", generator=g)
        act = d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}
        outs = {f: e.step(act) for f, e in envs.items()}
        assert torch.equal(outs["f32"][1].view(torch.int64), outs["bf16"][1].view(torch.int64))
        assert torch.equal(outs["f32"][2], outs["u8"][2])
        check((kind, L, t))
    assert (envs["f32"].obs > 0).any() and (envs["u8"].obs == 255).any()
    snap = envs["bf16"].state_dict()
    envs["bf16"].load_state_dict(snap)                        # td_observe_as path
    check("observe")
    with pytest.raises(ValueError):
        TDVecEnv("def", 12, 4, obs_format="bf16")
    for e in envs.values():
        e.close()


def test_attacker_action_at_an_unaligned_address():
    """The (3, 8) int64 attacker action normally rides in with the record as 16-byte asynchronous copies; a tensor
    that starts 8 bytes off a 16-byte boundary takes the plain-load path.  Same results either way."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    N = 257
    envs = [TDVecEnv(k, 10, N, seed=4, auto_reset=True) for k in ("atk", "atk", "2p", "2p")]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(6)
    raw = torch.zeros(N * 24 + 1, dtype=torch.int64, device="cuda")
    off = raw[1:].view(N, 3, 8)
    assert off.data_ptr() % 16 == 8 and off.is_contiguous()
    for t in range(80):
        a = torch.randint(0, 5, (N, 3, 8), dtype=torch.int64, device="cuda", generator=g)
        d = torch.randint(0, 601, (N,), dtype=torch.int64, device="cuda", generator=g)
        off.copy_(a)
        o1, r1, d1, i1 = envs[0].step(a)
        o2, r2, d2, i2 = envs[1].step(off)
        assert torch.equal(o1.view(torch.int32), o2.view(torch.int32)) and torch.equal(r1.view(torch.int64), r2.view(torch.int64))
        assert torch.equal(i1["RealAction"], i2["RealAction"]) and torch.equal(i1["FailCode"], i2["FailCode"])
        o3, r3, d3, i3 = envs[2].step({"Attacker": a, "Defender": d})
        o4, r4, d4, i4 = envs[3].step({"Attacker": off, "Defender": d})
        assert torch.equal(o3.view(torch.int32), o4.view(torch.int32)) and torch.equal(r3.view(torch.int64), r4.view(torch.int64))
        assert torch.equal(i3["RealAction"]["Attacker"], i4["RealAction"]["Attacker"])
    for e in envs:
        e.close()


def test_full_size_determinism_and_replayed_subset():
    """BASELINE.json config 2: 65,536 envs, Discrete actions; the first 256 envs replayed through the oracle;
    size-independent properties on the whole batch (broadcast planes equal the state, determinism)."""
    import torch
    from gym_td_b200 import mapgen
    from gym_td_b200.vec_env import TDVecEnv
    N, L, K, SUB = 65536, 10, 120, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    acts = torch.randint(0, 601, (K, N), dtype=torch.int64, device="cuda", generator=g)

    def run():
        env = TDVecEnv("def", L, N, seed=0, auto_reset=True)
        env.reset()
        sums, subs = [], []
        for k in range(K):
            obs, rew, done, info = env.step(acts[k])
            sums.append((obs.view(torch.int32).sum(dtype=torch.int64).item(), rew.sum().item(), done.sum().item()))
            subs.append((obs[:SUB].cpu().numpy(), rew[:SUB].cpu().numpy(), done[:SUB].cpu().numpy()))
        blob = env.engine.get_state_raw(0, 2048)
        hdr = np.stack([b[:64].view(E.HEADER_DTYPE)[0] for b in blob])
        o = env.obs[:2048].cpu().numpy()
        env.close()
        return sums, subs, hdr, o

    from gym_td_b200 import engine as E
    s1, sub1, hdr, o = run()
    s2, _, _, _ = run()
    assert s1 == s2                                                    # bitwise deterministic at full size
    # broadcast planes are pure functions of the record header
    assert np.array_equal(o[:, 11, 0, 0], (hdr["cost_def"] / 100.0).astype(np.float32))
    assert np.array_equal(o[:, 12, 3, 4], (hdr["cost_atk"] / 100.0).astype(np.float32))
    assert np.array_equal(o[:, 13, 9, 9], (hdr["steps"] / 1200).astype(np.float32))
    assert np.array_equal(o[:, 5, 2, 2], (hdr["base_LP"] / 5).astype(np.float32))
    assert (o[:, 10] == 0).all() and (o[:, 11].min(axis=(1, 2)) == o[:, 11].max(axis=(1, 2))).all()
    # replayed subset
    maps, _, _ = mapgen.generate_batch(np.arange(N, dtype=np.uint32), L)
    ocfg = PU.oracle_config_from_engine(PU.make_config())
    acts_h = acts[:, :SUB].cpu().numpy()
    for i in range(SUB):
        oenv = PU.oracle_env_from_map(maps[i], ocfg)
        oenv.set_pyrand(random.Random(i).getstate())
        mid = i
        for k in range(K):
            out = oenv.def_step(int(acts_h[k, i]), 1, False)
            assert float(sub1[k][1][i]).hex() == float(out.reward).hex() and bool(sub1[k][2][i]) == bool(out.done), (i, k)
            if out.done:
                py = oenv.e.pyrand
                mid = (mid + 1) % N
                oenv = PU.oracle_env_from_map(maps[mid], ocfg)
                oenv.e.pyrand = py
            assert np.array_equal(sub1[k][0][i].view(np.uint32), oenv.get_states().view(np.uint32)), (i, k)


@pytest.mark.parametrize("kind,L,N,multi", [("atk", 10, 65536, False), ("def", 20, 32768, True), ("2p", 30, 16384, False)])
def test_full_size_properties_of_the_other_baseline_configs(kind, L, N, multi):
    """BASELINE.json configs 3-5 at their full per-GPU sizes: bitwise determinism of two runs (a checksum of
    the per-step checksums) and size-independent invariants tying every observation to its env record."""
    import torch
    from gym_td_b200 import engine as E
    from gym_td_b200.vec_env import TDVecEnv
    K = 30
    g = torch.Generator(device="cuda").manual_seed(11)
    a_atk = torch.randint(0, 5, (4, N, 3, 8), dtype=torch.int64, device="cuda", generator=g) if kind != "def" else None
    if kind == "atk":
        a_def = None
    elif multi:
        a_def = (torch.rand((2, N, 6, L, L), device="cuda", generator=g) < 0.01).to(torch.int64)
    else:
        a_def = torch.randint(0, 6 * L * L + 1, (4, N), dtype=torch.int64, device="cuda", generator=g)

    def action(k):
        d = a_def[k % a_def.shape[0]] if a_def is not None else None
        a = a_atk[k % 4] if a_atk is not None else None
        return d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}

    def run():
        env = TDVecEnv(kind, L, N, seed=3, auto_reset=True, multi_action=multi)
        env.reset()
        digest = 0
        for k in range(K):
            obs, rew, done, info = env.step(action(k))
            digest = (digest * 1000003 + obs.view(torch.int32).sum(dtype=torch.int64).item()
                      + 31 * rew.view(torch.int64).sum().item() + 7 * int(done.sum().item())) % (1 << 61)
        M = 1024
        blob = env.engine.get_state_raw(0, M)
        hdr = np.stack([b[:64].view(E.HEADER_DTYPE)[0] for b in blob])
        o = env.obs[:M].cpu().numpy()
        flags = env.stats()
        env.close()
        return digest, hdr, o, flags

    d1, hdr, o, st = run()
    d2, _, _, _ = run()
    assert d1 == d2
    assert np.array_equal(o[:, 11, 0, 0], (hdr["cost_def"] / 100.0).astype(np.float32))
    assert np.array_equal(o[:, 12, L - 1, 0], (hdr["cost_atk"] / 100.0).astype(np.float32))
    assert np.array_equal(o[:, 13, 0, L - 1], (hdr["steps"] / 1200).astype(np.float32))
    assert np.array_equal(o[:, 5, 1, 1], (hdr["base_LP"] / 5).astype(np.float32))
    assert (o[:, 10] == 0).all()
    # one-hot planes against the list lengths of the record: towers by level (15-16) and by type (17-20),
    # enemies through the count planes (37-40 hold count / 8 per cell and type)
    assert np.array_equal(o[:, 15:17].sum(axis=(1, 2, 3)), hdr["n_towers"].astype(np.float32))
    assert np.array_equal(o[:, 17:21].sum(axis=(1, 2, 3)), hdr["n_towers"].astype(np.float32))
    assert np.array_equal(o[:, 37:41].sum(axis=(1, 2, 3)) * 8, hdr["n_enemies"].astype(np.float32))
    assert (o[:, 4].sum(axis=(1, 2)) == 1).all() and (o[:, 6:9].sum(axis=(1, 2, 3)) >= 1).all()
    assert (o[:, 25:29] <= o[:, 29:33]).all()                         # min ratio <= max ratio wherever enemies stand


@pytest.mark.parametrize("kind", ["def", "atk"])
def test_checkpoint_resume_is_bit_identical(kind):
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    N, L = 64, 10
    env = TDVecEnv(kind, L, N, seed=77, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    def act():
        return (torch.randint(0, 601, (N,), device="cuda", generator=g) if kind == "def"
                else torch.randint(0, 5, (N, 3, 8), device="cuda", generator=g))
    for _ in range(150):
        env.step(act())
    snap = env.state_dict()
    acts = [act() for _ in range(200)]
    def run():
        out = []
        for a in acts:
            obs, rew, done, info = env.step(a)
            out.append((obs.view(torch.int32).sum(dtype=torch.int64).item(), rew.clone(), done.clone()))
        return out
    first = run()
    obs0 = env.load_state_dict(snap)
    second = run()
    for (s1, r1, d1), (s2, r2, d2) in zip(first, second):
        assert s1 == s2 and torch.equal(r1.view(torch.int64), r2.view(torch.int64)) and torch.equal(d1, d2)
    env.close()


@pytest.mark.parametrize("chunks,graph,extra", [("1", "1", []), ("3", "1", []), ("4", "1", ["1"]), ("0", "1", []), ("3", "0", []),
                                                ("4", "1", ["0", "16"]), ("3", "1", ["1", "100"]),
                                                ("3", "1", ["0", "-1", "0"]), ("0", "1", ["0", "-1", "1"]), ("2", "0", ["0", "-1", "1"]),
                                                ("0", "1", ["0", "-1", "-1", "pageable"])])
def test_chunked_host_path_equals_device_path(chunks, graph, extra):
    """td_step_host cuts the batch into chunks inside one CUDA graph -- independent branches (default) or chained
    kernels, equal chunks or a short first chunk followed by growing ones -- or issues plain stream launches; small
    actions are read by the kernel straight from the page-locked host buffer (zero-copy: automatic / never / every
    action; ordinary pageable action buffers fall back to stream copies).  Results must not depend on any of it, nor on
    which cached graph serves a call."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "host_chunk_check.py"), chunks, graph] + extra,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "chunked host path ok" in out.stdout


def test_observations_rebuilt_from_compact_snapshots():
    """f4: keep 2.5 KB records instead of 18 KB observations, rebuild any observation later, bit-identical."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    N, L, T = 96, 10, 40
    env = TDVecEnv("def", L, N, seed=5, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    rb = env.engine.layout.record_bytes
    store = torch.empty((T, N, rb), dtype=torch.uint8, device="cuda")
    obs_ref = []
    for t in range(T):
        obs, _, _, _ = env.step(torch.randint(0, 601, (N,), device="cuda", generator=g))
        env.snapshot(store[t])
        obs_ref.append(obs.clone())
    assert rb * 7 < 45 * L * L * 4
    picks = torch.tensor([0, 7, 39, 22, 13], device="cuda")
    rebuilt = env.observe_snapshot(store[picks].contiguous())
    rebuilt = rebuilt.view(len(picks), N, 45, L, L)
    for k, t in enumerate(picks.tolist()):
        assert torch.equal(rebuilt[k].view(torch.int32), obs_ref[t].view(torch.int32)), t
    env.close()


def test_param_config_reaches_the_device_like_the_reference():
    """paramConfig(**kw) before make(): same observable behaviour as the reference with the same config."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    import gym_td_b200 as G
    from oracle import ref_harness as RH
    ref_loader.load()
    np.seterr(all="ignore")
    kw = dict(base_LP=None, defender_init_cost=45, attacker_init_cost=30, max_cost=80, reward_kill=0.25,
              tower_destruct_return=0.75, frozen_time=3)
    old = {k: getattr(G.config, k) for k in kw}
    G.paramConfig(**kw)
    try:
        with RH.ref_config_override(**kw):
            ref = RH.make_env("2p", 10, 77)
            mine = G.make("TD-2p-small-v0", seed=77)
            assert np.array_equal(ref._board.get_states(), mine._board.get_states())
            rs = np.random.RandomState(8)
            for t in range(300):
                a = {"Attacker": rs.randint(0, 5, size=(3, 8)), "Defender": RH.smart_defender_action(ref._board, rs)}
                o1, r1, d1, i1 = ref.step(a)
                o2, r2, d2, i2 = mine.step(a)
                assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)) and repr(r1) == repr(r2) and d1 == d2, t
                assert repr(i1["Win"]) == repr(i2["Win"])
                if d1:
                    break
            assert t == 299 or d1          # base_LP=None: only the step limit ends the episode
            mine.close()
    finally:
        G.paramConfig(**old)


def _round_road_episode(env, t, num_enemy):
    """The caller loop of the reference's balance script (balance.py:95-121), one episode."""
    done, mem, road, total = False, None, 0, 0.0
    while not done:
        if mem is not None:
            act = mem
        else:
            act = np.full((3, 8), 4, np.int64)
            act[road, :num_enemy] = t
            road = road + 1 if road + 1 < env.num_roads else 0
        _, r, done, info = env.step(act)
        mem = act if 2 in info["FailCode"] else None            # FC.COST_SHORTAGE
        total += r
    return info["Win"], total


def test_balance_scripts_through_facade_and_batched_agent():
    """the balance.py attack scripts. The single-env loop gives the same episodes through the
    reference env and the façade; the batched agent reproduces the façade's first episodes env by env."""
    import gym_td_b200 as G
    from gym_td_b200 import balance as B
    from oracle import ref_loader
    np.seterr(all="ignore")
    t, num_enemy = 1, 6                                         # min(100 // 15, 8)
    if ref_loader.available():
        from oracle import ref_harness as RH
        ref_loader.load()
        ref = RH.make_env("atk", 10, 1024)
        random.seed(1024)
        w1, r1 = _round_road_episode(ref, t, num_enemy)
        mine = G.make("TD-atk-small-v0", seed=1024)
        random.seed(1024)
        w2, r2 = _round_road_episode(mine, t, num_enemy)
        mine.close()
        assert repr(w1) == repr(w2) and repr(float(r1)) == repr(float(r2))
    K, base = 6, 4242
    vec = G.make_vec("TD-atk-small-v0", K, seed=base)
    agent = B.RoundRoadAttacker(vec, t)
    assert agent.num_enemy == num_enemy
    wins, rets = B.evaluate(vec, agent, 1)
    wins, rets = wins.cpu().numpy(), rets.cpu().numpy()
    for i in range(K):
        env = G.make("TD-atk-small-v0", seed=int(vec.map_seeds[i]))
        random.seed(base + i)
        w, r = _round_road_episode(env, t, num_enemy)
        env.close()
        assert int(wins[i, 0]) == int(bool(w)) and abs(rets[i, 0] - r) < 1e-9, (i, wins[i, 0], w, rets[i, 0], r)
    vec.close()
