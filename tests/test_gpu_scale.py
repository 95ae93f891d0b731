"""Parity at the BASELINE.json batch sizes: the batch runs at its full per-GPU size
with auto-reset on, and the first 256 envs are replayed step by step through the CPU oracle -- every output
(reward bits, done, win, AllowNextMove, RealAction, FailCode) and the whole float32 observation, bit for bit.
Config 2 runs a full episode and beyond (>= 1,250 steps, through the step-limit auto-reset of every env).
Also: results do not depend on how the global env range is cut into ranks (the env_offset sharding rule)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SUB = 256


def _actions(torch, kind, L, N, multi, K, seed, sparse=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a_atk = torch.randint(0, 5, (8, N, 3, 8), dtype=torch.int64, device="cuda", generator=g) if kind != "def" else None
    if kind == "atk":
        a_def = None
    elif multi:
        if sparse:
            a_def = (torch.rand((4, N, 6, L, L), device="cuda", generator=g) < 0.01).to(torch.int64)
        else:
            a_def = torch.randint(0, 3, (4, N, 6, L, L), dtype=torch.int64, device="cuda", generator=g)
    else:
        a_def = torch.randint(0, 6 * L * L + 1, (16, N), dtype=torch.int64, device="cuda", generator=g)

    def action(k):
        d = a_def[k % a_def.shape[0]] if a_def is not None else None
        a = a_atk[k % a_atk.shape[0]] if a_atk is not None else None
        return d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}
    return action


def _replay(kind, L, N, multi, steps, seed, sparse=False, sub=SUB):
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    from oracle.replay import Replayer
    env = TDVecEnv(kind, L, N, seed=seed, auto_reset=True, multi_action=multi)
    env.reset()
    torch.cuda.synchronize()
    action = _actions(torch, kind, L, N, multi, steps, 100 + seed, sparse)
    R = Replayer(env, sub)
    R.check_initial_obs()
    episodes0 = 0
    for k in range(steps):
        a = action(k)
        env.step(a)
        torch.cuda.synchronize()
        R.check_step(a)
        assert not R.mismatches, R.mismatches[:3]
    st = env.stats()
    env.close()
    s = R.summary()
    assert s["mismatches"] == 0 and s["compared_env_steps"] == sub * steps
    return s, st, R


def test_config2_full_episode_replay_at_65536_envs():
    """TD-def-small-v0, 65,536 envs, Discrete actions: 1,260 steps = every env passes the 1,200-step limit (or its
    base falls) and restarts on the next map of the pool at least once, replayed env by env."""
    s, st, R = _replay("def", 10, 65536, False, 1260, seed=0)
    assert st["episodes"] >= 65536                                # every env finished at least one episode
    assert all(m != i for i, m in enumerate(R.map_id))            # ... including each replayed one


@pytest.mark.parametrize("kind,L,N,multi,steps,sparse", [
    ("def", 20, 32768, True, 160, False),      # config 3: Box(6,20,20) uniform {0,1,2} (action_space.sample())
    ("def", 20, 32768, True, 160, True),       # config 3, sparse variant (flags 1 with p = 0.01)
    ("atk", 10, 65536, False, 420, False),     # config 4: cluster actions, scripted defender lv1 on device
    ("2p", 30, 16384, False, 200, False),      # config 5: Dict actions
])
def test_configs_3_4_5_replayed_subset_at_full_size(kind, L, N, multi, steps, sparse):
    s, st, R = _replay(kind, L, N, multi, steps, seed=3, sparse=sparse)
    assert s["replayed_envs"] == SUB


def test_results_do_not_depend_on_the_rank_split():
    """Two handles stepping global env ranges [0, n) and [n, 2n) (what ranks 0 and 1 of a torchrun job do with
    env_offset; on a multi-GPU box the second handle lives on cuda:1) against one handle stepping [0, 2n):
    identical per-env outputs every step, across auto-resets (the map pools overlap by `extra` maps so that an
    env's e-th episode uses the same global map in both layouts)."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    n, L, K, extra = 1024, 10, 1300, 16
    dev1 = 1 if torch.cuda.device_count() > 1 else 0
    whole = TDVecEnv("def", L, 2 * n, seed=5, auto_reset=True, n_maps=2 * n + extra)
    parts = [TDVecEnv("def", L, n, seed=5, auto_reset=True, env_offset=0, n_maps=n + extra, device=0),
             TDVecEnv("def", L, n, seed=5, auto_reset=True, env_offset=n, n_maps=n + extra, device=dev1)]
    whole.reset()
    for p in parts:
        p.reset()
    g = torch.Generator(device="cuda:0").manual_seed(9)
    max_ep = 0
    for k in range(K):
        a = torch.randint(0, 601, (2 * n,), dtype=torch.int64, device="cuda:0", generator=g)
        o, r, d, info = whole.step(a)
        for j, p in enumerate(parts):
            aj = a[j * n:(j + 1) * n].to(p.device).contiguous()
            o2, r2, d2, info2 = p.step(aj)
            sl = slice(j * n, (j + 1) * n)
            assert torch.equal(r[sl].view(torch.int64).cpu(), r2.view(torch.int64).cpu()), (k, j)
            assert torch.equal(d[sl].cpu(), d2.cpu()) and torch.equal(info["RealAction"][sl].cpu(), info2["RealAction"].cpu())
            if k % 50 == 0 or k > K - 5:
                assert torch.equal(o[sl].view(torch.int32).cpu(), o2.view(torch.int32).cpu()), (k, j)
    hdr = whole.engine.decode_state(whole.engine.get_state_raw(0, 1)[0])["header"]
    assert whole.stats()["episodes"] >= 2 * n
    eps = [whole.engine.decode_state(b)["header"]["episode"] for b in whole.engine.get_state_raw(0, 2 * n)]
    assert max(eps) <= extra, "the overlap of the map pools must cover every episode played"
    whole.close()
    for p in parts:
        p.close()


def test_two_handles_with_different_configs_coexist():
    """the game config is per handle (it travels in the kernel parameters).  Two live handles on one
    device with different configs, stepped alternately, each bit-identical to the oracle under ITS config."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    from oracle.replay import Replayer
    from tests import parity_util as PU
    cfg_a = PU.make_config()
    cfg_b = PU.make_config(base_LP=9, max_cost=60, defender_init_cost=40, reward_kill=0.5, max_episode_steps=90,
                           tower_distance=1, enemy_speed=[[.5, .5], [.25, .25], [.2, .2], [.2, .2]])
    n, L = 64, 10
    ea = TDVecEnv("def", L, n, seed=21, auto_reset=True, cfg=cfg_a)
    eb = TDVecEnv("def", L, n, seed=21, auto_reset=True, cfg=cfg_b)
    ea.reset(), eb.reset()
    torch.cuda.synchronize()
    ra, rb = Replayer(ea, n, cfg=cfg_a), Replayer(eb, n, cfg=cfg_b)
    g = torch.Generator(device="cuda").manual_seed(4)
    differ = False
    for k in range(260):
        a = torch.randint(0, 601, (n,), dtype=torch.int64, device="cuda", generator=g)
        ea.step(a)
        eb.step(a)
        torch.cuda.synchronize()
        ra.check_step(a), rb.check_step(a)
        assert not ra.mismatches and not rb.mismatches, (ra.mismatches[:2], rb.mismatches[:2])
        differ = differ or not torch.equal(ea.reward, eb.reward)
    assert differ
    ea.close(), eb.close()


@pytest.mark.parametrize("kind,multi", [("def", False), ("def", True), ("atk", False)])
def test_opponent_specialised_kernels_match_the_generic_ones(kind, multi):
    """The step kernels compiled for the default scripted opponent (level 1 on the device generator, variants 5-7 of
    td_engine.cu) and the generic ones (TD_OPT_GENERIC_KERNELS) are the same step: outputs, observation and the whole
    env record, every step, across auto-resets; full writes and in-place observation updates."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    n, L = 96, 10
    kw = dict(seed=77, auto_reset=True, multi_action=multi) if kind == "def" else dict(seed=77, auto_reset=True)
    for inc in (False, True):
        a_env = TDVecEnv(kind, L, n, incremental_obs=inc, **kw)
        b_env = TDVecEnv(kind, L, n, incremental_obs=inc, **kw)
        b_env.engine.set_option("generic_kernels", 1)
        a_env.reset(), b_env.reset()
        g = torch.Generator(device="cuda").manual_seed(5)
        for k in range(150 if multi else 400):
            if kind == "atk":
                act = torch.randint(0, 5, (n, 3, 8), dtype=torch.int64, device="cuda", generator=g)
            elif multi:
                act = (torch.rand((n, 6, L, L), device="cuda", generator=g) < 0.02).to(torch.int64)
            else:
                act = torch.randint(0, 6 * L * L + 1, (n,), dtype=torch.int64, device="cuda", generator=g)
            _, ra, da, _ = a_env.step(act)
            _, rb, db, _ = b_env.step(act)
            assert torch.equal(a_env.obs, b_env.obs), (kind, multi, inc, k)
            assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(a_env._slab, b_env._slab), (kind, multi, inc, k)
        torch.cuda.synchronize()
        assert (a_env.engine.get_state_raw() == b_env.engine.get_state_raw()).all()
        a_env.close(), b_env.close()


@pytest.mark.parametrize("kind,L,multi,inc", [("def", 10, False, False), ("def", 10, False, True), ("def", 20, True, False),
                                              ("2p", 30, False, True)])
def test_compressible_observation_memory_is_transparent(kind, L, multi, inc):
    """TDVecEnv keeps the observation (and a large RealAction slab) in a compressible allocation
    (td_alloc_compressible).  Compression is lossless and happens in L2: every output, the observation as tensors,
    clones and host copies see it, and the env record must equal those of an env writing into ordinary torch memory."""
    import torch
    from gym_td_b200.vec_env import TDVecEnv
    n = {10: 2048, 20: 512, 30: 256}[L]
    kw = dict(seed=11, auto_reset=True, incremental_obs=inc)
    if kind == "def":
        kw["multi_action"] = multi
    a = TDVecEnv(kind, L, n, obs_memory="plain", **kw)
    b = TDVecEnv(kind, L, n, obs_memory="compressible", **kw)
    assert a.obs_memory == "plain" and b.obs_memory in ("compressible", "plain (compression not granted)")
    if b.obs_memory != "compressible":
        pytest.skip("the driver did not grant a compressible allocation")
    assert b.slab_memory == "compressible"
    a.reset(), b.reset()
    assert torch.equal(a.obs, b.obs)
    g = torch.Generator(device="cuda").manual_seed(6)
    for k in range(120):
        d = ((torch.rand((n, 6, L, L), device="cuda", generator=g) < 0.02).to(torch.int64) if multi
             else torch.randint(0, 6 * L * L + 1, (n,), dtype=torch.int64, device="cuda", generator=g))
        atk = torch.randint(0, 5, (n, 3, 8), dtype=torch.int64, device="cuda", generator=g)
        act = d if kind == "def" else atk if kind == "atk" else {"Attacker": atk, "Defender": d}
        if k % 9 == 4:        # the host-buffer call, with the observation copied out of the compressible buffer
            hact = d.cpu().pin_memory() if kind == "def" else {"Attacker": atk.cpu().pin_memory(), "Defender": d.cpu().pin_memory()}
            a.step(act)
            h = b.step_host(hact, want_obs=True)
            assert torch.equal(h["obs"].view(torch.int32), a.obs.cpu().view(torch.int32)), (k,)
        else:
            a.step(act)
            b.step(act)
        assert torch.equal(a.obs, b.obs) and torch.equal(a._slab, b._slab), (kind, L, multi, inc, k)
        if k % 40 == 0:
            assert torch.equal(b.obs.clone(), a.obs) and torch.equal(b.obs[n // 2:].cpu(), a.obs[n // 2:].cpu())
    torch.cuda.synchronize()
    assert (a.engine.get_state_raw() == b.engine.get_state_raw()).all()
    a.close(), b.close()
    del b
    torch.cuda.empty_cache()


def test_compressible_allocator_abi():
    """td_alloc_compressible / td_free_compressible: zero-filled device memory, errors as return codes."""
    import ctypes as C
    import torch
    from gym_td_b200 import engine as E
    lib = E.lib()
    ptr, granted = C.c_void_p(), C.c_int(-1)
    assert lib.td_alloc_compressible(0, 0, C.byref(ptr), C.byref(granted)) == -1          # TD_E_INVALID
    rc = lib.td_alloc_compressible(0, 5 << 20, C.byref(ptr), C.byref(granted))
    if rc == -4:
        pytest.skip("no compressible memory on this device: " + lib.td_last_error(None).decode())
    assert rc == 0 and ptr.value and granted.value in (0, 1)
    buf = E.CompressibleBuffer(3 << 20, 0)
    t = buf.tensor((3 << 18,), torch.float32)
    assert t.is_cuda and float(t.abs().sum()) == 0.0
    t.fill_(2.0)
    assert float(t.sum()) == 2.0 * (3 << 18)
    del t, buf
    assert lib.td_free_compressible(ptr) == 0
    assert lib.td_free_compressible(ptr) == -1                                               # not ours any more
    assert lib.td_free_compressible(None) == 0
