"""Shared machinery of the GPU parity tests: drive the CUDA engine (through the C ABI) and the CPU
oracle (oracle/td_oracle.c) with the same maps, actions and opponent randomness, and compare every
output and the whole board state after every step.

Comparison bar: integers, f64 costs / LP / margins / rewards and the float32 observation are all
compared BIT FOR BIT (stricter than the 1e-6 relative tolerance north_star allows for the float
LP ratios and averages).
"""
import ctypes as C
import random

import numpy as np

from gym_td_b200 import engine as E
from gym_td_b200 import mapgen
from oracle import td_oracle as TO


LAST_MAX = {"towers": 0, "enemies": 0}      # longest lists seen by the last run_parity (coverage evidence)


def oracle_config_from_engine(cfg):
    """TdConfig (product) -> oracle Config; the two structs have the same field names."""
    o = TO.Config()
    for name, _ in E.TdConfig._fields_:
        src = getattr(cfg, name)
        if hasattr(src, "__len__"):
            for t in range(4):
                for l in range(2):
                    getattr(o, name)[t][l] = src[t][l]
        else:
            setattr(o, name, src)
    return o


def make_config(**overrides):
    c = E.TdConfig()
    E.lib().td_default_config(C.byref(c))
    for k, v in overrides.items():
        if k == "base_LP" and v is None:
            v = -1
        cur = getattr(c, k)
        if hasattr(cur, "__len__"):
            for t in range(4):
                for l in range(2):
                    cur[t][l] = v[t][l]
        else:
            setattr(c, k, v)
    return c


def oracle_env_from_map(m, ocfg):
    p = mapgen.planes(m)
    road_bits = (p["road"][0] | (p["road"][1] << 1) | (p["road"][2] << 2) | (p["road"][3] << 3)).astype(np.uint8)
    o = TO.OracleEnv(ocfg)
    o.init_from_planes(p["map_size"], p["num_roads"], p["start"], p["end"], road_bits, p["dist"], p["dir"])
    return o


def smart_defender_action(o, rs, p_nop=0.3, p_uniform=0.2):
    """Action generator that really exercises build / LvUp / destruct (uniform actions mostly fail)."""
    e = o.e
    L = e.L
    cells = L * L
    nop = 6 * cells
    u = rs.random_sample()
    if u < p_nop:
        return nop
    if u < p_nop + p_uniform:
        return int(rs.randint(nop + 1))
    kind = rs.randint(10)
    if kind < 6 or e.n_towers == 0:
        map6 = np.ctypeslib.as_array(e.map6)[:cells]
        free = np.flatnonzero(map6 == 0)
        if len(free) == 0:
            return nop
        road = (np.ctypeslib.as_array(e.road)[:cells] & 1).reshape(L, L)
        loc = int(free[rs.randint(len(free))])
        for _ in range(6):
            cand = int(free[rs.randint(len(free))])
            r, c = divmod(cand, L)
            if road[max(r - 2, 0):r + 3, max(c - 2, 0):c + 3].any():
                loc = cand
                break
        return int(rs.randint(4)) * cells + loc
    tw = e.towers[rs.randint(e.n_towers)]
    return (4 if kind < 9 else 5) * cells + tw.loc


def uniform_multi_action(L, rs):
    """Box(0, 2, (6, L, L), int64).sample(): the distribution of BASELINE config 3."""
    return rs.randint(0, 3, size=(6, L, L)).astype(np.int64)


def sparse_multi_action(L, rs, p=0.02):
    a = (rs.random_sample((6, L, L)) < p).astype(np.int64)
    a[rs.random_sample((6, L, L)) < p] = 2
    return a


def attacker_action(rs):
    mode = rs.randint(4)
    if mode == 0:
        return rs.randint(0, 5, size=(3, 8)).astype(np.int64)
    a = np.full((3, 8), 4, dtype=np.int64)
    if mode >= 2:
        a[rs.randint(3), :rs.randint(1, 9)] = rs.randint(4)
    return a


def compare_state(tag, eng, rec, o):
    """CUDA env record vs oracle env; raises AssertionError naming what differs."""
    st = eng.decode_state(rec)
    h, e = st["header"], o.e
    cells = e.L * e.L
    bad = []
    if float(h["cost_def"]).hex() != float(e.cost_def).hex():
        bad.append("cost_def %r != %r" % (float(h["cost_def"]), e.cost_def))
    if float(h["cost_atk"]).hex() != float(e.cost_atk).hex():
        bad.append("cost_atk %r != %r" % (float(h["cost_atk"]), e.cost_atk))
    for name, want in (("base_LP", e.base_LP if e.has_base_LP else -1), ("steps", e.steps),
                       ("defender_cd", e.defender_cd), ("attacker_cd", e.attacker_cd),
                       ("n_towers", e.n_towers), ("n_enemies", e.n_enemies)):
        if int(h[name]) != int(want):
            bad.append("%s %d != %d" % (name, int(h[name]), int(want)))
    if not bad:
        for i in range(e.n_towers):
            t, r = e.towers[i], st["towers"][i]
            got = (int(r["loc"]), int(r["type_lv"]) & 3, int(r["type_lv"]) >> 2, float(r["cd"]).hex())
            want = (t.loc, t.type, t.lv, float(t.cd).hex())
            if got != want:
                bad.append("tower %d %r != %r" % (i, got, want))
        for i in range(e.n_enemies):
            x, r = e.enemies[i], st["enemies"][i]
            got = (int(r["loc"]), int(r["type_lv"]) & 3, float(r["LP"]).hex(), float(r["margin"]).hex(),
                   int(r["slowdown"]))
            want = (x.loc, x.type, float(x.LP).hex(), float(x.margin).hex(), x.slowdown)
            if got != want:
                bad.append("enemy %d %r != %r" % (i, got, want))
        if not np.array_equal(st["map6"], np.ctypeslib.as_array(e.map6)[:cells]):
            bad.append("map6")
    if bad:
        raise AssertionError("%s: state mismatch: %s" % (tag, "; ".join(bad[:6])))


def run_parity(kind, L, n_envs, steps, seed=0, multi=False, opponent="device", difficulty=1, cfg_overrides=None,
               check_state_every=1, device=0, use_host_api=False, multi_mode="sparse", incremental=False):
    """Step `n_envs` instances on the GPU and in the oracle; compare everything each step.

    opponent: "device" (on-device CPython-compatible generator), "stream" (host-resolved type/road
    bytes, DEF only) or "none".  Returns the number of env-steps compared.
    """
    import torch
    dev = torch.device("cuda", device)
    LAST_MAX.update(towers=0, enemies=0)
    rs = np.random.RandomState(seed)
    cfg = make_config(**(cfg_overrides or {}))
    ocfg = oracle_config_from_engine(cfg)
    eng = E.Engine(kind, L, n_envs, device=device, cfg=cfg)
    seeds = (np.arange(n_envs, dtype=np.uint32) * 7 + 1000 * seed + 11).astype(np.uint32)
    maps, seeds, _ = mapgen.generate_batch(seeds, L)
    eng.upload_maps(maps)
    oracles = [oracle_env_from_map(maps[i], ocfg) for i in range(n_envs)]
    cells = L * L
    if opponent == "device" and kind != "2p":
        states = np.zeros((n_envs, 625), dtype=np.uint32)
        for i in range(n_envs):
            st = random.Random(int(seeds[i])).getstate()
            states[i] = np.asarray(st[1], dtype=np.uint64).astype(np.uint32)
            oracles[i].set_pyrand(st)
        eng.seed_opponent(states)
        eng.set_difficulty(difficulty)

    obs = torch.empty((n_envs, 45, L, L), dtype=torch.float32, device=dev)
    eng.reset(obs=obs)
    torch.cuda.synchronize()
    obs_h = obs.cpu().numpy()
    for i, o in enumerate(oracles):
        assert np.array_equal(obs_h[i].view(np.uint32), o.get_states().view(np.uint32)), "obs0 env %d" % i

    t = dict(
        reward=torch.zeros(n_envs, dtype=torch.float64, device=dev),
        done=torch.zeros(n_envs, dtype=torch.uint8, device=dev),
        win=torch.zeros(n_envs, dtype=torch.int8, device=dev),
        allow_next=torch.zeros(n_envs, dtype=torch.uint8, device=dev),
        fail_def=torch.zeros(n_envs, dtype=torch.int32, device=dev),
        fail_atk=torch.zeros((n_envs, 4), dtype=torch.int32, device=dev),
        real_atk=torch.zeros((n_envs, 3, 8), dtype=torch.int64, device=dev),
        real_def=torch.zeros((n_envs, 6, L, L) if multi else (n_envs,), dtype=torch.int64, device=dev),
    )
    def_shape = (n_envs, 6, L, L) if multi else (n_envs,)
    def_act = torch.zeros(def_shape, dtype=torch.int64, device=dev)
    atk_act = torch.zeros((n_envs, 3, 8), dtype=torch.int64, device=dev)
    opp = torch.zeros(n_envs, dtype=torch.uint8, device=dev)
    alive = np.ones(n_envs, dtype=bool)
    compared = 0
    outs = [None] * n_envs
    for step in range(1, steps + 1):
        a_def = np.zeros(def_shape, dtype=np.int64)
        a_atk = np.full((n_envs, 3, 8), 4, dtype=np.int64)
        a_opp = np.full(n_envs, 0xFF, dtype=np.uint8)
        for i, o in enumerate(oracles):
            if not alive[i]:
                if not multi and kind != "atk":
                    a_def[i] = 6 * cells
                continue
            if kind != "atk":
                a_def[i] = ((uniform_multi_action(L, rs) if multi_mode == "uniform" else sparse_multi_action(L, rs))
                            if multi else smart_defender_action(o, rs))
            if kind != "def":
                a_atk[i] = attacker_action(rs)
            if kind == "def" and opponent == "stream":
                tt, rd = rs.randint(4), rs.randint(o.e.num_roads)
                a_opp[i] = tt | (rd << 4)
        def_act.copy_(torch.from_numpy(a_def))
        atk_act.copy_(torch.from_numpy(a_atk))
        opp.copy_(torch.from_numpy(a_opp))
        io = E.Engine.make_io(def_action=def_act if kind != "atk" else None,
                              atk_action=atk_act if kind != "def" else None,
                              opponent=opp if (kind == "def" and opponent == "stream") else None,
                              multi_action=multi, auto_reset=False, obs=obs, reward=t["reward"], done=t["done"],
                              win=t["win"], allow_next=t["allow_next"], real_def=t["real_def"],
                              real_atk=t["real_atk"], fail_def=t["fail_def"], fail_atk=t["fail_atk"],
                              obs_incremental=incremental)
        # snapshot the records of finished envs: they must not be stepped in the comparison below
        eng.step(io, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        host = {k: v.cpu().numpy() for k, v in t.items()}
        obs_h = obs.cpu().numpy()
        recs = eng.get_state_raw() if (step % check_state_every == 0 or step == steps) else None
        for i, o in enumerate(oracles):
            if not alive[i]:
                continue
            tag = "%s L=%d env=%d step=%d" % (kind, L, i, step)
            real_multi = None
            if kind == "def":
                if multi:
                    if opponent == "stream":
                        out, real_multi = _def_multi_stream(o, a_def[i], a_opp[i])
                    else:
                        out, real_multi = o.def_step_multi(a_def[i], difficulty if opponent == "device" else -1)
                elif opponent == "stream":
                    out = _def_stream(o, int(a_def[i]), a_opp[i])
                else:
                    out = o.def_step(int(a_def[i]), difficulty if opponent == "device" else -1)
            elif kind == "atk":
                out = o.atk_step(a_atk[i], difficulty if opponent == "device" else -1)
            else:
                if multi:
                    out, real_multi = o.multi_step_multi(a_atk[i], a_def[i])
                else:
                    out = o.multi_step(a_atk[i], int(a_def[i]))
            assert float(host["reward"][i]).hex() == float(out.reward).hex(), \
                "%s reward %r != %r" % (tag, host["reward"][i], out.reward)
            assert bool(host["done"][i]) == bool(out.done), tag + " done"
            assert int(host["win"][i]) == int(out.win), tag + " win %d != %d" % (host["win"][i], out.win)
            allow = (1 if out.allow_next_def else 0) | (2 if out.allow_next_atk else 0)
            assert int(host["allow_next"][i]) == allow, tag + " allow_next"
            if kind != "atk":
                if multi:
                    assert np.array_equal(host["real_def"][i], real_multi), tag + " real_def (multi)"
                    assert int(host["fail_def"][i]) == 0
                else:
                    assert int(host["real_def"][i]) == int(out.real_def), tag + " real_def"
                    assert int(host["fail_def"][i]) == int(out.fail_def), \
                        tag + " fail_def %d != %d (a=%d)" % (host["fail_def"][i], out.fail_def, a_def[i])
            if kind != "def":
                assert np.array_equal(host["real_atk"][i], np.ctypeslib.as_array(out.real_atk)), tag + " real_atk"
                want = [out.n_fail_atk] + list(out.fail_atk[:out.n_fail_atk])
                got = host["fail_atk"][i].tolist()
                assert got[:1 + out.n_fail_atk] == want, tag + " fail_atk %r != %r" % (got, want)
            oo = o.get_states()
            if not np.array_equal(obs_h[i].view(np.uint32), oo.view(np.uint32)):
                ch = sorted(set(np.argwhere(obs_h[i] != oo)[:, 0].tolist()))
                raise AssertionError("%s obs differs in channels %s" % (tag, ch))
            if recs is not None:
                compare_state(tag, eng, recs[i], o)
            compared += 1
            LAST_MAX["towers"] = max(LAST_MAX["towers"], o.e.n_towers)
            LAST_MAX["enemies"] = max(LAST_MAX["enemies"], o.e.n_enemies)
            if out.done:
                alive[i] = False
        if not alive.any():
            break
    st = eng.stats()
    assert st["overflow_envs"] == 0, "capacity overflow flagged"
    eng.close()
    return compared


def _opp_cluster(o, byte):
    if byte == 0xFF or o.e.attacker_cd != 0:
        return
    o.summon_cluster(np.full(8, byte & 3, dtype=np.int64), (byte >> 4) & 3)
    o.e.attacker_cd = o.cfg.attacker_action_interval


def _def_stream(o, action, byte):
    """TDDefense.step with the scripted attacker's (type, road) supplied by the caller."""
    L_ = o.L_
    # decode with no opponent (difficulty -1), but the board step must come after the summon:
    # run the wrapper pieces by hand through the oracle's board API
    e = o.e
    e.attacker_cd = max(e.attacker_cd - 1, 0)
    e.defender_cd = max(e.defender_cd - 1, 0)
    cells = e.L * e.L
    out = o.out
    C.memset(C.byref(out), 0, C.sizeof(out))
    out.win, out.win_attacker = -1, -1
    out.real_def, out.fail_def = 6 * cells, 0
    if e.defender_cd == 0 and action != 6 * cells:
        act, loc = divmod(action, cells)
        ok = o.tower_build(act, loc) if act < 4 else (o.tower_lvup(loc) if act == 4 else o.tower_destruct(loc))
        if ok:
            e.defender_cd = o.cfg.defender_action_interval
            out.real_def = action
        out.fail_def = e.fail_code
    _opp_cluster(o, int(byte))
    out.reward = o.board_step()
    out.done = int(o.done())
    if out.done:
        out.win = 1 if (not e.has_base_LP or e.base_LP > 0) else 0
    out.allow_next_def = int(e.defender_cd <= 1)
    out.allow_next_atk = int(e.attacker_cd <= 1)
    return out


def _def_multi_stream(o, action, byte):
    """TDDefense.step in multi-action mode (TDDefense.py:40-60) with the scripted attacker's (type, road) supplied
    by the caller: the wrapper restated over the oracle's board calls, flagged cells in r-major / c / channel order."""
    e = o.e
    e.attacker_cd = max(e.attacker_cd - 1, 0)
    e.defender_cd = max(e.defender_cd - 1, 0)
    L = e.L
    out = o.out
    C.memset(C.byref(out), 0, C.sizeof(out))
    out.win, out.win_attacker = -1, -1
    real = np.zeros((6, L, L), dtype=np.int64)
    if e.defender_cd == 0:
        flagged = np.argwhere(np.transpose(action, (1, 2, 0)) == 1)          # sorted by (r, c, channel)
        for r, c, ch in flagged:
            loc = int(r) * L + int(c)
            ok = o.tower_build(int(ch), loc) if ch < 4 else (o.tower_lvup(loc) if ch == 4 else o.tower_destruct(loc))
            if ok:
                e.defender_cd = o.cfg.defender_action_interval
                real[ch, r, c] = 1
    _opp_cluster(o, int(byte))
    out.reward = o.board_step()
    out.done = int(o.done())
    if out.done:
        out.win = 1 if (not e.has_base_LP or e.base_LP > 0) else 0
    out.allow_next_def = int(e.defender_cd <= 1)
    out.allow_next_atk = int(e.attacker_cd <= 1)
    return out, real
