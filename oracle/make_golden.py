"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference or baseline/_ref):
    python -m oracle.make_golden
The fixtures travel to the GPU box (the reference source does not) and pin both the CPU oracle and
the CUDA engine there.  Everything is produced by executing the reference's own code through the gym
stub; only the multi-action wrapper (which raises in the reference) is the restated
harness of oracle/ref_harness.py.
"""
import hashlib
import json
import os
import random
import struct
import sys

import numpy as np

from . import ref_harness as RH
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest64(b):
    return np.frombuffer(hashlib.sha256(b).digest()[:8], dtype="<u8")[0]


def state_bytes(st):
    """Canonical packing of a state dict (ref_harness.board_state / OracleEnv.state_dict)."""
    out = [struct.pack("<ddiiiiii", st["cost_def"], st["cost_atk"], -1 if st["base_LP"] is None else st["base_LP"],
                       st["steps"], st["attacker_cd"], st["defender_cd"], len(st["towers"]), len(st["enemies"]))]
    for t in st["towers"]:
        out.append(struct.pack("<iiid", int(t[0]), int(t[1]), int(t[2]), float(t[3])))
    for e in st["enemies"]:
        out.append(struct.pack("<iiddi", int(e[0]), int(e[1]), float(e[2]), float(e[4]), int(e[6])))
    out.append(np.asarray(st["map6"], dtype="<i4").tobytes())
    return b"".join(out)


def coo(a):
    idx = np.argwhere(a != 0)
    return np.concatenate([idx, a[a != 0].reshape(-1, 1)], axis=1).astype(np.int32)


def record(kind, L, seed, steps, difficulty=1, random_agent=True, multi=False, overrides=None, tag=None):
    overrides = overrides or {}
    with RH.ref_config_override(**overrides):
        if multi:
            RH.set_multiple_actions(True)
        try:
            seed, env = RH.first_valid_seed(kind, L, seed, **({} if kind == "2p" else
                                                              dict(difficulty=difficulty, random_agent=random_agent)))
            random.seed(seed)
            py_state = np.asarray(random.getstate()[1], dtype=np.uint64).astype(np.uint32)
            st = env.np_random.get_state()
            np_state = np.concatenate([np.asarray(st[1], dtype=np.uint32), [np.uint32(st[2])]])
            m = RH.board_roads(env._board)
            rs = np.random.RandomState(seed + 77)
            obs0 = env._board.get_states()
            rec = dict(reward=[], done=[], win=[], allow=[], real_def=[], fail_def=[], real_atk=[], fail_atk=[],
                       real_is_def_only=[], obs_digest=[digest64(obs0.tobytes())], state_digest=[],
                       def_action=[], atk_action=[], def_coo=[], real_coo=[])
            samples = {0: obs0.copy()}
            for t in range(1, steps + 1):
                atk = np.full((3, 8), 4, dtype=np.int64)
                if kind != "def":
                    mode = rs.randint(4)
                    if mode == 0:
                        atk = rs.randint(0, 5, size=(3, 8)).astype(np.int64)
                    elif mode >= 2:
                        atk[rs.randint(3), :rs.randint(1, 9)] = rs.randint(4)
                d = None
                if kind != "atk":
                    d = RH.sparse_multi_action(env._board, rs, 0.02) if multi else RH.smart_defender_action(env._board, rs)
                if kind == "def":
                    obs, rew, done, info = RH.def_step_multi(env, d) if multi else env.step(d)
                elif kind == "atk":
                    obs, rew, done, info = env.step(atk)
                else:
                    a = {"Attacker": atk, "Defender": d}
                    obs, rew, done, info = RH.multi_step_multi(env, a) if multi else env.step(a)
                rec["reward"].append(float(rew))
                rec["done"].append(bool(done))
                w = info["Win"]
                if isinstance(w, dict):
                    w = w["Defender"]
                rec["win"].append(-1 if w is None else int(bool(w)))
                al = info["AllowNextMove"]
                if isinstance(al, dict):
                    rec["allow"].append((1 if al["Defender"] else 0) | (2 if al["Attacker"] else 0))
                else:
                    # single-agent envs report only their own flag; the other bit comes from the env object
                    rec["allow"].append((1 if env.defender_cd <= 1 else 0) | (2 if env.attacker_cd <= 1 else 0))
                ra, fc = info["RealAction"], info["FailCode"]
                only = 0
                r_def, r_atk, f_def, f_atk = 6 * L * L, atk, 0, []
                if kind == "def":
                    r_def, f_def = ra, fc
                elif kind == "atk":
                    r_atk, f_atk = ra, fc
                else:
                    if isinstance(ra, dict):
                        r_def, r_atk = ra["Defender"], ra["Attacker"]
                    else:
                        r_def, only = ra, 1
                    f_def, f_atk = fc["Defender"], fc["Attacker"]
                if multi and kind != "atk":
                    rec["real_coo"].append(np.concatenate([np.full((len(coo(r_def)), 1), t), coo(r_def)], axis=1))
                    rec["def_coo"].append(np.concatenate([np.full((len(coo(d)), 1), t), coo(d)], axis=1))
                    r_def = 0
                rec["real_def"].append(int(r_def))
                rec["fail_def"].append(int(f_def))
                rec["real_atk"].append(np.asarray(r_atk, dtype=np.int64))
                rec["fail_atk"].append([len(f_atk)] + list(f_atk) + [0] * (3 - len(f_atk)))
                rec["real_is_def_only"].append(only)
                rec["def_action"].append(0 if (d is None or multi) else int(d))
                rec["atk_action"].append(atk)
                rec["obs_digest"].append(digest64(obs.tobytes()))
                rec["state_digest"].append(digest64(state_bytes(RH.board_state(env))))
                if t in (1, steps // 2) or done or t == steps:
                    samples[t] = obs.copy()
                if done:
                    break
        finally:
            if multi:
                RH.set_multiple_actions(False)
    T = len(rec["reward"])
    name = tag or "%s_L%d_s%d%s%s" % (kind, L, seed, "_multi" if multi else "",
                                      "" if kind == "2p" else "_d%d%s" % (difficulty, "" if random_agent else "_np"))
    meta = dict(kind=kind, L=L, seed=int(seed), difficulty=difficulty, random_agent=bool(random_agent),
                multi=bool(multi), overrides=overrides, steps=T)
    empty = np.zeros((0, 5), dtype=np.int32)
    np.savez_compressed(
        os.path.join(OUT, "traj_%s.npz" % name), meta=json.dumps(meta),
        num_roads=m["num_roads"], start=np.asarray(m["start"], dtype=np.int32), end=m["end"],
        road=m["road"], dist=m["dist"].astype(np.uint8), dir=m["dir"].astype(np.uint8),
        py_state=py_state, np_state=np_state,
        def_action=np.asarray(rec["def_action"], dtype=np.int64), atk_action=np.asarray(rec["atk_action"], dtype=np.int8),
        def_coo=np.concatenate(rec["def_coo"]).astype(np.int32) if rec["def_coo"] else empty,
        real_coo=np.concatenate(rec["real_coo"]).astype(np.int32) if rec["real_coo"] else empty,
        reward=np.asarray(rec["reward"], dtype=np.float64), done=np.asarray(rec["done"], dtype=np.uint8),
        win=np.asarray(rec["win"], dtype=np.int8), allow=np.asarray(rec["allow"], dtype=np.uint8),
        real_def=np.asarray(rec["real_def"], dtype=np.int64), fail_def=np.asarray(rec["fail_def"], dtype=np.int32),
        real_atk=np.asarray(rec["real_atk"], dtype=np.int8), fail_atk=np.asarray(rec["fail_atk"], dtype=np.int32),
        real_is_def_only=np.asarray(rec["real_is_def_only"], dtype=np.uint8),
        obs_digest=np.asarray(rec["obs_digest"], dtype=np.uint64),
        state_digest=np.asarray(rec["state_digest"], dtype=np.uint64),
        sample_steps=np.asarray(sorted(samples), dtype=np.int32),
        sample_obs=np.stack([samples[k] for k in sorted(samples)]))
    print("traj_%s: %d steps, return %r" % (name, T, sum(rec["reward"])), flush=True)


def golden_maps(n_per_size=2000):
    """Map-generator fixtures: validity, num_roads, randint count and a digest for many seeds."""
    out = {}
    for L in (10, 20, 30):
        n = n_per_size if L == 10 else n_per_size // 4
        valid = np.zeros(n, dtype=np.uint8)
        nroads = np.zeros(n, dtype=np.uint8)
        dig = np.zeros(n, dtype=np.uint64)
        full = []
        for seed in range(n):
            res = RH.generate_roads(L, seed)
            if res is None:
                continue
            num_roads, roads = res
            from gym_TD.envs.TDBoard import TDBoard
            from gym.utils import seeding
            rng = seeding.CountingRandomState(seed)
            nr = int(rng.randint(low=1, high=4))
            b = TDBoard(L, nr, rng, 10, 0, 100, 5)
            m = RH.board_roads(b)
            valid[seed], nroads[seed] = 1, nr
            packed = (m["road"] | (m["dir"].astype(np.uint8) << 4)).astype(np.uint8)
            blob = packed.tobytes() + m["dist"].astype(np.uint8).tobytes() + \
                np.asarray(m["start"] + [0] * (3 - len(m["start"])) + [m["end"]], dtype="<i4").tobytes()
            dig[seed] = digest64(blob)
            if len(full) < 8:
                full.append((seed, packed.reshape(L, L), m["dist"].astype(np.uint8).reshape(L, L)))
        out["valid_%d" % L], out["num_roads_%d" % L], out["digest_%d" % L] = valid, nroads, dig
        out["full_seeds_%d" % L] = np.asarray([f[0] for f in full], dtype=np.int32)
        out["full_cells_%d" % L] = np.stack([f[1] for f in full])
        out["full_dist_%d" % L] = np.stack([f[2] for f in full])
        print("maps L=%d: %d seeds, %d invalid" % (L, n, int(n - valid.sum())), flush=True)
    np.savez_compressed(os.path.join(OUT, "maps.npz"), **out)


def reference_selftest():
    """The reference's own golden vector (TDBoard.py:674-756): RandomState(1024), 2 roads, 10x10."""
    from gym_TD.envs.TDBoard import TDBoard
    from gym_TD.envs.TDParam import config
    rng = np.random.RandomState()
    rng.seed(1024)
    b = TDBoard(10, 2, rng, config.defender_init_cost, config.attacker_init_cost, config.max_cost, config.base_LP)
    m = RH.board_roads(b)
    np.savez_compressed(os.path.join(OUT, "ref_selftest.npz"), obs0=b.get_states(), road=m["road"],
                        dist=m["dist"].astype(np.uint8), dir=m["dir"].astype(np.uint8),
                        start=np.asarray(m["start"], dtype=np.int32), end=m["end"])


def survey_kats():
    """The trajectories quoted in SURVEY.md 8(c): digest over obs0 || (obs_t || f64 reward_t)..."""
    from gym_TD.envs import TDDefense, TDAttack, TDMulti
    kats = []
    for L in (10, 20, 30):
        for kind in ("def", "2p", "atk"):
            rs = np.random.RandomState(0)
            if kind == "def":
                env = TDDefense(L, seed=1024, random_agent=False)
            elif kind == "2p":
                env = TDMulti(L, seed=1024, random_agent=False)
            else:
                random.seed(1024)
                env = TDAttack(L, seed=1024, random_agent=True)
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
                o, r, d, _ = env.step(a)
                h.update(o.tobytes())
                h.update(np.float64(r).tobytes())
                ret += r
                n += 1
                if d:
                    break
            kats.append(dict(kind=kind, L=L, steps=n, ret=repr(ret), sha=h.hexdigest()[:16]))
            print(kats[-1], flush=True)
    json.dump(kats, open(os.path.join(OUT, "survey_kats.json"), "w"), indent=1)


def extra():
    """Second batch of recorded trajectories (python oracle/make_golden.py --extra): the first batch stays as it is."""
    record("atk", 10, 510, 600, difficulty=2)
    record("atk", 30, 550, 300, difficulty=2)
    record("def", 20, 540, 400, difficulty=0)
    record("def", 30, 520, 200, multi=True)
    record("2p", 30, 530, 200, multi=True)
    record("atk", 20, 560, 600, overrides=dict(enemy_upgrade_at=0.05, max_cost=150, reward_kill=0.25),
           tag="atk_L20_override_c")


def extra2():
    """Third batch (python -m oracle.make_golden --extra2): the scripted opponents on the env's own np_random
    (random_agent=False) that the reference can run -- random_enemy_lv0, random_tower_lv0 -- and the multi-action
    defender against the np_random attacker."""
    record("def", 10, 610, 600, difficulty=0, random_agent=False)
    record("atk", 10, 620, 600, difficulty=0, random_agent=False)
    record("def", 20, 630, 300, multi=True, random_agent=False)


def main():
    ref_loader.load()
    np.seterr(all="ignore")
    os.makedirs(OUT, exist_ok=True)
    if "--extra" in sys.argv:
        extra()
        return 0
    if "--extra2" in sys.argv:
        extra2()
        return 0
    reference_selftest()
    survey_kats()
    golden_maps()
    for L, steps in ((10, 1200), (20, 600), (30, 400)):
        record("def", L, 100 + L, steps)
        record("atk", L, 200 + L, steps)
        record("2p", L, 300 + L, steps)
    record("def", 10, 410, 600, difficulty=0)
    record("def", 10, 420, 600, random_agent=False)
    record("atk", 10, 430, 600, difficulty=0)
    record("atk", 20, 440, 400, difficulty=2)
    record("def", 20, 450, 400, multi=True)
    record("2p", 10, 460, 400, multi=True)
    record("def", 10, 470, 1200, overrides=dict(base_LP=None, defender_action_interval=3, attacker_action_interval=2),
           tag="def_L10_override_a")
    record("atk", 10, 480, 1200, overrides=dict(base_LP=None, defender_action_interval=3, attacker_action_interval=2),
           tag="atk_L10_override_a")
    record("2p", 20, 490, 600, overrides=dict(defender_init_cost=60, attacker_init_cost=50, defender_cost_rate=.7),
           tag="2p_L20_override_b")
    return 0


if __name__ == "__main__":
    sys.exit(main())
