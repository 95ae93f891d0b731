"""Replay a subset of a TDVecEnv batch through the CPU oracle, step by step (TEST INFRASTRUCTURE ONLY).

Used by tests/ (parity at the BASELINE batch sizes) and by bench.py's post-timing `replay` check (every rank
replays its first K envs, SURVEY.md 8(d) config 2: "first 256 envs of each GPU stepped through the reference").
Never imported by gym_td_b200.

The oracle side follows the reference wrappers (TDDefense.step / TDAttack.step / TDMulti.step restated in
oracle/td_oracle.c) and the auto-reset contract of TDVecEnv: a finished env restarts on map
(map_id + 1) % n_maps of its handle's pool, the scripted opponent's generator keeps running across episodes
(like the reference's global `random` module).
"""
import numpy as np

from gym_td_b200 import engine as E
from gym_td_b200 import mapgen

from . import td_oracle as TO


def oracle_config(cfg=None):
    """gym_td_b200 TdConfig (or None = defaults) -> oracle Config (same field names)."""
    if cfg is None:
        return TO.default_config()
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


def oracle_env_from_map(m, ocfg):
    p = mapgen.planes(m)
    road_bits = (p["road"][0] | (p["road"][1] << 1) | (p["road"][2] << 2) | (p["road"][3] << 3)).astype(np.uint8)
    o = TO.OracleEnv(ocfg)
    o.init_from_planes(p["map_size"], p["num_roads"], p["start"], p["end"], road_bits, p["dist"], p["dir"])
    return o


class Replayer(object):
    """Oracle twins of envs [0, n_sub) of `env` (a TDVecEnv that was just reset())."""

    def __init__(self, env, n_sub, difficulty=1, cfg=None):
        self.env, self.kind, self.L = env, env.kind, env.map_size
        self.n_sub = min(int(n_sub), env.num_envs)
        self.multi = bool(env.multi_action)
        self.n_maps = env.engine.n_maps
        self.ocfg = oracle_config(cfg)
        self.difficulty = difficulty
        self.map_cache = {}
        self.map_id = list(range(self.n_sub))                      # td_reset: env i starts on map i % n_maps
        self.scripted = self.kind != "2p" and env.scripted
        self.twins = [oracle_env_from_map(self._map(i % self.n_maps), self.ocfg) for i in range(self.n_sub)]
        if self.scripted:
            states = env.engine.get_opponent(0, self.n_sub)        # continue the device generators where they are
            for o, st in zip(self.twins, states):
                o.set_pyrand((3, tuple(int(x) for x in st), None))
        self.steps = 0
        self.compared = 0
        self.mismatches = []

    def _map(self, j):
        if j not in self.map_cache:
            m = mapgen.generate(int(self.env.map_seeds[j]), self.L)
            assert m is not None
            self.map_cache[j] = m
        return self.map_cache[j]

    def check_initial_obs(self):
        obs = self.env.obs[:self.n_sub].cpu().numpy()
        for i, o in enumerate(self.twins):
            if not np.array_equal(obs[i].view(np.uint32), o.get_states().view(np.uint32)):
                self.mismatches.append("env %d: observation after reset" % i)

    def check_step(self, action):
        """Call after env.step(action): steps the twins with the same action and compares every output of
        the first n_sub envs bit for bit.  Returns the number of new mismatches."""
        env, K, kind = self.env, self.n_sub, self.kind
        if kind == "def":
            a_def, a_atk = action, None
        elif kind == "atk":
            a_def, a_atk = None, action
        else:
            a_def, a_atk = action["Defender"], action["Attacker"]
        a_def = a_def[:K].cpu().numpy() if a_def is not None else None
        a_atk = a_atk[:K].cpu().numpy() if a_atk is not None else None
        obs = env.obs[:K].cpu().numpy()
        rew = env.reward[:K].cpu().numpy()
        done = env._done[:K].cpu().numpy()
        win = env.win[:K].cpu().numpy()
        allow = env._allow[:K].cpu().numpy()
        real_def = env.real_def[:K].cpu().numpy() if env.real_def is not None else None
        fail_def = env.fail_def[:K].cpu().numpy() if env.fail_def is not None else None
        real_atk = env.real_atk[:K].cpu().numpy() if env.real_atk is not None else None
        fail_atk = env.fail_atk[:K].cpu().numpy() if env.fail_atk is not None else None
        before = len(self.mismatches)
        self.steps += 1
        diff = self.difficulty if self.scripted else -1
        for i in range(K):
            o = self.twins[i]
            real_multi = None
            if kind == "def":
                if self.multi:
                    out, real_multi = o.def_step_multi(a_def[i], diff)
                else:
                    out = o.def_step(int(a_def[i]), diff)
            elif kind == "atk":
                out = o.atk_step(a_atk[i], diff)
            elif self.multi:
                out, real_multi = o.multi_step_multi(a_atk[i], a_def[i])
            else:
                out = o.multi_step(a_atk[i], int(a_def[i]))
            bad = []
            if float(rew[i]).hex() != float(out.reward).hex():
                bad.append("reward %r != %r" % (float(rew[i]), out.reward))
            if bool(done[i]) != bool(out.done):
                bad.append("done")
            if int(win[i]) != int(out.win):
                bad.append("win %d != %d" % (win[i], out.win))
            if int(allow[i]) != ((1 if out.allow_next_def else 0) | (2 if out.allow_next_atk else 0)):
                bad.append("allow_next")
            if kind != "atk":
                if self.multi:
                    if not np.array_equal(real_def[i], real_multi):
                        bad.append("real_def (multi)")
                else:
                    if int(real_def[i]) != int(out.real_def):
                        bad.append("real_def")
                    if int(fail_def[i]) != int(out.fail_def):
                        bad.append("fail_def")
            if kind != "def":
                if not np.array_equal(real_atk[i], np.ctypeslib.as_array(out.real_atk)):
                    bad.append("real_atk")
                want = [out.n_fail_atk] + list(out.fail_atk[:out.n_fail_atk])
                if fail_atk[i].tolist()[:1 + out.n_fail_atk] != want:
                    bad.append("fail_atk")
            finished = bool(out.done)
            if finished and env.auto_reset:
                # auto-reset inside the step: next map of the pool, the opponent's generator keeps running
                py = o.e.pyrand
                self.map_id[i] = (self.map_id[i] + 1) % self.n_maps
                o = self.twins[i] = oracle_env_from_map(self._map(self.map_id[i]), self.ocfg)
                o.e.pyrand = py
            if not np.array_equal(obs[i].view(np.uint32), o.get_states().view(np.uint32)):
                bad.append("observation")
            if bad:
                self.mismatches.append("env %d step %d: %s" % (i, self.steps, ", ".join(bad)))
            self.compared += 1
        return len(self.mismatches) - before

    def summary(self):
        return {"replayed_envs": self.n_sub, "replayed_steps": self.steps, "compared_env_steps": self.compared,
                "mismatches": len(self.mismatches), "first_mismatches": self.mismatches[:3]}
