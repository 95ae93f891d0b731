"""Load the golden trajectories (generated from the unmodified reference by oracle/make_golden.py)
and replay them through the CPU oracle or the CUDA engine."""
import glob
import hashlib
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def digest64(b):
    return np.frombuffer(hashlib.sha256(b).digest()[:8], dtype="<u8")[0]


def state_bytes(st):
    out = [struct.pack("<ddiiiiii", st["cost_def"], st["cost_atk"], -1 if st["base_LP"] is None else st["base_LP"],
                       st["steps"], st["attacker_cd"], st["defender_cd"], len(st["towers"]), len(st["enemies"]))]
    for t in st["towers"]:
        out.append(struct.pack("<iiid", int(t[0]), int(t[1]), int(t[2]), float(t[3])))
    for e in st["enemies"]:
        out.append(struct.pack("<iiddi", int(e[0]), int(e[1]), float(e[2]), float(e[4]), int(e[6])))
    out.append(np.asarray(st["map6"], dtype="<i4").tobytes())
    return b"".join(out)


def trajectories():
    return sorted(glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


class Traj(object):
    def __init__(self, path):
        z = np.load(path, allow_pickle=False)
        self.name = os.path.basename(path)[5:-4]
        self.meta = json.loads(str(z["meta"]))
        self.z = z
        self.kind, self.L, self.T = self.meta["kind"], self.meta["L"], self.meta["steps"]
        self.multi = self.meta["multi"]

    def def_action(self, t):
        """Defender action of step t (1-based)."""
        if not self.multi:
            return int(self.z["def_action"][t - 1])
        return self._dense(self.z["def_coo"], t)

    def real_multi(self, t):
        return self._dense(self.z["real_coo"], t)

    def _dense(self, coo, t):
        a = np.zeros((6, self.L, self.L), dtype=np.int64)
        rows = coo[coo[:, 0] == t]
        a[rows[:, 1], rows[:, 2], rows[:, 3]] = rows[:, 4]
        return a

    def atk_action(self, t):
        return self.z["atk_action"][t - 1].astype(np.int64)

    def config_overrides(self):
        return dict(self.meta["overrides"])


def oracle_for(traj):
    """CPU oracle env initialised from a golden trajectory's map, config and generator states."""
    from oracle import td_oracle as TO
    d = {}
    for k, v in traj.config_overrides().items():
        d[k] = v
    cfg = TO.config_from_dict(d)
    o = TO.OracleEnv(cfg)
    z = traj.z
    o.init_from_planes(traj.L, int(z["num_roads"]), [int(x) for x in z["start"]], int(z["end"]), z["road"],
                       z["dist"].astype(np.int32), z["dir"].astype(np.int32))
    m = TO.MT()
    m.mt[:] = [int(x) for x in z["py_state"][:624]]
    m.pos = int(z["py_state"][624])
    o.e.pyrand = m
    n = TO.MT()
    n.mt[:] = [int(x) for x in z["np_state"][:624]]
    n.pos = int(z["np_state"][624])
    o.e.nprand = n
    return o


def oracle_step(traj, o, t):
    """Advance the oracle by golden step t; returns (out, real_multi_or_None)."""
    use_np = not traj.meta["random_agent"]
    diff = traj.meta["difficulty"]
    if traj.kind == "def":
        if traj.multi:
            return o.def_step_multi(traj.def_action(t), diff, use_np)
        return o.def_step(traj.def_action(t), diff, use_np), None
    if traj.kind == "atk":
        return o.atk_step(traj.atk_action(t), diff, use_np), None
    if traj.multi:
        return o.multi_step_multi(traj.atk_action(t), traj.def_action(t))
    return o.multi_step(traj.atk_action(t), traj.def_action(t)), None


def check_outputs(traj, t, reward, done, win, allow, real_def, fail_def, real_atk, fail_atk, real_multi=None,
                  real_is_def_only=None):
    z, i, tag = traj.z, t - 1, "%s step %d" % (traj.name, t)
    assert float(reward).hex() == float(z["reward"][i]).hex(), "%s reward %r != %r" % (tag, reward, z["reward"][i])
    assert bool(done) == bool(z["done"][i]), tag + " done"
    assert int(win) == int(z["win"][i]), tag + " win"
    assert int(allow) == int(z["allow"][i]), tag + " allow_next %d != %d" % (allow, z["allow"][i])
    if traj.kind != "atk":
        if traj.multi:
            assert np.array_equal(real_multi, traj.real_multi(t)), tag + " RealAction (multi)"
        else:
            assert int(real_def) == int(z["real_def"][i]), tag + " RealAction defender"
            assert int(fail_def) == int(z["fail_def"][i]), tag + " FailCode defender"
    if traj.kind != "def":
        only = bool(z["real_is_def_only"][i])
        if real_is_def_only is not None:
            assert bool(real_is_def_only) == only, tag + " RealAction dict-vs-int quirk"
        if not only:
            assert np.array_equal(np.asarray(real_atk, dtype=np.int64), z["real_atk"][i].astype(np.int64)), \
                tag + " RealAction attacker"
        if not traj.multi:
            n = int(z["fail_atk"][i][0])
            assert list(fail_atk[:1 + n]) == list(z["fail_atk"][i][:1 + n]), tag + " FailCode attacker"
