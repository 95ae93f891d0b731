"""The CPU oracle against the reference's golden vectors (CPU only; runs everywhere).

Pins: (1) the reference's own self-test vector (gym_TD/envs/TDBoard.py:674-756, restated below as data),
(2) the nine trajectory digests quoted in SURVEY.md 8(c), re-generated from the reference into
tests/golden/survey_kats.json, and (3) full step-by-step trajectories recorded from the reference
(tests/golden/traj_*.npz) for every env kind, board size, opponent level and action mode.
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import td_oracle as TO
from tests import golden_util as GU

# gym_TD/envs/TDBoard.py:691-729: RandomState(1024), TDBoard(10, 2, ...)
ROAD0 = ["0000000000", "0000000000", "0000000000", "0000000000", "1100000111",
         "0100000100", "0111111100", "0000000000", "0000000000", "0000000000"]
ROAD1 = ["0000000000", "0000000000", "0000000000", "0000000000", "1100000000",
         "0100000000", "0111100000", "0000100000", "0000100000", "0000100000"]
DIST = [[0] * 10, [0] * 10, [0] * 10, [0] * 10, [0, 1, 0, 0, 0, 0, 0, 11, 12, 13], [0, 2, 0, 0, 0, 0, 0, 10, 0, 0],
        [0, 3, 4, 5, 6, 7, 8, 9, 0, 0], [0, 0, 0, 0, 7, 0, 0, 0, 0, 0], [0, 0, 0, 0, 8, 0, 0, 0, 0, 0],
        [0, 0, 0, 0, 9, 0, 0, 0, 0, 0]]


def _selftest_truth():
    r0 = np.array([[int(c) for c in row] for row in ROAD0], dtype=np.float32)
    r1 = np.array([[int(c) for c in row] for row in ROAD1], dtype=np.float32)
    road = ((r0 + r1) > 0).astype(np.float32)
    g = np.zeros((45, 10, 10), dtype=np.float32)
    g[0], g[1], g[2] = road, r0, r1
    g[4, 4, 0] = 1
    g[6, 4, 9] = 1
    g[7, 9, 4] = 1
    g[5] = 1
    g[9] = np.asarray(DIST, dtype=np.float32) / 14
    g[11] = 10 / 100
    g[12] = 0 / 100
    g[14] = 1 - road
    g[21] = 1
    for i, c in enumerate((8, 15, 40, 30)):
        g[41 + i] = 10 / c / 8
    return g


def test_reference_selftest_vector():
    z = np.load(os.path.join(GU.GOLDEN, "ref_selftest.npz"))
    truth = _selftest_truth()
    assert np.array_equal(z["obs0"], truth)          # the fixture really is the reference's golden vector
    o = TO.OracleEnv()
    o.init_from_planes(10, 2, [int(x) for x in z["start"]], int(z["end"]), z["road"], z["dist"].astype(np.int32),
                       z["dir"].astype(np.int32))
    assert np.array_equal(o.get_states().view(np.uint32), truth.view(np.uint32))
    assert TO.lib().tdo_n_channels() == 45                                 # TDBoard.py:151-152 doctest
    for t in range(4):                                                       # TDBoard.py:749-751
        for j in range(2):
            assert o.summon_enemy(t, j) is False
    assert not o.done()                                                      # TDBoard.py:374-382 doctests
    o.e.base_LP = 0
    assert o.done()
    o.e.base_LP, o.e.steps = 5, 1200
    assert o.done()


@pytest.mark.parametrize("path", GU.trajectories(), ids=lambda p: os.path.basename(p)[5:-4])
def test_oracle_replays_reference_trajectory(path):
    traj = GU.Traj(path)
    o = GU.oracle_for(traj)
    z = traj.z
    assert GU.digest64(o.get_states().tobytes()) == z["obs_digest"][0]
    samples = {int(s): z["sample_obs"][k] for k, s in enumerate(z["sample_steps"])}
    assert np.array_equal(o.get_states(), samples[0])
    for t in range(1, traj.T + 1):
        out, real_multi = GU.oracle_step(traj, o, t)
        allow = (1 if out.allow_next_def else 0) | (2 if out.allow_next_atk else 0)
        GU.check_outputs(traj, t, out.reward, out.done, out.win, allow, out.real_def, out.fail_def,
                         np.ctypeslib.as_array(out.real_atk), [out.n_fail_atk] + list(out.fail_atk),
                         real_multi, out.real_is_def_only if traj.kind == "2p" and not traj.multi else None)
        obs = o.get_states()
        assert GU.digest64(obs.tobytes()) == z["obs_digest"][t], "%s obs digest step %d" % (traj.name, t)
        assert GU.digest64(GU.state_bytes(o.state_dict())) == z["state_digest"][t - 1], \
            "%s state digest step %d" % (traj.name, t)
        if t in samples:
            assert np.array_equal(obs.view(np.uint32), samples[t].view(np.uint32))
    assert bool(z["done"][traj.T - 1]) == o.done()


def test_survey_known_answers():
    """SURVEY.md 8(c) digests, replayed with the product map generator + the oracle (seed 1024)."""
    from gym_td_b200 import mapgen
    kats = json.load(open(os.path.join(GU.GOLDEN, "survey_kats.json")))
    want = {("def", 10): "28ad3519638b7afb", ("def", 20): "34a65ea2e37c1181", ("def", 30): "795b27cb35ba9b82",
            ("2p", 10): "c5c470cc7b628c73", ("2p", 20): "dbf1b2c65e7b2604", ("2p", 30): "b7968388b787056c",
            ("atk", 10): "49ab55d27ea55ca9", ("atk", 20): "cb845aaaab242dd6", ("atk", 30): "008acd194e65c072"}
    for k in kats:
        assert want[(k["kind"], k["L"])] == k["sha"]      # the fixture equals the digests quoted in SURVEY.md
        L, kind = k["L"], k["kind"]
        env_rng = np.random.RandomState(1024)
        m = mapgen.generate_from_stream(env_rng, L)         # TDGymBasic.reset: num_roads, then the roads
        p = mapgen.planes(m)
        bits = (p["road"][0] | (p["road"][1] << 1) | (p["road"][2] << 2) | (p["road"][3] << 3)).astype(np.uint8)
        o = TO.OracleEnv()
        o.init_from_planes(L, p["num_roads"], p["start"], p["end"], bits, p["dist"], p["dir"])
        o.set_nprand(env_rng)
        random.seed(1024)
        o.set_pyrand(random.getstate())
        rs = np.random.RandomState(0)
        h = hashlib.sha256()
        h.update(o.get_states().tobytes())
        ret, n = 0.0, 0
        while True:
            if kind == "def":
                out = o.def_step(int(rs.randint(6 * L * L + 1)), 1, True)
            elif kind == "2p":
                atk = rs.randint(0, 5, size=(3, 8))
                out = o.multi_step(atk, int(rs.randint(6 * L * L + 1)))
            else:
                out = o.atk_step(rs.randint(0, 5, size=(3, 8)), 1, False)
            h.update(o.get_states().tobytes())
            h.update(np.float64(out.reward).tobytes())
            ret += out.reward
            n += 1
            if out.done:
                break
        assert (n, repr(ret), h.hexdigest()[:16]) == (k["steps"], k["ret"], k["sha"]), (kind, L)
