"""Host-side mirror of the reference interface: config API, registry, spaces, sharding."""
import numpy as np
import pytest

import gym_td_b200 as G
from gym_td_b200 import dist, params, spaces


def test_param_api_matches_reference_behaviour():
    assert G.getConfig() is params.config.__dict__               # TDParam.py:102-103 returns the live dict
    old = params.config.max_cost
    G.paramConfig(max_cost=77, brand_new_key=1)                  # bare setattr, unknown keys accepted (:98-100)
    assert params.config.max_cost == 77 and params.config.brand_new_key == 1
    G.paramConfig(max_cost=old)
    del params.config.brand_new_key
    hp = G.getHyperParameters()
    assert hp == dict(max_episode_steps=1200, video_frames_per_second=50, allow_multiple_actions=False,
                      max_cluster_length=8, max_num_of_roads=3)
    hp["max_episode_steps"] = 5                                  # a copy (:117-118)
    assert G.hyper_parameters.max_episode_steps == 1200
    with pytest.raises(RuntimeError):
        G.hyper_parameters.max_episode_steps = 3                 # :112-113
    assert params.n_channels() == 45


def test_registry_has_the_twelve_ids():
    ids = sorted(G.REGISTRY)
    assert len(ids) == 12 and "TD-def-small-v0" in ids and "TD-2p-v0" in ids
    assert G.REGISTRY["TD-atk-middle-v0"] == ("TDAttack", {"map_size": 20})
    assert G.REGISTRY["TD-def-v0"] == ("TDDefense", {})


def test_spaces_contains_semantics():
    d = spaces.Discrete(601)
    assert d.contains(600) and d.contains(np.int64(0)) and not d.contains(601) and not d.contains(1.0)
    b = spaces.Box(low=0, high=4, shape=(3, 8), dtype=np.int64)
    assert b.contains(np.full((3, 8), 4)) and not b.contains(np.full((3, 8), 5)) and not b.contains(np.zeros((3, 7)))
    assert not b.contains(np.zeros((3, 8), dtype=np.float64))
    dd = spaces.Dict({"Attacker": b, "Defender": d})
    assert dd.contains({"Attacker": np.zeros((3, 8), dtype=np.int64), "Defender": 3})
    assert not dd.contains({"Attacker": np.zeros((3, 8), dtype=np.int64)})
    assert b.sample().shape == (3, 8)


def test_rank_env_ranges_are_disjoint():
    """The sharding rule bench.py and examples/rollout_feed.py use: rank r owns global env indices
    [1_000_000 r, 1_000_000 r + n), whatever the world size."""
    ranges = [dist.global_env_indices(r, 65536) for r in range(8)]
    assert [rg[0] for rg in ranges] == [1_000_000 * r for r in range(8)] == [dist.rank_env_offset(r) for r in range(8)]
    assert all(a[-1] < b[0] for a, b in zip(ranges, ranges[1:]))


def test_distance_plane_division_is_correctly_rounded_for_every_operand():
    """The step kernel writes dist / (max dist + 1) with a hand-rolled float32 division (SmallDiv in
    csrc/td_obs.cuh: refined reciprocal, quotient, exact remainder, correction).  Exact rational
    arithmetic with round-to-nearest-even shows it returns the IEEE quotient for every 0 <= a <= 255,
    1 <= b <= 256 and for any hardware reciprocal within one ulp of 1 / b."""
    from fractions import Fraction

    def rn(x):                                      # Fraction -> nearest float32 (ties to even), as a Fraction
        if x == 0:
            return Fraction(0)
        s, x = (-1 if x < 0 else 1), abs(x)
        e = x.numerator.bit_length() - x.denominator.bit_length()
        if Fraction(2) ** e > x:
            e -= 1
        ulp = Fraction(2) ** (e - 23)
        q, r = divmod(x, ulp)
        q = int(q)
        if r * 2 > ulp or (r * 2 == ulp and q & 1):
            q += 1
        return s * q * ulp

    def fma(a, b, c):
        return rn(a * b + c)

    for b_int in range(1, 257):
        b = Fraction(b_int)
        r0 = rn(1 / b)
        ulp = Fraction(2) ** (r0.numerator.bit_length() - r0.denominator.bit_length() - 24)
        for x in (r0, r0 + ulp, r0 - ulp):         # rcp.approx: anything within 1 ulp
            r = fma(x, fma(-b, x, Fraction(1)), x)
            for a_int in range(0, 256):
                a = Fraction(a_int)
                q = rn(a * r)
                got = fma(fma(-b, q, a), r, q)
                assert got == rn(a / b), (a_int, b_int)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host cores (no GPU): one JSON line on stdout with the contract's
    keys, the metric/config of the B200 arm, and a zero-copy e2e object."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # nothing but the JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("TD-def-small-v0")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
