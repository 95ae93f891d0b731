"""Host map generator (C++, td_mapgen*) against maps produced by the reference's TDRoadGen/TDBoard."""
import os

import numpy as np
import pytest

from gym_td_b200 import mapgen
from tests import golden_util as GU


def _digest(m):
    L = m.map_size
    cells = np.ctypeslib.as_array(m.cells)[:L * L]
    dist = np.ctypeslib.as_array(m.dist)[:L * L]
    start = [m.start[i] if i < m.num_roads else 0 for i in range(3)]
    blob = cells.tobytes() + dist.tobytes() + np.asarray(start + [m.end], dtype="<i4").tobytes()
    return GU.digest64(blob)


@pytest.mark.parametrize("L", [10, 20, 30])
def test_golden_maps(L):
    z = np.load(os.path.join(GU.GOLDEN, "maps.npz"))
    valid, nroads, dig = z["valid_%d" % L], z["num_roads_%d" % L], z["digest_%d" % L]
    n = len(valid)
    maps, seeds, ok = mapgen.generate_batch(np.arange(n), L, skip_invalid=False, threads=4)
    assert np.array_equal(ok.astype(np.uint8), valid)            # same seeds are invalid (seed-skip rule)
    for s in range(n):
        if valid[s]:
            assert maps[s].num_roads == nroads[s], (L, s)
            assert _digest(maps[s]) == dig[s], (L, s)
    for k, s in enumerate(z["full_seeds_%d" % L]):
        p = mapgen.planes(maps[int(s)])
        cells = np.ctypeslib.as_array(maps[int(s)].cells)[:L * L].reshape(L, L)
        assert np.array_equal(cells, z["full_cells_%d" % L][k])
        assert np.array_equal(p["dist"], z["full_dist_%d" % L][k])


def test_skip_invalid_advances_seed():
    z = np.load(os.path.join(GU.GOLDEN, "maps.npz"))
    valid = z["valid_10"]
    bad = int(np.flatnonzero(valid == 0)[0])
    maps, seeds, ok = mapgen.generate_batch([bad], 10, skip_invalid=True)
    nxt = bad + 1
    while not valid[nxt]:
        nxt += 1
    assert int(seeds[0]) == nxt and ok[0] == 1
    assert mapgen.generate(bad, 10) is None


def test_stream_generation_continues_like_numpy():
    rs = np.random.RandomState(7)
    a = mapgen.generate_from_stream(rs, 20)
    b = mapgen.generate_from_stream(rs, 20)           # second reset of the same env: next draws of the stream
    assert a is not None and b is not None
    assert bytes(a.cells) == bytes(mapgen.generate(7, 20).cells)
    assert bytes(a.cells) != bytes(b.cells)


def test_map_invariants():
    maps, _, _ = mapgen.generate_batch(np.arange(200) + 5000, 30)
    for m in maps:
        p = mapgen.planes(m)
        L = 30
        road = p["road"][0]
        assert road.sum() < 6 * L and p["dist"][p["end"] // L, p["end"] % L] == 0
        assert p["max_dist"] == p["dist"].max() < 2 * L
        for i, s in enumerate(p["start"]):
            assert p["road"][i + 1][s // L, s % L] == 1
            # follow the direction field from the start: must reach the end in dist steps
            loc, steps = s, 0
            while loc != p["end"]:
                d = p["dir"][loc // L, loc % L]
                loc += (1, -1, L, -L)[d]
                steps += 1
                assert road[loc // L, loc % L] == 1 and steps < 2 * L
            assert steps == p["dist"][s // L, s % L]


def test_live_against_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference not present on this box")
    from oracle import ref_harness as RH
    ref_loader.load()
    from gym_TD.envs.TDBoard import TDBoard
    from gym.utils import seeding
    for L, n in ((10, 150), (20, 60), (30, 40), (13, 40)):
        for seed in range(900, 900 + n):
            rng = seeding.CountingRandomState(seed)
            rng.budget = RH.MAPGEN_BUDGET
            try:
                nr = int(rng.randint(low=1, high=4))
                b = TDBoard(L, nr, rng, 10, 0, 100, 5)
            except (ValueError, IndexError, seeding.BudgetExceeded):
                assert mapgen.generate(seed, L) is None, (L, seed)
                continue
            m = mapgen.generate(seed, L)
            assert m is not None, (L, seed)
            p, q = mapgen.planes(m), RH.board_roads(b)
            assert p["num_roads"] == q["num_roads"] and p["start"] == q["start"] and p["end"] == q["end"]
            bits = p["road"][0] | (p["road"][1] << 1) | (p["road"][2] << 2) | (p["road"][3] << 3)
            assert np.array_equal(bits.reshape(-1), q["road"]) and np.array_equal(p["dist"].reshape(-1), q["dist"])
            assert np.array_equal(p["dir"].reshape(-1), q["dir"]) and p["n_randint"] == rng.n_randint
