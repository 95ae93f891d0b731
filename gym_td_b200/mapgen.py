"""Host map generation (C++ road generator in csrc/td_mapgen.cpp behind td_mapgen / td_mapgen_batch).

Contract: the map for `seed` is the one the reference builds from numpy.random.RandomState(seed) in
TDGymBasic.reset (num_roads drawn first from the same stream, gym_TD/envs/TDGymBasic.py:42-51).
Seeds for which the reference generator raises or does not terminate are invalid and skipped
(s <- s + 1), the rule the oracle harness applies too.
"""
import ctypes as C
import os

import numpy as np

from . import engine

DEFAULT_BUDGET = 100000


def generate(seed, map_size, num_roads=0, budget=DEFAULT_BUDGET):
    """Return a TdMap for RandomState(seed), or None when the seed is invalid."""
    m = engine.TdMap()
    rc = engine.lib().td_mapgen(int(seed) & 0xFFFFFFFF, int(map_size), int(num_roads), int(budget), C.byref(m))
    if rc < 0:
        raise engine.TdError(rc, "td_mapgen failed")
    return m if rc == 1 else None


def generate_from_stream(rs, map_size, num_roads=0, budget=DEFAULT_BUDGET):
    """Draw one map from a live numpy.random.RandomState (advanced in place), like TDGymBasic.reset does
    with self.np_random.  Returns a TdMap or None (stream advanced either way)."""
    st = rs.get_state()
    buf = np.empty(625, dtype=np.uint32)
    buf[:624] = st[1]
    buf[624] = st[2]
    m = engine.TdMap()
    rc = engine.lib().td_mapgen_stream(buf.ctypes.data, int(map_size), int(num_roads), int(budget), C.byref(m))
    if rc < 0:
        raise engine.TdError(rc, "td_mapgen_stream failed")
    rs.set_state((st[0], buf[:624].copy(), int(buf[624]), 0, 0.0))
    return m if rc == 1 else None


def generate_batch(seeds, map_size, num_roads=0, budget=DEFAULT_BUDGET, skip_invalid=True, threads=None):
    """Generate len(seeds) maps on host threads.

    Returns (maps, seeds_used, valid): a ctypes array of TdMap, the (possibly advanced) seeds, and a
    validity mask (all ones when skip_invalid).
    """
    seeds = np.array(seeds, dtype=np.uint32, copy=True).reshape(-1)
    n = seeds.shape[0]
    maps = (engine.TdMap * n)()
    valid = np.zeros(n, dtype=np.int32)
    threads = threads or min(os.cpu_count() or 1, 32)
    rc = engine.lib().td_mapgen_batch(seeds.ctypes.data, n, int(map_size), int(num_roads), int(budget),
                                      int(bool(skip_invalid)), int(threads), maps, valid.ctypes.data)
    if rc < 0:
        raise engine.TdError(rc, "td_mapgen_batch failed")
    return maps, seeds, valid


def planes(m):
    """TdMap -> dict of numpy planes in the reference's layout (map[0..5] of TDBoard.py:31-59)."""
    L = m.map_size
    cells = np.ctypeslib.as_array(m.cells)[:L * L].reshape(L, L)
    dist = np.ctypeslib.as_array(m.dist)[:L * L].reshape(L, L)
    road = np.stack([(cells >> k) & 1 for k in range(4)]).astype(np.int32)
    return dict(map_size=L, num_roads=m.num_roads, start=[m.start[i] for i in range(m.num_roads)], end=m.end,
                road=road, dist=dist.astype(np.int32), dir=((cells >> 4) & 3).astype(np.int32),
                max_dist=m.max_dist, n_randint=m.n_randint)
