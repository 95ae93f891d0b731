"""The N > 1 path on CPU: two gloo ranks take their env ranges by the rule bench.py and examples/rollout_feed.py use
(dist.rank_env_offset), and all-reduce the statistics vector and the max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as td
import torch.multiprocessing as mp

from gym_td_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    n_envs = 500 + rank
    idx = D.global_env_indices(rank, n_envs)
    lo, hi = idx[0], idx[-1] + 1
    assert lo == D.rank_env_offset(rank) == 1_000_000 * rank
    stats = dict(return_sum=0.5 * (hi - lo), episodes=hi - lo, length_sum=10 * (hi - lo), wins=rank,
                 kills=lo, leaks=hi, steps=1200 * (hi - lo))
    red = D.reduce_stats(stats, "cpu")
    t = torch.tensor([float(1 + rank)], dtype=torch.float64)      # the bench's max-over-ranks timing rule
    td.all_reduce(t, op=td.ReduceOp.MAX)
    out.put((rank, lo, hi, red, float(t)))
    td.destroy_process_group()


def test_two_rank_stats_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, lo0, hi0, red0, t0), (_, lo1, hi1, red1, t1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 500, 1_000_000, 1_000_501)        # disjoint global index ranges
    assert red0 == red1 and t0 == t1 == 2.0
    assert red0["episodes"] == 1001 and red0["return_sum"] == 500.5 and red0["steps"] == 1200 * 1001
    assert red0["wins"] == 1 and red0["kills"] == 1_000_000 and red0["leaks"] == 500 + 1_000_501
