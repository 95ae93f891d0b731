"""td_step_host throughput against the chunk count and the launch mode (graph / plain streams).
    python tools/e2e_sweep.py [workload]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from gym_td_b200.vec_env import TDVecEnv

name = sys.argv[1] if len(sys.argv) > 1 else "def-small"
env_id, kind, L, n, multi, _ = B.WORKLOADS[name]
env = TDVecEnv(kind, L, n, seed=0, auto_reset=True, multi_action=multi)
env.reset()
action, dp, ap = B.make_actions(torch, name, kind, L, n, multi, env.device, 1234)
for k in range(300):
    env.step(action(k))
hd = [dp[i].cpu().pin_memory() for i in range(4)] if dp is not None else None
ha = [ap[i].cpu().pin_memory() for i in range(4)] if ap is not None else None


def haction(k):
    d = hd[k % 4] if hd is not None else None
    a = ha[k % 4] if ha is not None else None
    return d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}


s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for k in range(100):
    env.step(action(k))
e.record()
torch.cuda.synchronize()
dev_ms = s.elapsed_time(e) / 100
print("%s device step %.4f ms = %.3e env-steps/s" % (name, dev_ms, n / dev_ms * 1e3))
def measure(tag):
    for k in range(8):
        env.step_host(haction(k))
    best = 1e9
    for r in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(60):
            env.step_host(haction(k))
        best = min(best, (time.perf_counter() - t0) / 60)
    print("%s: %.4f ms/step  %.3e env-steps/s  (%.2f of device)" % (tag, best * 1e3, n / best, dev_ms / (best * 1e3)))


quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
measure("default (automatic: zero-copy inputs where they pay, independent chunks for copied inputs)")
env.engine.set_option("host_zero_copy", 0)      # the sweeps below are about the copy path
if not quick:
    for graph in (1, 0):
        for chunks in (1, 2, 4):
            env.engine.set_option("host_graph", graph)
            env.engine.set_option("host_chunks", chunks)
            measure("graph=%d chunks=%d" % (graph, chunks))
# independent chunk kernels (each starts when its own actions are in); first > 0: chunks of first, 3 first, 9 first, ...
env.engine.set_option("host_graph", 1)
for chain in ((1, 0) if quick else (0,)):
    env.engine.set_option("host_chain", chain)
    for chunks in ((1, 2, 3, 4, 8, 16) if quick else (2, 3, 4, 5, 6, 8)):
        for first in ((0, n // 32, n // 16, n // 8) if quick else (0, 512, 1024, 2048, 4096, 8192)):
            if chunks == 1 and first:
                continue
            env.engine.set_option("host_chunks", chunks)
            env.engine.set_option("host_first_chunk", first)
            measure("graph=1 %s chunks=%d first=%d" % ("chained" if chain else "unchained", chunks, first))
env.close()
