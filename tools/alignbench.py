#!/usr/bin/env python
"""Observation-kernel bandwidth vs board size (sector alignment of planes / env stride)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_td_b200.vec_env import TDVecEnv
dev = torch.device("cuda", 0)
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for L in [10, 12, 16, 20, 24, 28, 30, 32]:
    N = int(1.18e9 // (45 * L * L * 4))
    env = TDVecEnv("def", L, N, seed=0, auto_reset=True, n_maps=2048)
    env.reset()
    s = torch.cuda.current_stream().cuda_stream
    nbytes = N * 45 * L * L * 4
    t = timeit(lambda: env.engine.observe(env.obs, s))
    x = env.obs.view(-1)
    tf = timeit(lambda: x.fill_(1.5))
    print(json.dumps(dict(L=L, N=N, plane_bytes=L*L*4, plane_mod32=(L*L*4) % 32, env_mod128=(45*L*L*4) % 128,
                          observe_ms=round(t, 4), observe_GBs=round(nbytes / t / 1e6), fill_GBs=round(nbytes / tf / 1e6))), flush=True)
    env.close(); del env, x
    torch.cuda.empty_cache()
