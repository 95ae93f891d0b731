#!/usr/bin/env python
"""Write-bandwidth context for the roofline: pure fills vs copy vs the observation kernel alone.

    python tools/membench.py [L] [n_envs]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_td_b200.vec_env import TDVecEnv

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda", 0)

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

out = {}
nbytes = N * 45 * L * L * 4
x = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
y = torch.empty_like(x)
out["zero_GBs"] = nbytes / timeit(lambda: x.zero_()) / 1e6
out["fill_GBs"] = nbytes / timeit(lambda: x.fill_(1.5)) / 1e6
out["copy_GBs_rw"] = 2 * nbytes / timeit(lambda: y.copy_(x)) / 1e6
big = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); big2 = torch.empty_like(big)
out["copy_2GiB_GBs_rw"] = 2 * big.numel() * 2 / timeit(lambda: big2.copy_(big), iters=10) / 1e6
out["zero_2GiB_GBs"] = big.numel() * 2 / timeit(lambda: big.zero_(), iters=10) / 1e6
del big, big2, y
env = TDVecEnv("def", L, N, seed=0, auto_reset=True)
env.reset()
acts = torch.randint(0, 6 * L * L + 1, (16, N), dtype=torch.int64, device=dev)
for k in range(1300): env.step(acts[k % 16])
s = torch.cuda.current_stream().cuda_stream
t_obs = timeit(lambda: env.engine.observe(env.obs, s))
out["observe_kernel_ms"] = t_obs
out["observe_GBs"] = nbytes / t_obs / 1e6
k = [0]
def full():
    env.step(acts[k[0] % 16]); k[0] += 1
t_full = timeit(full, iters=100)
out["step_ms"] = t_full
out["step_alg_GBs"] = (nbytes + N * 32) / t_full / 1e6
obs_keep = env.obs
env.obs = None
def logic():
    io = env._io(acts[k[0] % 16], None); env.engine.step(io, s); k[0] += 1
t_logic = timeit(logic, iters=100)
out["step_no_obs_ms"] = t_logic
print(json.dumps(out, indent=1))
