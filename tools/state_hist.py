"""List-length statistics of a running batch (how long are the tower / enemy lists the step kernel has to load?).
    python tools/state_hist.py def-small [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from gym_td_b200 import engine as E
from gym_td_b200.vec_env import TDVecEnv

name = sys.argv[1] if len(sys.argv) > 1 else "def-small"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1300
env_id, kind, L, n, multi, _ = B.WORKLOADS[name]
n = min(n, 16384)
env = TDVecEnv(kind, L, n, seed=0, auto_reset=True, multi_action=multi)
env.reset()
action, _, _ = B.make_actions(torch, name, kind, L, n, multi, env.device, 1234)
for k in range(steps):
    env.step(action(k))
    if (k + 1) % 325 == 0:
        torch.cuda.synchronize()
        blob = env.engine.get_state_raw(0, 4096)
        hdr = np.stack([b[:64].view(E.HEADER_DTYPE)[0] for b in blob])
        nt, ne = hdr["n_towers"].astype(int), hdr["n_enemies"].astype(int)
        q = lambda x: " ".join("p%d=%d" % (p, np.percentile(x, p)) for p in (50, 75, 90, 99, 100))
        print("%s step %4d towers mean %.1f %s | enemies mean %.1f %s | nt>8 %.2f nt>12 %.2f ne>6 %.2f ne>8 %.2f ne>10 %.2f"
              % (name, k + 1, nt.mean(), q(nt), ne.mean(), q(ne), (nt > 8).mean(), (nt > 12).mean(), (ne > 6).mean(),
                 (ne > 8).mean(), (ne > 10).mean()))
env.close()
