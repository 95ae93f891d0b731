"""Quick GPU bring-up driver (not a pytest file): python tests/gpu_quick.py"""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_util as PU
from oracle import td_oracle
td_oracle.build()
cases = [
    dict(kind="def", L=10, n_envs=8, steps=300, seed=1, opponent="none"),
    dict(kind="def", L=10, n_envs=8, steps=300, seed=1, opponent="stream"),
    dict(kind="def", L=10, n_envs=16, steps=1200, seed=2, opponent="device"),
    dict(kind="atk", L=10, n_envs=16, steps=600, seed=3, opponent="none"),
    dict(kind="atk", L=10, n_envs=16, steps=1200, seed=3, opponent="device"),
    dict(kind="2p", L=10, n_envs=16, steps=1200, seed=4, opponent="none"),
    dict(kind="def", L=20, n_envs=16, steps=600, seed=5, opponent="device", multi=True),
    dict(kind="2p", L=30, n_envs=8, steps=600, seed=6, opponent="none", multi=True),
    dict(kind="def", L=15, n_envs=8, steps=300, seed=7, opponent="device"),
]
ok = True
for c in cases:
    t = time.time()
    try:
        n = PU.run_parity(**c)
        print("OK  ", c, n, "env-steps %.1fs" % (time.time() - t), flush=True)
    except Exception as e:
        ok = False
        print("FAIL", c, repr(e)[:600], flush=True)
        traceback.print_exc(limit=3)
sys.exit(0 if ok else 1)
