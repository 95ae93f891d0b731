"""Tiny parity run for compute-sanitizer (racecheck / memcheck): python tests/gpu_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_util as PU
from oracle import td_oracle
td_oracle.build()
n = 0
n += PU.run_parity("def", 10, n_envs=6, steps=120, seed=1, opponent="device")
n += PU.run_parity("atk", 10, n_envs=6, steps=120, seed=2, opponent="device", difficulty=2)
n += PU.run_parity("2p", 20, n_envs=4, steps=80, seed=3, opponent="none", multi=True)
n += PU.run_parity("def", 15, n_envs=4, steps=60, seed=4, opponent="device")
print("sanitize run ok", n)
