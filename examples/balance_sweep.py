#!/usr/bin/env python
"""The reference's balance study (balance.py) on the batched env: attacker win rate of every enemy type's
round-road script and of random clusters against the scripted defender, thousands of episodes per cell.

    python examples/balance_sweep.py [--env TD-atk-middle-v0] [--envs 4096] [--episodes 2]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import gym_td_b200 as G
from gym_td_b200 import balance as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="TD-atk-middle-v0")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--episodes", type=int, default=2)
    args = ap.parse_args()
    print("%-12s %-10s %9s %12s %10s" % ("script", "difficulty", "win rate", "mean return", "seconds"))
    for difficulty in (0, 1, 2):
        for name in ("type 0", "type 1", "type 2", "type 3", "random"):
            env = G.make_vec(args.env, args.envs, seed=0, difficulty=difficulty)
            agent = B.RandomAttacker(env) if name == "random" else B.RoundRoadAttacker(env, int(name[-1]))
            t0 = time.time()
            wins, rets = B.evaluate(env, agent, args.episodes)
            torch.cuda.synchronize()
            print("%-12s %-10d %9.3f %12.3f %10.1f" % (name, difficulty, wins.float().mean().item(),
                                                      rets.mean().item(), time.time() - t0))
            env.close()


if __name__ == "__main__":
    main()
