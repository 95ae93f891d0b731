#!/usr/bin/env python
"""Config 5 of BASELINE.json in miniature: the batched env feeds a policy network in place.

    python examples/rollout_feed.py [--env TD-2p-large-v0] [--envs 16384] [--steps 50]
    torchrun --nproc-per-node 8 examples/rollout_feed.py            # one rank per GPU, stats all-reduced by NCCL

The observation tensor written by the step kernel is read directly by a small CNN (library convolutions:
the learner is not part of this repo's hot path); actions are sampled on the GPU and handed back to the next
step; nothing crosses PCIe.  A RolloutBuffer per player (one for TD-def / TD-atk, two for TD-2p) masks the actions
by AllowNextMove, records the penalised rewards and computes GAE at the end of every horizon -- the host loop of
the reference's trainer (train/main.py:79-176, train/PPO/Model.py:134-192) as device kernels.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import gym_td_b200 as G
from gym_td_b200 import dist as D
from gym_td_b200.rollout import RolloutBuffer


class TinyPolicy(torch.nn.Module):
    """45 x L x L observation -> defender logits over 6 L^2 + 1 actions, attacker logits (3, 8, 5), value."""

    def __init__(self, L):
        super().__init__()
        self.L = L
        self.body = torch.nn.Sequential(torch.nn.Conv2d(45, 32, 3, padding=1), torch.nn.ReLU(),
                                        torch.nn.Conv2d(32, 32, 3, padding=1), torch.nn.ReLU())
        self.defender = torch.nn.Conv2d(32, 6, 1)
        self.nop = torch.nn.Parameter(torch.zeros(1))
        self.attacker = torch.nn.Linear(32, 3 * 8 * 5)
        self.value = torch.nn.Linear(32, 1)

    def forward(self, obs):
        h = self.body(obs)
        pooled = h.mean(dim=(2, 3))
        d = torch.cat([self.defender(h).flatten(1), self.nop.expand(obs.shape[0], 1)], dim=1)
        a = self.attacker(pooled).view(-1, 3, 8, 5)
        return d, a, self.value(pooled).squeeze(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="TD-2p-large-v0")
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--horizon", type=int, default=16)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # the policy reads env.obs in place and nothing else writes it: let the step update it incrementally
    env = G.make_vec(args.env, args.envs, seed=0, device=local, env_offset=D.rank_env_offset(rank), incremental_obs=True)
    L, kind = env.map_size, env.kind
    policy = TinyPolicy(L).cuda(local).to(memory_format=torch.channels_last).eval()
    roles = {"def": ["defender"], "atk": ["attacker"], "2p": ["defender", "attacker"]}[kind]
    bufs = [RolloutBuffer(env, horizon=args.horizon, role=r, keep_actions=False) for r in roles]
    values = torch.zeros((args.horizon, args.envs), dtype=torch.float32, device=env.device)
    flushes = 0
    obs = env.reset()
    obs_ptr = obs.data_ptr()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for t in range(args.steps):
            d_logits, a_logits, v = policy(obs)                       # reads the env's observation tensor in place
            d_act = torch.distributions.Categorical(logits=d_logits.float()).sample()
            a_act = torch.distributions.Categorical(logits=a_logits.float()).sample()
            action = d_act if kind == "def" else a_act if kind == "atk" else {"Attacker": a_act, "Defender": d_act}
            values[bufs[0].ptr] = v.float()
            for b in bufs:
                b.mask(action)
            obs, reward, done, info = env.step(action)
            assert obs.data_ptr() == obs_ptr
            full = [b.record(action) for b in bufs]
            if full[0]:
                nv = policy(obs)[2].float()
                for b in bufs:                                  # zero-sum game: the attacker's critic is -V
                    b.flush(values if b.role == "defender" or kind != "2p" else -values,
                            nv if b.role == "defender" or kind != "2p" else -nv)
                flushes += 1
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stats = env.allreduce_stats()
    if rank == 0:
        import json
        print("%s: %d envs x %d ranks, %d steps with the policy in the loop: %.3g env-steps/s; episodes %d, wins %d"
              % (args.env, args.envs, world, args.steps, args.envs * world * args.steps / wall, stats["episodes"],
                 stats["wins"]))
        print(json.dumps({"env_id": args.env, "envs_per_gpu": args.envs, "ranks": world, "steps": args.steps,
                          "env_steps_per_sec_with_policy": args.envs * world * args.steps / wall,
                          "players_recorded": roles, "horizon": args.horizon, "gae_flushes": flushes,
                          "episode_stats": stats, "nccl": world > 1}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
