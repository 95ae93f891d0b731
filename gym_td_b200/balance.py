"""Batched scripted attack agents of the reference's balance study, on TDVecEnv tensors.

The reference's `balance.py` measures win rates of fixed attack scripts against the scripted defender
(`td_atk_random` balance.py:11-61, `td_atk_single_round_road` balance.py:63-123).  Its loops drive one env at
a time through `env.step`; they run unchanged through the façade classes in `envs.py`.  This module is the same
agent logic for N envs in lockstep: the action memory (`mem`), the road cursor and the per-episode bookkeeping
are device tensors, so a balance sweep is one `step` kernel plus a few elementwise ops per step.
"""
import torch

from . import engine as E
from . import params

FC_COST_SHORTAGE = 2        # gym_TD/utils/fail_code.py
FC_IMPOSSIBLE_CLUSTER = 1


def _fail_mask(fail_atk, code):
    """`code in info['FailCode']` for every env: fail_atk is (N, 4) = [n codes, code road 0..2]."""
    n = fail_atk[:, :1]
    idx = torch.arange(E.ROADS, device=fail_atk.device).unsqueeze(0)
    return ((fail_atk[:, 1:] == code) & (idx < n)).any(dim=1)


class RoundRoadAttacker(object):
    """balance.py:63-123: a full cluster of one enemy type, one road per action, round robin over the roads;
    an action that failed with COST_SHORTAGE is repeated until it goes through."""

    def __init__(self, env, enemy_type):
        if env.kind != "atk":
            raise ValueError("RoundRoadAttacker drives a TD-atk vector env")
        self.env, self.t = env, int(enemy_type)
        cfg = params.config
        self.num_enemy = int(min(cfg.max_cost // cfg.enemy_cost[self.t][0], params.hyper_parameters.max_cluster_length))
        N, dev = env.num_envs, env.device
        self.road = torch.zeros(N, dtype=torch.int64, device=dev)
        self.mem = torch.full((N, E.ROADS, E.CLUSTER), E.NT, dtype=torch.int64, device=dev)
        self.has_mem = torch.zeros(N, dtype=torch.bool, device=dev)
        self._slot = (torch.arange(E.CLUSTER, device=dev) < self.num_enemy).view(1, 1, E.CLUSTER)
        self._rows = torch.arange(E.ROADS, device=dev).view(1, E.ROADS, 1)

    def act(self, num_roads):
        """num_roads: (N,) int64 roads of every env's current map."""
        fresh = torch.where((self._rows == self.road.view(-1, 1, 1)) & self._slot,
                            torch.full_like(self.mem, self.t), torch.full_like(self.mem, E.NT))
        use_mem = self.has_mem.view(-1, 1, 1)
        action = torch.where(use_mem, self.mem, fresh).contiguous()
        nxt = self.road + 1
        nxt = torch.where(nxt >= num_roads, torch.zeros_like(nxt), nxt)
        self.road = torch.where(self.has_mem, self.road, nxt)
        self._last = action
        return action

    def observe(self, done, info):
        short = _fail_mask(info["FailCode"], FC_COST_SHORTAGE)
        self.mem = torch.where(short.view(-1, 1, 1), self._last, self.mem)
        self.has_mem = short & ~done
        self.road = torch.where(done, torch.zeros_like(self.road), self.road)      # a new episode starts at road 0


class RandomAttacker(object):
    """balance.py:11-61: uniform random clusters; COST_SHORTAGE repeats the action, IMPOSSIBLE_CLUSTER drops it."""

    def __init__(self, env, seed=0):
        if env.kind != "atk":
            raise ValueError("RandomAttacker drives a TD-atk vector env")
        self.env = env
        self.gen = torch.Generator(device=env.device)
        self.gen.manual_seed(seed)
        N, dev = env.num_envs, env.device
        self.mem = torch.full((N, E.ROADS, E.CLUSTER), E.NT, dtype=torch.int64, device=dev)
        self.has_mem = torch.zeros(N, dtype=torch.bool, device=dev)

    def act(self, num_roads=None):
        fresh = torch.randint(0, E.NT + 1, self.mem.shape, generator=self.gen, device=self.mem.device)
        self._last = torch.where(self.has_mem.view(-1, 1, 1), self.mem, fresh).contiguous()
        return self._last

    def observe(self, done, info):
        impossible = _fail_mask(info["FailCode"], FC_IMPOSSIBLE_CLUSTER)
        short = _fail_mask(info["FailCode"], FC_COST_SHORTAGE) & ~impossible
        self.mem = torch.where(short.view(-1, 1, 1), self._last, self.mem)
        self.has_mem = short & ~done


def roads_of(obs):
    """Number of roads of every env's current map, read from the start one-hot planes 6..8 of the observation."""
    return obs[:, 6:9].sum(dim=(1, 2, 3)).round().to(torch.int64)


def evaluate(env, agent, episodes):
    """Run until every env finished `episodes` episodes (auto-reset on).  Returns (wins, returns): (N, episodes)
    tensors, the per-episode attacker win flag and undiscounted reward sum -- balance.py's `wins` / `rwds`."""
    N, dev = env.num_envs, env.device
    wins = torch.zeros((N, episodes), dtype=torch.int8, device=dev)
    rets = torch.zeros((N, episodes), dtype=torch.float64, device=dev)
    count = torch.zeros(N, dtype=torch.int64, device=dev)
    acc = torch.zeros(N, dtype=torch.float64, device=dev)
    obs = env.reset()
    rows = torch.arange(N, device=dev)
    while True:
        num_roads = roads_of(obs)
        obs, reward, done, info = env.step(agent.act(num_roads))
        agent.observe(done, info)
        acc = acc + reward
        live = done & (count < episodes)
        slot = count.clamp(max=episodes - 1)
        rets[rows, slot] = torch.where(live, acc, rets[rows, slot])
        wins[rows, slot] = torch.where(live, info["Win"], wins[rows, slot])
        count = count + live.to(torch.int64)
        acc = torch.where(done, torch.zeros_like(acc), acc)
        if bool((count >= episodes).all()):
            return wins, rets
