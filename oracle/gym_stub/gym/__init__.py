"""Minimal stand-in for OpenAI `gym` (not installed in this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference package
(`gym_TD`, found under baseline/_ref or /root/reference) can be imported and
used as the parity oracle / CPU baseline.  It provides exactly the surface the
reference touches ("Shim 1"):
  gym.Env, gym.spaces.{Box,Discrete,Dict}, gym.utils.seeding.np_random,
  gym.envs.registration.register, gym.make.
Seed contract: `seeding.np_random(seed)` returns a legacy
`numpy.random.RandomState(seed)` (the reference calls `.randint(low, high)`,
`.shuffle`, `.random` on it, i.e. the gym<=0.21 API).
"""
from . import spaces, utils, envs  # noqa: F401
from .envs.registration import make, register  # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return []
