"""On-device rollout buffer for the reference's PPO loop.

Replaces the host-side bookkeeping of train/main.py:79-176 and train/PPO/{Callbacks.py:20-34, Model.py:134-192}
-- AllowNextMove masking, the RealAction penalty, record() and the GAE flush() -- with three small kernels
behind the C ABI (td_rollout_mask / td_rollout_record / td_gae).  Buffers are [horizon, n] torch tensors in HBM,
env index fastest; observations are not copied (the learner reads `env.obs` in place, or rebuilds them).

Two-player env (TD-2p-*, BASELINE config 5): one buffer per player (`role="defender"` / `"attacker"`).  The
reference's trainer never drives TDMulti; the bookkeeping applied per player is the one it applies to the
single-agent envs -- the defender sees the board reward and its Discrete RealAction, the attacker the negated
reward (TDAttack.py:50) and its (3, 8) cluster RealAction, each masked by its own AllowNextMove bit.
"""

import torch

from . import engine as E


class RolloutBuffer(object):
    def __init__(self, env, horizon=128, gamma=0.99, lam=0.95, penalty=0.3, keep_actions=True, role=None):
        if env.multi_action:
            raise ValueError("the rollout bookkeeping covers Discrete defender actions and (3, 8) attacker clusters")
        if env.kind == "2p":
            if role not in ("defender", "attacker"):
                raise ValueError("TD-2p needs one buffer per player: role='defender' or role='attacker'")
        elif role not in (None, {"def": "defender", "atk": "attacker"}[env.kind]):
            raise ValueError("a TD-%s env has no %s to record" % (env.kind, role))
        self.env, self.horizon, self.gamma, self.lam, self.penalty = env, int(horizon), gamma, lam, penalty
        self.role = role or {"def": "defender", "atk": "attacker"}[env.kind]
        self.which = 0 if self.role == "defender" else 1
        self.negate = env.kind == "2p" and self.which == 1       # the 2p env reports the defender's reward
        n, dev = env.num_envs, env.device
        self.rewards = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((horizon, n), dtype=torch.uint8, device=dev)
        self.advs = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        self.returns = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        shape = (horizon, n) if self.which == 0 else (horizon, n, E.ROADS, E.CLUSTER)
        self.actions = torch.zeros(shape, dtype=torch.int64, device=dev) if keep_actions else None
        self.ptr = 0

    def _stream(self):
        return torch.cuda.current_stream(self.env.device).cuda_stream

    def _own(self, actions):
        if isinstance(actions, dict):
            return actions["Defender" if self.which == 0 else "Attacker"]
        return actions

    def mask(self, actions):
        """train/main.py:130-132 -- in place: envs whose last AllowNextMove was False play empty_action()."""
        eng = self.env.engine
        actions = self._own(actions)
        eng._check(eng._lib.td_rollout_mask(eng._h, self.which, actions.data_ptr(), self.env._allow.data_ptr(),
                                            self._stream()))
        return actions

    def record(self, actions):
        """train/PPO/Callbacks.py:21-23 + Model.py:134-140, after env.step(actions): row `ptr` of the buffers."""
        env, eng, t = self.env, self.env.engine, self.ptr
        actions = self._own(actions)
        real = env.real_def if self.which == 0 else env.real_atk
        reward = env.reward.neg() if self.negate else env.reward
        arow = self.actions[t].data_ptr() if self.actions is not None else None
        eng._check(eng._lib.td_rollout_record(eng._h, self.which, actions.data_ptr(), real.data_ptr(),
                                              reward.data_ptr(), env._done.data_ptr(), float(self.penalty),
                                              self.rewards[t].data_ptr(), self.dones[t].data_ptr(), arow,
                                              self._stream()))
        self.ptr = (t + 1) % self.horizon
        return self.ptr == 0                      # True when the horizon is full (time to flush)

    def flush(self, values, next_value):
        """train/PPO/Model.py:166-192 -- GAE over the horizon.  values: [horizon, n] float32 critic outputs for
        the recorded states, next_value: [n] float32 for the states after the last step."""
        assert values.shape == self.rewards.shape and values.dtype == torch.float32 and values.is_contiguous()
        assert next_value.shape == (self.env.num_envs,) and next_value.dtype == torch.float32
        rc = E.lib().td_gae(self.horizon, self.env.num_envs, self.rewards.data_ptr(), self.dones.data_ptr(),
                            values.data_ptr(), next_value.data_ptr(), float(self.gamma), float(self.lam),
                            self.advs.data_ptr(), self.returns.data_ptr(), self._stream())
        if rc != 0:
            raise E.TdError(rc, E.lib().td_last_error(None).decode())
        return self.advs, self.returns
