"""Golden vectors for the rollout consumer, produced by the reference's own train/PPO code.

TEST INFRASTRUCTURE ONLY.  python -m oracle.make_golden_rollout   (build container; needs /root/reference)
The reference classes are imported unmodified from /root/reference/train; a fixed critic stands in for the net.
"""
import os
import sys
import types

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    import torch
    from . import ref_loader
    ref_loader.load()
    sys.path.insert(0, os.path.join(ref_loader.find_reference_root() if os.path.isdir(
        os.path.join(ref_loader.find_reference_root(), "train")) else "/root/reference", "train"))
    if not os.path.isdir(sys.path[0]):
        sys.path[0] = "/root/reference/train"
    import PPO
    from PPO import Callbacks

    rs = np.random.RandomState(5)
    for name, T, n, act_shape in (("def", 16, 6, ()), ("atk", 12, 5, (3, 8))):
        cfg = types.SimpleNamespace(horizon=T, num_actors=n, gamma=0.99, lam=0.95, device="cpu", learning_rate=3e-4)
        state_shape = (3, 4, 4)

        class Critic(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.tensor(0.37))

            def forward(self, s):
                v = (s.float().mean(dim=(1, 2, 3)) * 3.1 - 0.4).unsqueeze(1)
                p = torch.zeros((s.shape[0], 601 if act_shape == () else 5) + tuple(act_shape))
                return p, v

        ppo = PPO.PPO(None, None, Critic(), state_shape, act_shape, cfg)
        states = rs.rand(T + 1, n, *state_shape).astype(np.float32)
        rewards_in = (rs.randn(T, n) * 2).astype(np.float64)
        dones = rs.rand(T, n) < 0.15
        if act_shape == ():
            actions = rs.randint(0, 601, size=(T, n)).astype(np.int64)
        else:
            actions = rs.randint(0, 5, size=(T, n) + act_shape).astype(np.int64)
        real = actions.copy()
        flip = rs.rand(T, n) < 0.3
        for t in range(T):
            for i in range(n):
                if flip[t, i]:
                    if act_shape == ():
                        real[t, i] = 600
                    else:
                        real[t, i, rs.randint(3), rs.randint(8)] = 4 if actions[t, i].flat[0] != 4 else 3
        for t in range(T):
            infos = [{"RealAction": real[t, i]} for i in range(n)]
            r = rewards_in[t].copy()
            Callbacks_rewards = r
            for i, action in enumerate(actions[t]):                       # PPO_train, Callbacks.py:21-23
                if (action != infos[i]["RealAction"]).any():
                    Callbacks_rewards[i] -= 0.3
            ppo.record(states[t], actions[t], Callbacks_rewards, dones[t])
        ppo.flush(torch.tensor(states[T]))
        with torch.no_grad():
            values = np.stack([Critic()(torch.tensor(states[t]))[1].numpy()[:, 0] for t in range(T + 1)])
        np.savez_compressed(os.path.join(OUT, "rollout_%s.npz" % name), actions=actions, real=real,
                            rewards_in=rewards_in, dones=dones, values=values[:T].astype(np.float32),
                            next_value=values[T].astype(np.float32), rewards=ppo._PPO__rewards,
                            advs=ppo._PPO__advs[:, :, 0], returns=ppo._PPO__returns[:, :, 0],
                            gamma=0.99, lam=0.95, penalty=0.3)
        print(name, "advs", ppo._PPO__advs[:2, :3, 0])
    return 0


if __name__ == "__main__":
    sys.exit(main())
