"""CPU restatement (numpy) of the reference's rollout bookkeeping -- TEST INFRASTRUCTURE ONLY.

Follows train/main.py:130-132 (masking), train/PPO/Callbacks.py:21-23 (penalty), train/PPO/Model.py:134-140
(record) and :166-192 (flush / GAE) with the arithmetic types NumPy 2 (NEP 50) gives those expressions:
float32 buffers, `gamma * v[j+1]` in float32, everything touched by the np.float64 `next_nonterminal` in
float64, float32 stores.  Pinned by tests/golden/rollout_*.npz, produced by the reference's own PPO class
(oracle/make_golden_rollout.py).
"""
import numpy as np


def mask_actions(actions, allow, empty):
    out = np.array(actions, copy=True)
    for i in range(len(out)):
        if not allow[i]:
            out[i] = empty
    return out


def record_row(actions, real, reward, done, penalty=0.3):
    """-> (rewards float32 [n], dones bool [n])"""
    r = np.array(reward, dtype=np.float64, copy=True)
    for i in range(len(r)):
        if np.any(np.asarray(actions[i]) != np.asarray(real[i])):
            r[i] -= penalty
    return r.astype(np.float32), np.asarray(done, dtype=bool)


def gae(rewards, dones, values, next_value, gamma, lam):
    """rewards/values float32 [T, n], dones bool [T, n], next_value float32 [n] -> advs, returns float32 [T, n]."""
    T, n = rewards.shape
    advs = np.zeros((T, n), dtype=np.float32)
    g32 = np.float32(gamma)
    gl = gamma * lam
    for i in range(n):
        last = 0.0
        for j in reversed(range(T)):
            nn = 1.0 - float(dones[j, i])
            if j == T - 1:
                g = gamma * float(next_value[i])
            else:
                g = float(np.float32(g32 * values[j + 1, i]))
            delta = (float(rewards[j, i]) + g * nn) - float(values[j, i])
            last = delta + (gl * nn) * last
            advs[j, i] = np.float32(last)
    return advs, (advs + values).astype(np.float32)
