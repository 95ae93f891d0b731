"""Subprocess body of test_gpu_env.py::test_chunked_host_path_equals_device_path: argv = chunks, graph (0/1)
[, chain (0/1), envs in the first chunk (-1 = automatic), zero-copy inputs (-1 automatic / 0 / 1), 'pageable' = ordinary host memory for the actions]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_td_b200.vec_env import TDVecEnv


def main():
    for kind, L, N in (("def", 10, 1000), ("atk", 10, 514), ("2p", 20, 301)):
        a = TDVecEnv(kind, L, N, seed=9, auto_reset=True)
        b = TDVecEnv(kind, L, N, seed=9, auto_reset=True)
        a.reset()
        b.reset()
        b.engine.set_option("host_chunks", int(sys.argv[1]))
        b.engine.set_option("host_graph", int(sys.argv[2]))
        if len(sys.argv) > 3:
            b.engine.set_option("host_chain", int(sys.argv[3]))
        if len(sys.argv) > 4:
            b.engine.set_option("host_first_chunk", int(sys.argv[4]))
        if len(sys.argv) > 5:
            b.engine.set_option("host_zero_copy", int(sys.argv[5]))
        g = torch.Generator(device="cuda").manual_seed(1)
        for t in range(60):
            d = torch.randint(0, 6 * L * L + 1, (N,), device="cuda", generator=g)
            k = torch.randint(0, 5, (N, 3, 8), device="cuda", generator=g)
            act = d if kind == "def" else k if kind == "atk" else {"Attacker": k, "Defender": d}
            pin = (lambda t: t.cpu()) if "pageable" in sys.argv else (lambda t: t.cpu().pin_memory())
            hact = pin(d) if kind == "def" else pin(k) if kind == "atk" else {"Attacker": pin(k), "Defender": pin(d)}
            obs, rew, done, info = a.step(act)
            h = b.step_host(hact, want_obs=(t % 7 == 0))
            torch.cuda.synchronize()
            assert torch.equal(rew.cpu().view(torch.int64), h["reward"].view(torch.int64)), (kind, t)
            assert torch.equal(done.cpu(), h["done"].bool()) and torch.equal(a.win.cpu(), h["win"])
            assert torch.equal(a._allow.cpu(), h["allow"])
            if kind != "atk":
                assert torch.equal(a.real_def.cpu(), h["real_def"]) and torch.equal(a.fail_def.cpu(), h["fail_def"])
            if kind != "def":
                assert torch.equal(a.real_atk.cpu(), h["real_atk"]) and torch.equal(a.fail_atk.cpu(), h["fail_atk"])
            assert torch.equal(obs.view(torch.int32), b.obs.view(torch.int32))
            if t % 7 == 0:
                assert torch.equal(h["obs"].view(torch.int32), obs.cpu().view(torch.int32))
        a.close()
        b.close()
    print("chunked host path ok", sys.argv[1:])


if __name__ == "__main__":
    main()
