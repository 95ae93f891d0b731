"""Where the host-side time of td_step_host goes: the Python wrapper against the bare ctypes call, and the fixed
cost of a call (small batches).    python tools/e2e_breakdown.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_td_b200.vec_env import TDVecEnv


def run(n, chunks=0, chain=0, first=-1):
    env = TDVecEnv("def", 10, n, seed=0, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.randint(0, 601, (n,), dtype=torch.int64, device="cuda", generator=g) for _ in range(4)]
    for k in range(300):
        env.step(acts[k % 4])
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(200):
        env.step(acts[k % 4])
    e.record()
    torch.cuda.synchronize()
    dev = s.elapsed_time(e) / 200
    hacts = [a.cpu().pin_memory() for a in acts]
    env.engine.set_option("host_chunks", chunks)
    env.engine.set_option("host_chain", chain)
    env.engine.set_option("host_first_chunk", first)

    def timed(f, reps=200):
        for k in range(10):
            f(k)
        best = 1e9
        for r in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(reps):
                f(k)
            best = min(best, (time.perf_counter() - t0) / reps)
        return best * 1e3

    wrapper = timed(lambda k: env.step_host(hacts[k % 4]))
    # the bare C call with prebuilt structs (what a C caller pays)
    h = env._host_buffers()
    io = env._io(h["def_dev"], None)
    hio = env._hio_cache
    stream = torch.cuda.current_stream(env.device).cuda_stream
    ptrs = [a.data_ptr() for a in hacts]

    def bare(k):
        hio.def_action_host = ptrs[k % 4]
        env.engine.step_host(io, hio, stream)

    c_call = timed(bare)

    # floors: (a) device-resident action, one launch + one sync per step; (b) the kernel reads the action straight from
    # the page-locked host buffer (no copy node); (c) td_step_host without any output to the host
    from gym_td_b200 import engine as E
    dio = env._io(acts[0], None)

    def dev_sync(k):
        env.engine.step(dio, stream)
        torch.cuda.current_stream().synchronize()

    a_floor = timed(dev_sync)
    zio = E.TdStepIO.from_buffer_copy(dio)

    def zero_copy_in(k):
        zio.def_action_dev = ptrs[k % 4]
        env.engine.step(zio, stream)
        torch.cuda.current_stream().synchronize()

    b_floor = timed(zero_copy_in)
    hio2 = E.TdHostIO()

    def no_out(k):
        hio2.def_action_host = ptrs[k % 4]
        env.engine.step_host(io, hio2, stream)

    c_floor = timed(no_out)
    pio = E.TdStepIO.from_buffer_copy(dio)
    pio.packed_out_dev = h["packed"].data_ptr()

    def dev_in_zero_out(k):
        env.engine.step(pio, stream)
        torch.cuda.current_stream().synchronize()

    d_floor = timed(dev_in_zero_out)
    print("        floors: launch+sync, action on device %.4f (+%.1f us) | + packed outputs to host %.4f (+%.1f us) | action read from host memory by the kernel %.4f (+%.1f us) | td_step_host, no outputs %.4f (+%.1f us)"
          % (a_floor, (a_floor - dev) * 1e3, d_floor, (d_floor - dev) * 1e3, b_floor, (b_floor - dev) * 1e3, c_floor, (c_floor - dev) * 1e3))
    print("n=%6d chunks=%d chain=%d first=%5d: device %.4f ms | TDVecEnv.step_host %.4f ms (+%.1f us) | bare td_step_host %.4f ms (+%.1f us)"
          % (n, chunks, chain, first, dev, wrapper, (wrapper - dev) * 1e3, c_call, (c_call - dev) * 1e3))
    env.close()


for n in (256, 4096, 16384, 65536):
    run(n, 1, 1, 0)          # one chunk
for n in (16384, 65536):
    run(n)                   # automatic: independent chunks, short first chunk
