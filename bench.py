#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched gym-TD board step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload def-small] [--impl reference]

One "step" = one lockstep pass of the fused step kernel over the whole batch of game instances
of one GPU (action decode, scripted opponent, board dynamics, reward/done and the full
(45, L, L) float32 observation write), with finished instances restarted in place.
Under torchrun (N > 1) every rank steps its own batch (weak scaling, no data-path collective);
NCCL carries only the episode-statistics all-reduce.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (env id, kind, L, envs per GPU, multi_action, algorithmic bytes per env-step = SURVEY.md 8(d))
    "def-small": ("TD-def-small-v0", "def", 10, 65536, False, 4 * 45 * 100 + 8 + 24),
    "def-middle-multi": ("TD-def-middle-v0", "def", 20, 32768, True, 4 * 45 * 400 + 19200 + 19216),
    "def-middle-multi-sparse": ("TD-def-middle-v0", "def", 20, 32768, True, 4 * 45 * 400 + 19200 + 19216),
    "atk-small": ("TD-atk-small-v0", "atk", 10, 65536, False, 4 * 45 * 100 + 192 + 208),
    "2p-large": ("TD-2p-large-v0", "2p", 30, 16384, False, 4 * 45 * 900 + 200 + 216),
    "def-middle": ("TD-def-middle-v0", "def", 20, 32768, False, 4 * 45 * 400 + 8 + 24),
    "def-large": ("TD-def-large-v0", "def", 30, 16384, False, 4 * 45 * 900 + 8 + 24),
}
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="def-small", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--preroll", type=int, default=1300, help="untimed steps to de-synchronise episodes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=5, help="timed segments of --steps steps each; the median is reported")
    ap.add_argument("--replay", type=int, default=64, help="envs per rank replayed through the CPU oracle after timing (0 = off)")
    ap.add_argument("--replay-steps", type=int, default=100)
    ap.add_argument("--no-side-workloads", action="store_true", help="skip BASELINE configs 3-5 (atk-small, def-middle-multi, 2p-large)")
    ap.add_argument("--obs-memory", default="auto", choices=["auto", "compressible", "plain"],
                    help="memory of the observation tensor (TDVecEnv obs_memory): compressible allocation or ordinary torch memory")
    ap.add_argument("--side-workloads-multi", action="store_true", help="run the side workloads under torchrun too")
    return ap.parse_args()


def config_of(args, n_envs):
    env_id, kind, L, _, multi, bpe = WORKLOADS[args.workload]
    return {
        "workload": "%s batched, %d envs/GPU, %s actions, scripted opponent lv1, auto-reset from a map pool"
                    % (env_id, n_envs, ("Box(6,L,L) multi, flags 1 with p=0.01" if args.workload.endswith("sparse") else
                                          "Box(6,L,L) multi, uniform {0,1,2}") if multi else
                       ("Discrete" if kind == "def" else "cluster (3,8)" if kind == "atk" else "Dict")),
        "env_id": env_id, "map_size": L, "envs_per_gpu": n_envs,
        "global_envs": n_envs * args.gpus, "parallelism": "env-shard x%d" % args.gpus,
        "l2": "per-step output (%.0f MB) and env records exceed the 126 MB L2; no flush needed"
              % (n_envs * 45 * L * L * 4 / 1e6),
        "algorithmic_bytes_per_env_step": bpe,
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)

class Clocks(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline

def cpu_reference_sample(kind, L, steps_per_worker, pool=None):
    from oracle import cpu_baseline as CB
    own = pool is None
    pool = pool or CB.ReferencePool()
    n, wall = pool.run(kind, L, steps_per_worker)
    cores = pool.workers
    if own:
        pool.close()
    return n, wall, cores


def run_reference_arm(args):
    """The reference's own CPU implementation of the path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    env_id, kind, L, n_envs, multi, _ = WORKLOADS[args.workload]
    n_envs = args.envs or n_envs
    from oracle import cpu_baseline as CB
    cfg = config_of(args, n_envs)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 state / f32 observation", "data": "synthetic", "config": cfg}
    cores = os.cpu_count() or 1
    if CB.reference_available() and not multi:
        kind_name = "reference"
        per_worker = 1000                      # env-steps per worker per bench step (~0.15 s of Python)
        pool = CB.ReferencePool(cores)
        for _ in range(max(args.warmup, 1)):
            pool.run(kind, L, per_worker)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            n, _w = pool.run(kind, L, per_worker)
            total += n
        wall = time.perf_counter() - t0
        pool.close()
        sample = ("unmodified Python reference (gym_TD via gym stub), %d processes x %d env-steps per step, "
                  "%d steps, independent env instances, random actions" % (cores, per_worker, args.steps))
    else:
        kind_name = "port"
        per_thread = 20000
        for _ in range(max(min(args.warmup, 3), 1)):
            CB.run_port(L, per_thread, cores)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            n, _w = CB.run_port(L, per_thread, cores)
            total += n
        wall = time.perf_counter() - t0
        sample = "C restatement (oracle/td_oracle.c), %d threads x %d env-steps per step" % (cores, per_thread)
    value = total / wall
    line.update(value=value, ms_per_step=1e3 * wall / args.steps,
                cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": kind_name, "sample": sample},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                gpu_launches=0)
    emit(line)


# ------------------------------------------------------------------------------------------------

_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; keep it for the ONE JSON line only."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def pin_rank_to_cores(local, local_world):
    """One slice of the host's cores per rank, so that eight Python processes do not migrate over each other
    (td_step_host is bound by launch + sync latency on the host side)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(local_world, 1)
        if local_world > 1 and per >= 1:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
            return per
    except (AttributeError, OSError):
        pass
    return None


def make_actions(torch, name, kind, L, n_envs, multi, dev, seed):
    """Pre-generated synthetic actions, resident in HBM (excluded from every timed region)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    A = 16
    def_pool = atk_pool = None
    if kind != "atk":
        if multi:
            if name.endswith("sparse"):              # SURVEY 8(d) config 3, second variant: each flag 1 with p = 0.01
                def_pool = (torch.rand((4, n_envs, 6, L, L), device=dev, generator=g) < 0.01).to(torch.int64)
            else:                                    # action_space.sample(): uniform {0, 1, 2}
                def_pool = torch.randint(0, 3, (4, n_envs, 6, L, L), dtype=torch.int64, device=dev, generator=g)
        else:
            def_pool = torch.randint(0, 6 * L * L + 1, (A, n_envs), dtype=torch.int64, device=dev, generator=g)
    if kind != "def":
        atk_pool = torch.randint(0, 5, (A, n_envs, 3, 8), dtype=torch.int64, device=dev, generator=g)

    def action(k):
        d = def_pool[k % def_pool.shape[0]] if def_pool is not None else None
        a = atk_pool[k % atk_pool.shape[0]] if atk_pool is not None else None
        return d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}

    return action, def_pool, atk_pool


def timed_repeats(torch, dist, env, action, steps, repeats, world, dev, on_first=None):
    """`repeats` timed segments of exactly `steps` steps each, CUDA events on the launching stream, barrier +
    synchronize on both sides of every segment, max over ranks per segment.  Returns the list of ms."""
    out = []
    for r in range(repeats):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if r == 0 and on_first is not None:
            on_first()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for k in range(steps):
            env.step(action(k))
        end.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()))
    return out


def replay_check(torch, dist, env, action, n_sub, n_steps, world, dev):
    """Parity inside the benchmark run: restart the batch, step it at full size and replay the first n_sub envs of
    THIS rank through the CPU oracle (oracle/replay.py -- the checker, never the thing measured), every output
    and the whole observation bit for bit.  Counts are summed over ranks."""
    from oracle.replay import Replayer
    env.incremental_obs = False
    env.reset()
    torch.cuda.synchronize()
    R = Replayer(env, n_sub)
    R.check_initial_obs()
    for k in range(n_steps):
        a = action(k)
        env.step(a)
        torch.cuda.synchronize()
        R.check_step(a)
    s = R.summary()
    v = torch.tensor([s["replayed_envs"], s["compared_env_steps"], s["mismatches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return {"replayed_envs": int(v[0].item()), "replayed_steps": n_steps, "compared_env_steps": int(v[1].item()),
            "mismatches": int(v[2].item()), "first_mismatches": s["first_mismatches"], "ranks": world,
            "against": "oracle/td_oracle.c (C restatement pinned on the reference's golden vectors)",
            "what": "first %d envs of every rank, batch at full size, auto-reset on, all outputs + observation bit-exact"
                    % s["replayed_envs"]}


def median(xs):
    ys = sorted(xs)
    return ys[len(ys) // 2] if len(ys) % 2 else 0.5 * (ys[len(ys) // 2 - 1] + ys[len(ys) // 2])


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(name, compressed=False):
    """ncu DRAM bytes per launch of the workload's step kernel (profiles/traffic.json), by observation memory."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return tr.get(name, {}).get("compressible" if compressed else "plain", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def roofline_of(name, kind, n_envs, bytes_per_env_step, kernel_ms, obs_memory="plain"):
    peak, src = load_peak()
    achieved = bytes_per_env_step * n_envs / (kernel_ms * 1e-3) / 1e9
    compressed = obs_memory == "compressible"
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": load_traffic(name, compressed), "peak_source": src, "kernel": "td_step_kernel<%s>" % kind,
           "algorithmic_bytes_per_launch": bytes_per_env_step * n_envs}
    if compressed:
        out["note"] = ("the observation tensor is a compressible allocation: the kernel stores every algorithmic byte, L2 "
                       "compresses the lines on their way to HBM, so `traffic` (ncu DRAM bytes) is below the algorithmic bytes "
                       "and `frac` -- algorithmic bytes over time against the uncompressed copy peak -- can exceed 1; "
                       "obs_memory.plain has the same kernel on ordinary memory")
    return out


def side_workload(torch, dist, args, name, rank, world, local, dev):
    """One of BASELINE.json's other configs (3: def-middle multi-action, 4: atk-small, 5: 2p-large), timed the same
    way (device events, median of the repeats) plus a short oracle replay at the full batch size."""
    from gym_td_b200 import dist as D
    from gym_td_b200.vec_env import TDVecEnv
    env_id, kind, L, n_envs, multi, bpe = WORKLOADS[name]
    env = TDVecEnv(kind, L, n_envs, seed=args.seed, device=local, difficulty=1, auto_reset=True,
                   env_offset=D.rank_env_offset(rank), multi_action=multi, obs_memory=args.obs_memory)
    env.reset()
    action, _, _ = make_actions(torch, name, kind, L, n_envs, multi, dev, 1234 + rank)
    for k in range(args.preroll + args.warmup):
        env.step(action(k))
    steps = min(args.steps, 100)
    ms = timed_repeats(torch, dist, env, action, steps, 3, world, dev)
    m = median(ms) / steps
    out = {"env_id": env_id, "envs_per_gpu": n_envs, "value": n_envs * max(world, 1) / (m * 1e-3), "unit": UNIT,
           "ms_per_step": m, "steps": steps, "timed_repeats": len(ms),
           "repeat_ms_per_step": [x / steps for x in ms],
           "roofline": roofline_of(name, kind, n_envs, bpe, m, env.obs_memory), "obs_memory": env.obs_memory}
    if args.replay > 0:
        out["replay"] = replay_check(torch, dist, env, action, min(args.replay, 32), min(args.replay_steps, 60),
                                     world, dev)
    env.close()
    del env
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return 0

    import torch
    import torch.distributed as dist
    from gym_td_b200 import dist as D
    from gym_td_b200.vec_env import TDVecEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cores_per_rank = pin_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world if world > 1 else 1

    env_id, kind, L, n_envs, multi, bytes_per_env_step = WORKLOADS[args.workload]
    n_envs = args.envs or n_envs
    env = TDVecEnv(kind, L, n_envs, seed=args.seed, device=local, difficulty=1, auto_reset=True,
                   env_offset=D.rank_env_offset(rank), multi_action=multi, obs_memory=args.obs_memory)
    env.reset()
    action, def_pool, atk_pool = make_actions(torch, args.workload, kind, L, n_envs, multi, dev, 1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.preroll):
        env.step(action(k))
    for k in range(args.warmup):
        env.step(action(k))
    env.engine.reset_stats()
    barrier()

    clocks = Clocks(local)

    def start_clocks():
        if rank == 0:
            clocks.start()
            time.sleep(0.25)
            torch.cuda.synchronize()

    # W warm-up steps are done; now `repeats` segments of exactly K steps each (median of 5)
    seg_ms = timed_repeats(torch, dist, env, action, args.steps, args.repeats, world, dev, on_first=start_clocks)
    clk = clocks.stop() if rank == 0 else None
    ms = median(seg_ms)
    stats = env.allreduce_stats()          # NCCL: the only collective of the path
    value = n_envs * n_gpus * args.steps / (ms * 1e-3)

    # The observation tensor normally lives in a compressible allocation (TDVecEnv obs_memory, td_alloc_compressible):
    # the same kernel, the same values, fewer HBM bytes on the way out.  The same step into an ordinary torch tensor
    # is timed next to it so that the line carries both.
    obs_mem = {"headline": env.obs_memory}
    if rank == 0 and world == 1 and env.obs_memory == "compressible":
        keep = env.obs
        env.obs = torch.empty_like(keep)
        for k in range(args.warmup):
            env.step(action(k))
        torch.cuda.synchronize()
        s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s1.record()
        for k in range(args.steps):
            env.step(action(k))
        e1.record()
        torch.cuda.synchronize()
        ms1 = s1.elapsed_time(e1) / args.steps
        env.obs = keep
        env.step(action(0))
        obs_mem["plain"] = {"ms_per_step": ms1, "value": n_envs / (ms1 * 1e-3), "unit": UNIT,
                            "roofline_frac": roofline_of(args.workload, kind, n_envs, bytes_per_env_step, ms1)["frac"],
                            "note": "the same step writing into an ordinary (cudaMalloc / torch) tensor"}
        obs_mem["note"] = ("headline: observation tensor in a compressible allocation (cuMemCreate, "
                           "CU_MEM_ALLOCATION_COMP_GENERIC): lossless L2 compression on the way to HBM, transparent to "
                           "readers; DRAM traffic falls below the algorithmic bytes")

    inc = None          # measured on a fresh env further down

    # end to end through the host-buffer API: pinned host actions in, reward/done/info out, every step
    e2e = None
    e2e_obs = None
    if not args.no_e2e:
        hd = ha = None
        if kind != "atk":
            hd = [def_pool[i].cpu().pin_memory() for i in range(min(4, def_pool.shape[0]))]
        if kind != "def":
            ha = [atk_pool[i].cpu().pin_memory() for i in range(4)]

        def haction(k):
            d = hd[k % len(hd)] if hd is not None else None
            a = ha[k % len(ha)] if ha is not None else None
            return d if kind == "def" else a if kind == "atk" else {"Attacker": a, "Defender": d}

        ke = max(10, args.steps // 3)
        for k in range(3):
            env.step_host(haction(k))
        walls = []
        for r in range(3):
            barrier()
            t0 = time.perf_counter()
            for k in range(ke):
                out = env.step_host(haction(k))
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            tw = torch.tensor([wall], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            walls.append(float(tw.item()))
        h2d, d2h = env.host_bytes_per_step(False)
        e2e = {"value": n_envs * n_gpus * ke / median(walls), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": ke, "timed_repeats": len(walls),
               "note": "td_step_host: actions in pinned host memory -> device (8-byte Discrete actions are read by the "
                       "step kernel over PCIe, larger ones by copy nodes in front of independent chunk kernels), fused "
                       "step, reward/done/win/allow/RealAction/FailCode -> pinned host memory (one packed record per env, "
                       "stored by the kernel), stream sync every step (host wall clock, max over ranks, median of the "
                       "repeats); the observation stays in HBM for the on-device learner (see e2e_host_obs for the "
                       "variant that also copies it)"}
        if rank == 0 and world == 1:
            ko = 5
            env.step_host(haction(0), want_obs=True)
            t0 = time.perf_counter()
            for k in range(ko):
                env.step_host(haction(k), want_obs=True)
            wall = time.perf_counter() - t0
            h2d, d2h = env.host_bytes_per_step(True)
            e2e_obs = {"value": n_envs * ko / wall, "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": ko,
                       "note": "as e2e plus the full float32 observation copied to pinned host memory (PCIe-bound)"}

    replay = None
    if args.replay > 0:
        replay = replay_check(torch, dist, env, action, args.replay, args.replay_steps, world, dev)
    env.close()
    del env
    torch.cuda.empty_cache()

    # Opt-in variant (td_step_io.obs_incremental): the same float32 tensor, updated in place instead
    # of rewritten.  Reported next to the headline, never as the headline: `value` always writes all 45 planes.
    # A fresh env, as a user would create it (TDVecEnv picks the observation memory for the mode).
    if rank == 0 and world == 1:
        e1 = TDVecEnv(kind, L, n_envs, seed=args.seed, device=local, difficulty=1, auto_reset=True,
                      env_offset=D.rank_env_offset(rank), multi_action=multi, incremental_obs=True, obs_memory=args.obs_memory)
        e1.reset()
        for k in range(min(args.preroll, 400) + args.warmup):
            e1.step(action(k))
        torch.cuda.synchronize()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for k in range(args.steps):
            e1.step(action(k))
        e2.record()
        torch.cuda.synchronize()
        ms2 = s2.elapsed_time(e2)
        inc = {"value": n_envs * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
               "obs_memory": e1.obs_memory,
               "note": "observation updated in place (changed planes + old/new tower and enemy cells); "
                       "bit-identical tensor, fewer bytes written; not comparable to the algorithmic-bytes roofline"}
        e1.close()
        del e1
        torch.cuda.empty_cache()

    # Opt-in reduced-precision observation planes (td_step_io.obs_format): same layout, bf16 / u8
    # elements.  Reported next to the headline like the in-place update; never the headline.
    reduced = None
    if rank == 0 and world == 1 and not multi and L in (10, 20, 30):
        reduced = {}
        for fmt in ("bf16", "u8"):
            e2 = TDVecEnv(kind, L, n_envs, seed=args.seed, device=local, difficulty=1, auto_reset=True,
                          env_offset=D.rank_env_offset(rank), obs_format=fmt, obs_memory=args.obs_memory)
            e2.reset()
            for k in range(min(args.preroll, 400) + args.warmup):
                e2.step(action(k))
            torch.cuda.synchronize()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for k in range(args.steps):
                e2.step(action(k))
            e3.record()
            torch.cuda.synchronize()
            ms3 = s3.elapsed_time(e3) / args.steps
            reduced[fmt] = {"ms_per_step": ms3, "value": n_envs / (ms3 * 1e-3), "unit": UNIT,
                            "obs_bytes_per_env_step": 45 * L * L * e2.obs.element_size()}
            e2.close()
            del e2
            torch.cuda.empty_cache()
        reduced["note"] = ("observation written as bfloat16 (round-to-nearest-even of the float32 value) / uint8 "
                           "(rint(min(255 v, 255))); fewer bytes, not the reference tensor: not comparable to the headline")
    del def_pool, atk_pool
    torch.cuda.empty_cache()

    # BASELINE.json configs 3-5 next to the headline (N = 1 by default: they would triple every SCALE run)
    sides = {}
    if args.workload == "def-small" and not args.no_side_workloads and (world == 1 or args.side_workloads_multi):
        for name in ("atk-small", "def-middle-multi", "2p-large"):
            sides[name] = side_workload(torch, dist, args, name, rank, world, local, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    kernel_ms = ms / args.steps                      # one fused kernel per step, timed on its own stream
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 state / f32 observation", "data": "synthetic",
        "config": config_of(args, n_envs),
        "preroll_steps": args.preroll,
        "obs_memory": obs_mem,
        "timing": {"timed_repeats": len(seg_ms), "steps_per_repeat": args.steps, "statistic": "median",
                   "repeat_ms_per_step": [x / args.steps for x in seg_ms],
                   "spread": (max(seg_ms) - min(seg_ms)) / ms if ms > 0 else None},
        "gpu_launches": args.steps * len(seg_ms),
        "clocks": clk,
        "roofline": roofline_of(args.workload, kind, n_envs, bytes_per_env_step, kernel_ms, obs_mem["headline"]),
        "episode_stats": stats,
    }
    if cores_per_rank:
        line["host_cores_per_rank"] = cores_per_rank
    if inc is not None:
        line["incremental_obs"] = inc
    if reduced:
        line["reduced_precision_obs"] = reduced
    if e2e is not None:
        line["e2e"] = e2e
        if e2e_obs is not None:
            line["e2e_host_obs"] = e2e_obs
    if replay is not None:
        line["replay"] = replay
    if sides:
        line["workloads"] = sides
    if n_gpus == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as CB
        cores = os.cpu_count() or 1
        if CB.reference_available() and not multi:
            per_worker = 6000                          # ~1 s of Python per worker
            pool = CB.ReferencePool(cores)
            pool.run(kind, L, 300)
            n, wall = 0, 0.0
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 12.0:
                a, b = pool.run(kind, L, per_worker)
                n += a
                wall += b
            pool.close()
            line["cpu_baseline"] = {"value": n / wall, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": "unmodified Python reference %s via gym stub: %d processes, %d env-steps "
                                              "in %.1f s, independent instances, random actions" % (env_id, cores, n, wall)}
        n, wall = CB.run_port(L, 400000, cores)
        port = {"value": n / wall, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "C restatement oracle/td_oracle.c, TD-def L=%d, %d threads x 400000 env-steps" % (L, cores)}
        if "cpu_baseline" in line:
            line["cpu_port"] = port
        else:
            line["cpu_baseline"] = port
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
