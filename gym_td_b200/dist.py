"""Multi-GPU plumbing: env sharding and the episode-statistics reduction.

The board step has no exchange between instances: every rank steps its own env index
range and the only collective is one SUM all-reduce of a 7-element statistics vector per report
interval (NCCL on GPUs; gloo in the CPU tests)."""
import torch

STAT_KEYS = ["return_sum", "episodes", "length_sum", "wins", "kills", "leaks", "steps"]


RANK_STRIDE = 1_000_000      # env i of GPU g uses seed 1_000_000 * g + i


def rank_env_offset(rank):
    """Global index of a rank's first env: the `env_offset` its TDVecEnv is created with.  Env i of rank r has the
    global index g = RANK_STRIDE * r + i; map seeds and opponent generators derive from seed + g, so a rank's
    results depend on its own global indices only -- never on the world size (bench.py, examples/rollout_feed.py)."""
    return RANK_STRIDE * int(rank)


def global_env_indices(rank, n_envs):
    lo = rank_env_offset(rank)
    return range(lo, lo + int(n_envs))


def reduce_stats(stats, device="cpu"):
    """SUM-all-reduce a statistics dict over the default process group (no-op without one)."""
    import torch.distributed as dist
    v = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return {k: (float(x) if k == "return_sum" else int(round(x))) for k, x in zip(STAT_KEYS, v.tolist())}
