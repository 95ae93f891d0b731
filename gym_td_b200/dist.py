"""Multi-GPU plumbing: env sharding and the episode-statistics reduction.

The board step has no exchange between instances: every rank steps its own env index
range and the only collective is one SUM all-reduce of a 7-element statistics vector per report
interval (NCCL on GPUs; gloo in the CPU tests)."""
import torch

STAT_KEYS = ["return_sum", "episodes", "length_sum", "wins", "kills", "leaks", "steps"]


def shard_range(n_global, rank, world):
    """Contiguous env index range [lo, hi) of `rank`; ranges differ in size by at most one."""
    base, rem = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_seeds(seed, lo, hi):
    """Global seeding rule: the env with global index g uses seed + g for its map search and its opponent."""
    return [(int(seed) + g) & 0xFFFFFFFF for g in range(lo, hi)]


def reduce_stats(stats, device="cpu"):
    """SUM-all-reduce a statistics dict over the default process group (no-op without one)."""
    import torch.distributed as dist
    v = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return {k: (float(x) if k == "return_sum" else int(round(x))) for k, x in zip(STAT_KEYS, v.tolist())}
