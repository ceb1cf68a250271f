"""Multi-GPU plumbing: contiguous env-index sharding and the one collective of the path, a
sum-reduction of per-rank episode statistics.  Environments never interact
(one physics world per env in the reference, environments.py:229), so there is no data-path
collective; torch.distributed is used with NCCL on GPUs and gloo in the CPU tests."""
import numpy as np

STAT_KEYS = ['env_steps', 'successes', 'reward_sum', 'resets', 'overflow_env_steps']


def shard_range(total_envs, rank, world):
    """Rank r of R owns [r*N/R, (r+1)*N/R) (remainder spread over the first ranks)."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(local, device=None):
    """local: dict with STAT_KEYS -> float.  Returns the job-wide sums on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64,
                     device=device if device is not None else 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(STAT_KEYS, t.tolist())}


def max_over_ranks(x, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else 'cpu')
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
