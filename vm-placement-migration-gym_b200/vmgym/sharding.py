"""Multi-GPU plumbing: envs are closed systems (own state, own random streams), so N envs shard
contiguously over the ranks with no data-path collective.  Seeds derive from the GLOBAL env id, so the trajectory of
env i does not depend on the number of GPUs."""
from __future__ import annotations

import numpy as np


def shard_range(n_global: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the global env ids owned by `rank` (contiguous, sizes differ by at most one)."""
    base, extra = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_seeds(base_seed: int, n_global: int, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_range(n_global, rank, world)
    return int(base_seed) + np.arange(lo, hi, dtype=np.int64)


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (device timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
