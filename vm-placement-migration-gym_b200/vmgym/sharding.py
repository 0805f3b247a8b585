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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU with HOST-resident observations (HostVecEnv): run the rank — and therefore allocate its pinned
    buffers, first touch — on the NUMA node its GPU hangs off, so that eight ranks do not funnel their PCIe traffic through
    one socket's memory.  Returns the previous CPU affinity (to restore later) or None when the topology cannot be read."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None
