"""CPU suite: the N>1 host logic (env sharding, seed derivation, max-over-ranks timing) under a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vmgym.sharding import max_over_ranks, shard_range, shard_seeds


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 2), (65536, 8), (10, 4), (3, 8)])
def test_shards_partition_the_env_ids(n, world):
    ids = np.concatenate([np.arange(*shard_range(n, r, world)) for r in range(world)])
    assert np.array_equal(ids, np.arange(n))
    seeds = np.concatenate([shard_seeds(7, n, r, world) for r in range(world)])
    assert np.array_equal(seeds, 7 + np.arange(n))       # independent of the number of ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_global, rank, world)
        mine = torch.zeros(n_global, dtype=torch.int64)
        mine[lo:hi] = shard_seeds(100, n_global, rank, world).tolist().__len__() and torch.from_numpy(shard_seeds(100, n_global, rank, world))
        dist.all_reduce(mine)                                   # every id owned by exactly one rank
        assert torch.equal(mine, 100 + torch.arange(n_global))
        got = max_over_ranks([1.0 + rank, 5.0 - rank])
        assert got == [float(world), 5.0]
        # whole-job throughput = units of all ranks / slowest rank's time
        units = torch.tensor([hi - lo], dtype=torch.float64)
        dist.all_reduce(units)
        assert int(units.item()) == n_global
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 4097, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
