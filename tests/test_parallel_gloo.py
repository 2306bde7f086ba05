"""The N>1 host logic on CPU with the gloo backend, world_size 2: env sharding needs no collective, agent sharding
all-gathers the entity view each sub-step."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from social_navigation_pyenvs_b200 import parallel


def test_env_shard_partitions_every_env_once():
    for E in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = np.zeros(E, int)
            for r in range(world):
                sl = parallel.env_shard(E, r, world)
                seen[sl] += 1
            assert (seen == 1).all()
    assert parallel.agent_shard(65536, 3, 8) == (3 * 8192, 8192)
    with pytest.raises(ValueError):
        parallel.agent_shard(10, 0, 3)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 64
        full = torch.arange(5 * N, dtype=torch.float64).reshape(5, N)
        off, n = parallel.agent_shard(N, rank, world)
        buf = torch.full((5, N), -1.0, dtype=torch.float64)
        buf[:, off:off + n] = full[:, off:off + n]            # only the own columns are valid
        for _ in range(3):                                     # three "sub-steps": own slice changes, gather again
            parallel.all_gather_columns(buf, off, n, world)
            assert torch.equal(buf, full)
            full = full + 1.0
            buf[:, off:off + n] = full[:, off:off + n]
        t = parallel.max_over_ranks(1.0 + rank, "cpu", world)
        s = parallel.sum_over_ranks(10.0, "cpu", world)
        out[rank] = (t, s)
    finally:
        dist.destroy_process_group()


def test_agent_sharded_gather_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
        assert dict(out) == {0: (2.0, 20.0), 1: (2.0, 20.0)}
