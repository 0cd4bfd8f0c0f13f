"""Multi-rank plumbing on CPU: world_size-2 (and 3) gloo runs of the shard + all-gather path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pmvs_b200 import shard


def fake_records(ids):
    """Deterministic stand-in for refined patches: a pure function of the patch id (like the real path, whose RNG is
    keyed by id and whose results do not depend on batch position)."""
    ids = np.asarray(ids, dtype=np.float64)
    return np.stack([np.sin(ids), np.cos(ids), ids * 0.5, ids % 3, ids % 5, ids % 7, 1.0 / (1 + ids), (ids % 4 == 0) * 1.0], axis=1)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n, rank, world)
    local = torch.from_numpy(fake_records(np.arange(lo, hi)))
    full = shard.allgather_records(local, n, rank, world)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 101), (3, 64), (2, 1)])
def test_allgather_matches_single_process(world, n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = fake_records(np.arange(n))
    for rank, full in got:
        assert full.shape == want.shape and np.array_equal(full, want), rank


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65, 1000):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
