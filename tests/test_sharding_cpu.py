"""CPU: the N>1 path — contiguous pair sharding and the single gather of fixed-size
result records, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200slam.sharding import RECORD_WIDTH, gather_records, pack_records, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 256, 4541):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_items, rank, world)
    ids = np.arange(lo, hi)
    rec = pack_records(ids, ids * 3 + 1, ids % 5, ids * 2)
    out = gather_records(torch.from_numpy(rec), n_items)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 256])
def test_gather_world2_gloo(n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ids = np.arange(n_items)
    np.testing.assert_array_equal(got, pack_records(ids, ids * 3 + 1, ids % 5, ids * 2))
    assert got.shape == (n_items, RECORD_WIDTH)
