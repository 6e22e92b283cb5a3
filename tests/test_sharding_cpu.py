"""CPU: the N>1 path — contiguous pair sharding and the single gather of fixed-size
result records, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200slam.sharding import RECORD_WIDTH, RecordGather, gather_records, pack_records, shard_bounds


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


def _worker_inplace(rank, world, port, n_items, width, q):
    """The product path's collective: every rank writes its records into ITS slice of the persistent buffer
    and ONE in-place all-gather completes the others; repeated steps reuse the buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = RecordGather(n_items, width, dtype=torch.uint8, device="cpu", pad_value=0)
    assert (g.lo, g.hi) == shard_bounds(n_items, rank, world) and g.local.shape == (g.cap, width)
    outs = []
    for step in range(3):
        ids = np.arange(g.lo, g.hi)
        rec = ((ids[:, None] * 7 + np.arange(width)[None] + step) % 251).astype(np.uint8)
        g.local[: g.n_local] = torch.from_numpy(rec)
        g.all_gather()
        outs.append(g.ordered().numpy().copy())
    if rank == 1:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items,width", [(7, 64), (256, 3584)])
def test_record_gather_in_place_world2_gloo(n_items, width):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_inplace, args=(r, 2, port, n_items, width, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ids = np.arange(n_items)
    for step, got in enumerate(outs):
        want = ((ids[:, None] * 7 + np.arange(width)[None] + step) % 251).astype(np.uint8)
        np.testing.assert_array_equal(got, want)


def test_unpack_records_layout():
    """Host view of the record layout of csrc/records.cu (header 16 x int32 | q u16 | t u16 | d u16 | inlier u8)."""
    from b200slam.frontend import unpack_records
    S, n = 5, 3
    rb = (64 + 7 * S + 63) // 64 * 64
    buf = np.zeros((n, rb), np.uint8)
    for r in range(n):
        hdr = np.array([4, 17 + r, 3, 100 + r] + [0] * 12, np.int32)
        hdr[4:13] = np.eye(3, dtype=np.float32).reshape(-1).view(np.int32)
        hdr[13:16] = np.array([0.0, 0.5, -1.0], np.float32).view(np.int32)
        buf[r, :64] = hdr.view(np.uint8)
        body = np.array([[1, 2, 3, 4, 0], [9, 8, 7, 6, 0], [30, 31, 32, 33, 0]], np.uint16) + r
        buf[r, 64:64 + 6 * S] = body.reshape(-1).view(np.uint8)
        buf[r, 64 + 6 * S:64 + 7 * S] = [1, 0, 1, 1, 0]
    u = unpack_records(buf, S)
    assert u["n_matches"].tolist() == [4, 4, 4] and u["best_h"].tolist() == [17, 18, 19] and u["pair_id"].tolist() == [100, 101, 102]
    np.testing.assert_array_equal(u["R"][1], np.eye(3, dtype=np.float32))
    np.testing.assert_array_equal(u["t"][2], np.array([0.0, 0.5, -1.0], np.float32))
    assert u["q"][2].tolist() == [3, 4, 5, 6, 2] and u["t_idx"][0].tolist() == [9, 8, 7, 6, 0] and u["d"][1].tolist() == [31, 32, 33, 34, 1]
    assert u["inlier"][0].tolist() == [1, 0, 1, 1, 0]


def test_shared_block_tables():
    """SharedBlocks.build (b2s_hamming_knn2_shared): tile prefix sums and per-pair first tiles for a
    frame sequence (pair p = frames p, p+1) and for a one-query-vs-map sweep — pure index arithmetic."""
    import numpy as np
    import torch
    from b200slam.frontend import SharedBlocks
    rows = np.array([300, 1, 128, 0, 257, 129], np.int32)
    row0 = (np.arange(6) * 320).astype(np.int32)
    sb = SharedBlocks.build(row0, rows, np.arange(5), np.arange(1, 6), torch.device("cpu"))
    tiles = [3, 1, 1, 0, 3, 2]
    assert sb.n_blocks == 6 and sb.total_tiles == sum(tiles) and sb.max_rows == 300
    assert sb.tile0.tolist() == [0, 3, 4, 5, 5, 8]
    assert sb.q_xtile.tolist() == [0, 3, 4, 5, 5] and sb.t_xtile.tolist() == [3, 4, 5, 5, 8]
    assert sb.row0.tolist() == row0.tolist() and sb.rows.tolist() == rows.tolist()
    # sweep: blocks = keyframes..., query last; every pair's query side is the last block
    rows = np.array([130, 128, 5, 2000], np.int32)
    sb = SharedBlocks.build(np.array([0, 130, 258, 263], np.int32), rows, np.full(3, 3), np.arange(3), torch.device("cpu"))
    assert sb.tile0.tolist() == [0, 2, 3, 4] and sb.total_tiles == 4 + 16
    assert sb.q_xtile.tolist() == [4, 4, 4] and sb.t_xtile.tolist() == [0, 2, 3]
