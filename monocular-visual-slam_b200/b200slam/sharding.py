"""Multi-GPU path: frame pairs are independent, so ranks take contiguous blocks of the pair batch (or of the
map's keyframes) and the only collective is ONE all-gather of fixed-size result records.
One process per B200 (torchrun), NCCL over NVLink; the same classes run over gloo on CPU tensors (tests).

* ``shard_bounds``        who owns what.
* ``RecordGather``        the persistent all-gather buffer: every rank's last kernel writes its records straight
                          into its own slice (``local``), ``all_gather()`` is the in-place collective.
* ``ShardedFrontend``     match -> select -> RANSAC -> pose -> records on this rank's pairs + the gather, captured
                          as ONE CUDA graph (the NCCL all-gather is a node of it): BASELINE configs #2 / #3.
* ``ShardedSweep``        relocalization sweep (config #5): the map's keyframes sharded by id and resident on each
                          GPU, the query broadcast once, every rank verifies its local top-k, the gathered
                          candidates hold the global top-k.
"""
from __future__ import annotations

import numpy as np

RECORD_WIDTH = 4  # legacy int32 record: (n_matches, best_h, inlier_count, pair_id)


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_capacity(n_items: int, world: int) -> int:
    return -(-n_items // world) if world > 0 else n_items


def pack_records(pair_ids, n_matches, best_h, inlier_count):
    """-> (n_local, RECORD_WIDTH) int32 array (the small legacy record; the full record is csrc/records.cu)."""
    return np.stack([np.asarray(n_matches, np.int32), np.asarray(best_h, np.int32),
                     np.asarray(inlier_count, np.int32), np.asarray(pair_ids, np.int32)], axis=1)


def gather_records(local, n_items: int, group=None):
    """All-gather per-pair result records (torch tensor, (n_local, W), on the backend's device) from every rank;
    returns the (n_items, W) tensor ordered by pair id.  One fixed-size collective (shards padded to the largest)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    g = RecordGather(n_items, int(local.shape[1]), dtype=local.dtype, device=local.device, group=group, pad_value=-1)
    g.local[: local.shape[0]] = local
    g.all_gather()
    return g.ordered()


def shutdown_process_group(timeout_s: float = 20.0):
    """Destroy the default process group at the end of a run.  CUDA graphs that captured NCCL collectives must be
    released BEFORE their communicator is destroyed (ncclCommDestroy waits for them forever otherwise — measured: a
    15-minute hang): the caller drops its ShardedFrontend / ShardedSweep objects first, this collects what is left,
    and a timer ends the process normally if the teardown still does not return."""
    import gc
    import os
    import sys
    import threading

    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dist.barrier()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(timeout_s, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


class RecordGather:
    """Persistent gather buffer [world, cap, width]: ``local`` (this rank's [cap, width] slice) is where the last
    kernel of a step writes; ``all_gather()`` completes the other slices in place (NCCL recognises
    sendbuff == recvbuff + rank * count and moves nothing locally).  ``ordered()`` drops the padding."""

    def __init__(self, n_items: int, width: int, *, dtype=None, device=None, group=None, pad_value=0):
        import torch
        import torch.distributed as dist

        self.group = group
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.rank = dist.get_rank(group) if self.dist_on else 0
        self.n_items, self.width = int(n_items), int(width)
        self.cap = shard_capacity(self.n_items, self.world)
        self.lo, self.hi = shard_bounds(self.n_items, self.rank, self.world)
        self.buf = torch.full((self.world, self.cap, self.width), pad_value, dtype=dtype or torch.uint8, device=device)
        self.local = self.buf[self.rank]

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def all_gather(self, async_op: bool = False):
        """Enqueue the collective behind the current stream's work (capturable into a CUDA graph with NCCL).
        async_op=True returns the work handle instead of making the current stream wait for the collective: the
        caller's next kernels then overlap it, and `handle.wait()` orders whatever reads or rewrites the buffer."""
        if self.world > 1:
            import torch.distributed as dist

            return dist.all_gather_into_tensor(self.buf.view(-1), self.local.reshape(-1), group=self.group, async_op=async_op)
        return None

    def ordered(self):
        """[n_items, width] in pair order (a copy; padding rows of short shards dropped)."""
        import torch

        if self.cap * self.world == self.n_items:
            return self.buf.reshape(self.n_items, self.width)
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_items, r, self.world)
            parts.append(self.buf[r, : hi - lo])
        return torch.cat(parts, dim=0)


class ShardedFrontend:
    """The hot path on this rank's block of a global pair batch + the gather of every rank's result records.

    ``n_pairs_global`` pairs are split with ``shard_bounds``; the caller builds the PairBatch of ITS pairs
    (``self.lo .. self.hi``).  ``step(batch)`` runs match -> select -> hypotheses -> score -> winner (-> refit ->
    decomposition when cfg.with_pose), the record kernel writes into this rank's slice of the gather buffer, and the all-gather
    follows on the same stream.  ``capture(batch)`` records exactly that — collective included — into one CUDA
    graph (``replay()``); if the NCCL build cannot be captured the kernels are replayed and the collective is
    issued eagerly behind them (``gather_in_graph`` says which).  After a step every rank holds every pair's
    record (``records()`` -> device uint8 [n_pairs_global, record_bytes])."""

    def __init__(self, cfg, n_pairs_global: int, *, variant=None, group=None, device=None, n_buffers: int = 2):
        """n_buffers gather buffers are used in turn and the all-gather is asynchronous: the collective of launch set
        k runs under the kernels of launch set k + 1 (they do not depend on it); `wait()` joins what is outstanding."""
        from . import _capi
        from .frontend import Frontend, record_bytes

        torch = _capi.require_cuda()
        if not cfg.max_matches:
            raise ValueError("ShardedFrontend needs max_matches (record stride)")
        self.cfg = cfg            # cfg.with_pose decides whether the records carry R | t (else zeros)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fe = Frontend(self.cfg, variant=_capi.VARIANT_I8MMA1 if variant is None else variant)
        self.rec_bytes = record_bytes(cfg.max_matches)
        self.gathers = [RecordGather(n_pairs_global, self.rec_bytes, dtype=torch.uint8, device=self.dev, group=group)
                        for _ in range(max(1, n_buffers))]
        for g in self.gathers:
            g.buf.view(torch.int32).reshape(g.world, g.cap, -1)[:, :, 3] = -1   # padding slots: pair id -1
        self.gather = self.gathers[0]                  # the buffer of the most recent step
        self._works, self._k = [None] * len(self.gathers), 0
        self.lo, self.hi, self.world, self.rank = self.gather.lo, self.gather.hi, self.gather.world, self.gather.rank
        self._graph, self._batch, self.res = None, None, None
        self.gather_in_graph = False

    def step(self, batch, K=None):
        if batch.n_pairs != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns pairs [{self.lo}, {self.hi}) but the batch has {batch.n_pairs}")
        i = self._k % len(self.gathers)
        self._k += 1
        g = self.gathers[i]
        if self._works[i] is not None:             # the collective that last read this buffer must be done before it is rewritten
            self._works[i].wait()
            self._works[i] = None
        self.res = self.fe.run(batch, K=K, records=g.local[: batch.n_pairs], pair_id0=self.lo)
        self._works[i] = g.all_gather(async_op=True)
        self.gather = g
        return self.res

    def wait(self):
        """Order the current stream behind every outstanding all-gather (end of a step, before the records are read)."""
        for i, w in enumerate(self._works):
            if w is not None:
                w.wait()
                self._works[i] = None

    def capture(self, batches, K=None, collective_in_graph: bool = True):
        """One eager pass (lazy init, workspaces, NCCL warm-up), then the capture of step(b) for every b of `batches`
        (one PairBatch or a list: a step of several launch sets) followed by wait()."""
        import torch

        batches = batches if isinstance(batches, (list, tuple)) else [batches]

        def body():
            for b in batches:
                self.step(b, K)
            self.wait()
        body()
        torch.cuda.synchronize()
        self._k = 0
        g = torch.cuda.CUDAGraph()
        self.gather_in_graph = False
        if collective_in_graph and self.world > 1:
            try:
                with torch.cuda.graph(g):
                    body()
                self.gather_in_graph = True
            except Exception:                      # NCCL / torch build that cannot capture the collective
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
        if not self.gather_in_graph:
            self._eager_gathers = []
            with torch.cuda.graph(g):
                for j, b in enumerate(batches):
                    gg = self.gathers[j % len(self.gathers)]
                    self.res = self.fe.run(b, K=K, records=gg.local[: b.n_pairs], pair_id0=self.lo)
                    self.gather = gg
            self._eager_gathers = [self.gather]     # only the last launch set's records can be gathered behind the graph
        self._graph = g
        torch.cuda.synchronize()
        return self.gather_in_graph

    def replay(self):
        self._graph.replay()
        if not self.gather_in_graph:
            for gg in self._eager_gathers:
                gg.all_gather()
        return self.res

    def records(self):
        return self.gather.ordered()

    def close(self):
        """Release the captured graph (it holds the NCCL all-gather node: see shutdown_process_group)."""
        self.wait()
        self._graph = None


class ShardedSweep:
    """BASELINE config #5 on N GPUs: keyframes sharded by id (contiguous blocks), each shard resident on its GPU
    (``MapSweep``), the query frame broadcast from rank 0, every rank cross-check-matches the query against
    ITS keyframes, verifies its local top-k geometrically and contributes k candidate records + its per-keyframe
    match counts to ONE all-gather.  The global top-k by match count is among the world x k gathered candidates."""

    def __init__(self, kf_desc, kf_kp, frame_ids, cfg, *, top: int = 5, max_query_rows: int = 2048, group=None, device=None,
                 n_keyframes_global: int | None = None):
        """kf_desc / kf_kp / frame_ids: THIS rank's keyframes (lists of (n_i, 32) uint8 / (n_i, 2) float32, ids)."""
        from . import _capi
        from .frontend import MapSweep, record_bytes

        torch = _capi.require_cuda()
        import torch.distributed as dist

        self.group = group
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.rank = dist.get_rank(group) if self.dist_on else 0
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.sweep = MapSweep(kf_desc, kf_kp, frame_ids, cfg, top=top, max_query_rows=max_query_rows, device=self.dev)
        self.top, self.stride = top, cfg.max_matches
        self.n_global = n_keyframes_global if n_keyframes_global is not None else len(kf_desc) * self.world
        self.cap = shard_capacity(self.n_global, self.world)
        rb = record_bytes(cfg.max_matches)
        self.slot_bytes = top * rb + ((4 * self.cap + 63) // 64) * 64          # k candidate records | per-keyframe counts
        self.gather = RecordGather(self.world, self.slot_bytes, dtype=torch.uint8, device=self.dev, group=group)
        slot = self.gather.local[0]
        self.cand = slot[: top * rb].view(top, rb)
        self.counts = slot[top * rb: top * rb + 4 * self.cap].view(torch.int32)
        self.counts.fill_(-1)
        self._graph = None

    def query(self, q_desc_dev=None, q_kp_dev=None, n_rows: int | None = None):
        """Rank 0 passes the query frame (device uint8 [n, 32], float32 [n, 2]); the others pass None."""
        import torch
        import torch.distributed as dist

        sw = self.sweep
        if self.rank == 0 and q_desc_dev is not None:
            sw.set_query(q_desc_dev, q_kp_dev)
        if self.world > 1:                          # the 64 KB + 16 KB query, once
            dist.broadcast(sw.query_desc_slot(), src=0, group=self.group)
            dist.broadcast(sw.query_kp_slot(), src=0, group=self.group)
        sw.run(records=self.cand, counts_out=self.counts[: sw.n_kf])
        self.gather.all_gather()
        return self.gather.buf

    def capture(self):
        import torch

        self.query()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                self.query()
            self._graph = g
        except Exception:
            torch.cuda.synchronize()
            self._graph = None
        return self._graph is not None

    def replay(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self.query()

    def close(self):
        self._graph = None

    def result_host(self):
        """-> (per-keyframe match counts in global keyframe order, the global top-k candidates as a dict of arrays
        sorted by (-count, frame id))."""
        from .frontend import unpack_records

        buf = self.gather.buf.cpu().numpy().reshape(self.world, self.slot_bytes)
        rb = (self.slot_bytes - ((4 * self.cap + 63) // 64) * 64) // self.top
        counts, cands = [], []
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_global, r, self.world)
            counts.append(buf[r, self.top * rb: self.top * rb + 4 * self.cap].copy().view(np.int32)[: hi - lo])
            cands.append(buf[r, : self.top * rb].reshape(self.top, rb))
        rec = unpack_records(np.concatenate(cands, axis=0), self.stride)
        ok = np.flatnonzero(rec["pair_id"] >= 0)
        order = ok[np.lexsort((rec["pair_id"][ok], -rec["n_matches"][ok]))][: self.top]
        return np.concatenate(counts), {k: v[order] for k, v in rec.items()}
