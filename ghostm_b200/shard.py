"""db-chunk sharding across GPUs: chunk-parallel front, query-parallel back, one exchange.

The reference runs one db chunk after the other on one device and applies Merge once per
(db chunk, candidate chunk) in ascending db order, carrying the per-query hit lists between
calls (aligner.cpp:114-174); the carried hits take part in the unstable sort, so the result of
a query depends on that order - but only on ITS OWN candidates and carried hits (Merge works
per same-name run, aligner.cpp:697-742).  That gives two independent axes:

  front  seed search + SW extension of db chunk c: independent of every other chunk.  Chunk c
         lives on rank c % N (the index, 4 B per residue, is never replicated).
  back   Merge + TraceBack of one query slice: independent of every other slice, needs the
         scored candidates of ALL chunks in ascending chunk order, the .pos tables (DB::GetID)
         and the residues (TraceBack windows) - 1 B per residue, replicated on every rank.

Between them every rank hands every other rank the candidates (score, end) of that
rank's query slice: one all-to-all per round of N chunks, 8 B per candidate over NVLink.  The
slices are cut at same-name run boundaries; every slice then sees, for every chunk, exactly
the Merge calls (candidate-chunk segments, aligner.cpp:131-171) the single-device run makes,
restricted to its own queries - including the calls that bring it no candidate, because Merge
re-sorts the carried lists on every call (aligner.cpp:702).  All ranks work on the same step at
the same time: there is no pipeline to fill and the serial part of a step is the exchange.

The transport is torch.distributed (NCCL on device tensors, gloo on CPU tensors in the CPU
tests); the engines are abstract so that the host logic is testable without a GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

MAX_SEGMENTS = 15     # candidate chunks per (query chunk, db chunk) the meta exchange carries
BLOCK_WORDS = 2       # words per exchanged candidate: SW score and forward end.  The candidate's
                      # region start is not needed behind the exchange: Merge only copies it into
                      # the hit and TraceBack overwrites it (aligner.cpp:941)

Segment = Tuple[int, int]


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> List[int]:
    """db chunks whose index lives on `rank` (round robin: chunk c on rank c % world)."""
    return list(range(rank, n_chunks, world))


def slice_bounds(name_break, n_queries: int, world: int) -> np.ndarray:
    """world + 1 ascending query indices cutting [0, n_queries) into slices of about equal size
    whose boundaries fall on same-name run starts (name_break[i] != 0 or i == 0; None = every
    query is its own run)."""
    if name_break is None:
        starts = np.arange(n_queries, dtype=np.int64)
    else:
        nb = np.asarray(name_break).astype(bool).copy()
        if n_queries:
            nb[0] = True
        starts = np.flatnonzero(nb[:n_queries]).astype(np.int64)
    bounds = np.zeros(world + 1, dtype=np.uint32)
    for r in range(1, world):
        target = (r * n_queries + world // 2) // world
        i = int(np.searchsorted(starts, target, side="left"))
        b = int(starts[i]) if i < starts.shape[0] else n_queries
        bounds[r] = max(b, int(bounds[r - 1]))
    bounds[world] = n_queries
    return bounds


class Front:
    """Seed search + SW extension side of one rank."""

    def prepare(self, chunk_id: int) -> List[Segment]:
        """Search + score chunk_id for all queries; returns the candidate-chunk segments
        [(first_query, end_query)] in the order the reference would Merge them."""
        raise NotImplementedError

    def pack(self, bounds: np.ndarray):
        """-> (counts int32 tensor [n_queries], data int32 tensor [2 * total], totals uint64
        ndarray [len(bounds) - 1]); data holds one [score | end] block per slice."""
        raise NotImplementedError

    def empty(self, bounds: np.ndarray):
        """What pack returns for a rank that owns no chunk in this round."""
        raise NotImplementedError


class Back:
    """Merge + TraceBack side of one rank; its queries are the rank's slice, renumbered from 0."""

    def install(self, chunk_id: int, counts, data, total: int) -> None:
        raise NotImplementedError

    def merge(self, first: int, end: int) -> None:
        raise NotImplementedError

    def finish(self) -> None:
        pass


def _meta_pack(segs: Sequence[Segment], totals: np.ndarray, world: int) -> np.ndarray:
    if len(segs) > MAX_SEGMENTS:
        raise RuntimeError(f"{len(segs)} candidate chunks in one db chunk exceed the {MAX_SEGMENTS} "
                           "the sharded driver exchanges: raise -l (max_list_length)")
    m = np.zeros(1 + 2 * MAX_SEGMENTS + world, dtype=np.int64)
    m[0] = len(segs)
    for i, (f, e) in enumerate(segs):
        m[1 + 2 * i], m[2 + 2 * i] = f, e
    m[1 + 2 * MAX_SEGMENTS:] = totals.astype(np.int64)
    return m


def _meta_unpack(m: np.ndarray):
    n = int(m[0])
    segs = [(int(m[1 + 2 * i]), int(m[2 + 2 * i])) for i in range(n)]
    return segs, m[1 + 2 * MAX_SEGMENTS:]


def exchange(dist, rank: int, world: int, bounds: np.ndarray, counts, data, totals: np.ndarray,
             segs: Sequence[Segment]):
    """All-to-all of one round: returns, per source rank s, (counts of my slice, the
    [score | end] block of my slice, its candidate count, the segments of s's chunk)."""
    import torch
    dev = counts.device
    meta = torch.from_numpy(_meta_pack(segs, totals, world)).to(dev)
    meta_all = torch.empty(world * meta.shape[0], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(meta_all, meta)
    meta_all = meta_all.cpu().numpy().reshape(world, -1)
    n_slice = int(bounds[rank + 1]) - int(bounds[rank])
    recv_tot = [int(_meta_unpack(meta_all[s])[1][rank]) for s in range(world)]
    counts_in = torch.empty(world * n_slice, dtype=counts.dtype, device=dev)
    dist.all_to_all_single(counts_in, counts, [n_slice] * world,
                           [int(bounds[p + 1]) - int(bounds[p]) for p in range(world)])
    data_in = torch.empty(BLOCK_WORDS * sum(recv_tot), dtype=data.dtype, device=dev)
    dist.all_to_all_single(data_in, data, [BLOCK_WORDS * m for m in recv_tot], [BLOCK_WORDS * int(t) for t in totals])
    inbox, off = [], 0
    for s in range(world):
        inbox.append((counts_in[s * n_slice:(s + 1) * n_slice], data_in[off:off + BLOCK_WORDS * recv_tot[s]],
                      recv_tot[s], _meta_unpack(meta_all[s])[0]))
        off += BLOCK_WORDS * recv_tot[s]
    return inbox


def exchange_local(outboxes, bounds: np.ndarray):
    """The same exchange between simulated ranks of one process (tests): outboxes[s] =
    (counts, data, totals, segs) of rank s -> inboxes[r] as `exchange` returns them on rank r."""
    world = len(outboxes)
    inboxes = []
    for r in range(world):
        inbox = []
        for s in range(world):
            counts, data, totals, segs = outboxes[s]
            off = BLOCK_WORDS * int(sum(int(t) for t in totals[:r]))
            m = int(totals[r])
            inbox.append((counts[int(bounds[r]):int(bounds[r + 1])].clone(),
                          data[off:off + BLOCK_WORDS * m].clone(), m, list(segs)))
        inboxes.append(inbox)
    return inboxes


def back_round(back: Back, inbox, round0: int, n_chunks: int, base: int, stop: int) -> None:
    """Merge the chunks of one round into the slice [base, stop), ascending chunk order."""
    for s, (cnt, data, m, segs) in enumerate(inbox):
        c = round0 + s
        if c >= n_chunks:
            break
        if not segs:
            continue                    # SearchNext returned an empty list: no Merge call
        back.install(c, cnt, data, m)
        for f, e in segs:               # one Merge call per candidate chunk, as on one device
            f2, e2 = min(max(f, base), stop), max(min(e, stop), base)
            if f2 >= e2:
                f2 = e2 = base          # no candidates for this slice: carried lists only
            back.merge(f2 - base, e2 - base)


def front_round(front: Front, chunk_id: int, n_chunks: int, bounds: np.ndarray):
    """-> the outbox (counts, data, totals, segs) of one rank for the round holding chunk_id."""
    if chunk_id < n_chunks:
        segs = front.prepare(chunk_id)
        counts, data, totals = front.pack(bounds)
        return counts, data, totals, segs
    counts, data, totals = front.empty(bounds)
    return counts, data, totals, []


def shard_step(front: Front, back: Back, dist, rank: int, world: int, n_chunks: int,
               bounds: np.ndarray, before_back=None, timers=None) -> None:
    """One query batch on this rank: per round of `world` chunks search + extend the owned
    chunk, exchange by query slice, merge the round's chunks into the own slice; finally
    TraceBack of the survivors.  `dist` is torch.distributed (unused when world == 1);
    `before_back`, if given, is called after every exchange (e.g. a stream synchronisation);
    `timers`, if a dict, accumulates host wall seconds per phase (front / exchange / back)."""
    import time
    base, stop = int(bounds[rank]), int(bounds[rank + 1])

    def lap(key, t0):
        if timers is not None:
            timers[key] = timers.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    for round0 in range(0, n_chunks, world):
        t0 = time.perf_counter()
        counts, data, totals, segs = front_round(front, round0 + rank, n_chunks, bounds)
        t0 = lap("front", t0)
        if world == 1:
            inbox = [(counts, data, int(totals[0]), segs)]
        else:
            inbox = exchange(dist, rank, world, bounds, counts, data, totals, segs)
        if before_back is not None:
            before_back()
        t0 = lap("exchange", t0)
        if stop > base:
            back_round(back, inbox, round0, n_chunks, base, stop)
        t0 = lap("back", t0)
    t0 = time.perf_counter()
    if stop > base:
        back.finish()
    lap("back", t0)


# ---- engines over the C ABI (GPU) ----------------------------------------------------------

class GpuFront(Front):
    """gm_search / gm_score / gm_candidates_pack of one context; buffers are torch CUDA tensors
    so that NCCL can send them."""

    def __init__(self, ctx, n_queries: int, capacity: int, device, stats=None):
        import torch
        self.ctx, self.stats, self.n = ctx, stats, n_queries
        self.launches = 0                # kernels launched by pack (scan + pack)
        self.counts = torch.zeros(n_queries, dtype=torch.int32, device=device)
        self.data = torch.empty(BLOCK_WORDS * capacity, dtype=torch.int32, device=device)

    def prepare(self, chunk_id: int) -> List[Segment]:
        from . import capi
        counts, _ = self.ctx.search(chunk_id, self.stats)
        segs, first = [], 0
        while True:                      # Aligner::Execute's candidate-chunk loop, aligner.cpp:131-171
            end, n, last = capi.chunk_rule(counts, first, self.ctx.opt.max_list_length)
            if n == 0:
                break
            self.ctx.score(first, end, n, self.stats, fetch=False)
            segs.append((first, end))
            if last:
                break
            first = end
        return segs

    def pack(self, bounds: np.ndarray):
        totals = self.ctx.candidates_pack(bounds, self.counts.data_ptr(), self.data.data_ptr(),
                                          self.data.numel())
        self.launches += 2
        return self.counts, self.data[:BLOCK_WORDS * int(totals.sum())], totals

    def empty(self, bounds: np.ndarray):
        self.counts.zero_()
        return self.counts, self.data[:0], np.zeros(len(bounds) - 1, dtype=np.uint64)


class GpuBack(Back):
    def __init__(self, ctx, stats=None):
        self.ctx, self.stats = ctx, stats
        self.launches = 0                # kernels launched by install (scan)

    def install(self, chunk_id: int, counts, data, total: int) -> None:
        self.ctx.candidates_import(chunk_id, counts.data_ptr(), data.data_ptr() if total else 0, total)
        self.launches += 1

    def merge(self, first: int, end: int) -> None:
        self.ctx.merge(first, end, self.stats)

    def finish(self) -> None:
        self.ctx.traceback_pending(self.stats)
