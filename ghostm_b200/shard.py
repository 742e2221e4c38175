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
the same time; the back stage of a round runs on its own host thread and stream (BackWorker),
overlapped with the front stage of the next round or batch, so the serial part is the exchange.

The transport is torch.distributed (NCCL on device tensors, gloo on CPU tensors in the CPU
tests); the engines are abstract so that the host logic is testable without a GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

INLINE_SEGMENTS = 4   # candidate chunks per (query chunk, db chunk) that ride in the first meta
                      # exchange; a rank with more (small -l) triggers a second, variable-length one
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
    """[n_segs | first INLINE_SEGMENTS segments | per-destination candidate totals]."""
    m = np.zeros(1 + 2 * INLINE_SEGMENTS + world, dtype=np.int64)
    m[0] = len(segs)
    for i, (f, e) in enumerate(segs[:INLINE_SEGMENTS]):
        m[1 + 2 * i], m[2 + 2 * i] = f, e
    m[1 + 2 * INLINE_SEGMENTS:] = totals.astype(np.int64)
    return m


def _meta_unpack(m: np.ndarray):
    n = int(m[0])
    segs = [(int(m[1 + 2 * i]), int(m[2 + 2 * i])) for i in range(min(n, INLINE_SEGMENTS))]
    return n, segs, m[1 + 2 * INLINE_SEGMENTS:]


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
    unpacked = [_meta_unpack(meta_all[s]) for s in range(world)]
    max_segs = max(u[0] for u in unpacked)
    all_segs = [list(u[1]) for u in unpacked]
    if max_segs > INLINE_SEGMENTS:     # any number of candidate chunks: a second, padded gather
        extra = np.zeros(2 * max_segs, dtype=np.int64)
        for i, (f, e) in enumerate(segs):
            extra[2 * i], extra[2 * i + 1] = f, e
        extra_all = torch.empty(world * extra.shape[0], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(extra_all, torch.from_numpy(extra).to(dev))
        extra_all = extra_all.cpu().numpy().reshape(world, -1)
        all_segs = [[(int(extra_all[s][2 * i]), int(extra_all[s][2 * i + 1])) for i in range(unpacked[s][0])]
                    for s in range(world)]
    n_slice = int(bounds[rank + 1]) - int(bounds[rank])
    recv_tot = [int(unpacked[s][2][rank]) for s in range(world)]
    counts_in = torch.empty(world * n_slice, dtype=counts.dtype, device=dev)
    dist.all_to_all_single(counts_in, counts, [n_slice] * world,
                           [int(bounds[p + 1]) - int(bounds[p]) for p in range(world)])
    data_in = torch.empty(BLOCK_WORDS * sum(recv_tot), dtype=data.dtype, device=dev)
    dist.all_to_all_single(data_in, data, [BLOCK_WORDS * m for m in recv_tot], [BLOCK_WORDS * int(t) for t in totals])
    inbox, off = [], 0
    for s in range(world):
        inbox.append((counts_in[s * n_slice:(s + 1) * n_slice], data_in[off:off + BLOCK_WORDS * recv_tot[s]],
                      recv_tot[s], all_segs[s]))
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


class BackWorker:
    """One host thread that runs the back stage (install + Merge + TraceBack calls) in submission
    order while the caller already drives the next front round: the C ABI calls block their
    calling thread only (ctypes drops the GIL), and the two stages live in different contexts on
    different streams.  Exceptions surface at the next submit() / drain()."""

    def __init__(self):
        import collections
        import queue
        import threading
        self._q = queue.Queue()
        self._err = None
        self._event = threading.Event
        self.round_tickets = collections.deque()   # shard_step: one ticket per submitted round
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            fn, ticket = self._q.get()
            try:
                if fn is None:
                    return
                if self._err is None:
                    fn()
            except BaseException as e:   # noqa: BLE001 - re-raised on the submitting thread
                self._err = e
            finally:
                ticket.set()
                self._q.task_done()

    def _check(self):
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def submit(self, fn):
        """Queue fn; -> a threading.Event that is set when fn has run (or was skipped after an
        earlier failure)."""
        self._check()
        ticket = self._event()
        self._q.put((fn, ticket))
        return ticket

    def drain(self) -> None:
        self._q.join()
        self._check()

    def close(self) -> None:
        self._q.put((None, self._event()))
        self._t.join()


def shard_step(front: Front, back: Back, dist, rank: int, world: int, n_chunks: int,
               bounds: np.ndarray, before_back=None, timers=None, worker: BackWorker = None,
               after_back=None, before_first_back=None, max_outstanding: int = 1) -> None:
    """One query batch on this rank: per round of `world` chunks search + extend the owned
    chunk, exchange by query slice, merge the round's chunks into the own slice; finally
    TraceBack of the survivors.  `dist` is torch.distributed (unused when world == 1).
    `before_back`, if given, is called after every exchange and returns an object whose
    .synchronize() is awaited before the back stage touches the received buffers (e.g. a CUDA
    event), or None.  With a `worker` the back stage of a round runs on that thread, overlapped
    with the next front round (and the next batch); `before_first_back` (e.g. the upload of the
    slice's queries into the back context) is queued in front of it and `after_back` (e.g. the
    result download) behind it; at most `max_outstanding` back jobs may be unfinished when a front
    round starts (GpuFront rotates max_outstanding + 1 pack buffers).  `timers`, if a dict,
    accumulates host wall seconds per phase."""
    import time
    base, stop = int(bounds[rank]), int(bounds[rank + 1])

    def lap(key, t0):
        if timers is not None:
            timers[key] = timers.get(key, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def run(fn):
        if worker is not None:
            return worker.submit(fn)
        fn()
        return None

    if before_first_back is not None:
        run(before_first_back)
    for round0 in range(0, n_chunks, world):
        t0 = time.perf_counter()
        if worker is not None:   # the pack buffer of this round was read by the job n_buffers rounds back
            while len(worker.round_tickets) > max_outstanding:
                worker.round_tickets.popleft().wait()
        counts, data, totals, segs = front_round(front, round0 + rank, n_chunks, bounds)
        t0 = lap("front", t0)
        if world == 1:
            inbox = [(counts, data, int(totals[0]), segs)]
        else:
            inbox = exchange(dist, rank, world, bounds, counts, data, totals, segs)
        ready = before_back() if before_back is not None else None
        t0 = lap("exchange", t0)
        if stop > base:
            def job(inbox=inbox, round0=round0, ready=ready):
                if ready is not None:
                    ready.synchronize()
                back_round(back, inbox, round0, n_chunks, base, stop)
            ticket = run(job)
            if ticket is not None:
                worker.round_tickets.append(ticket)
        t0 = lap("back", t0)
    t0 = time.perf_counter()

    def last():
        if stop > base:
            back.finish()
        if after_back is not None:
            after_back()
    run(last)
    lap("back", t0)


# ---- engines over the C ABI (GPU) ----------------------------------------------------------

class GpuFront(Front):
    """gm_search / gm_score / gm_candidates_pack of one context; buffers are torch CUDA tensors
    so that NCCL can send them.  `n_buffers` pack buffers rotate, so a pack may be overwritten only
    after n_buffers - 1 later rounds (the back stage / the all-to-all may still read the previous one)."""

    def __init__(self, ctx, n_queries: int, capacity: int, device, stats=None, n_buffers: int = 2):
        import torch
        self.ctx, self.stats, self.n = ctx, stats, n_queries
        self.launches = 0                # kernels launched by pack (scan + pack)
        self.bufs = [(torch.zeros(n_queries, dtype=torch.int32, device=device),
                      torch.empty(BLOCK_WORDS * capacity, dtype=torch.int32, device=device))
                     for _ in range(n_buffers)]
        self.turn = 0

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

    def _next(self):
        self.turn = (self.turn + 1) % len(self.bufs)
        return self.bufs[self.turn]

    def pack(self, bounds: np.ndarray):
        counts, data = self._next()
        totals = self.ctx.candidates_pack(bounds, counts.data_ptr(), data.data_ptr(), data.numel())
        self.launches += 2
        return counts, data[:BLOCK_WORDS * int(totals.sum())], totals

    def empty(self, bounds: np.ndarray):
        counts, data = self._next()
        counts.zero_()
        return counts, data[:0], np.zeros(len(bounds) - 1, dtype=np.uint64)


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


class GpuPipeline:
    """One rank of the sharded run over the C ABI: a front context (index chunks c % world == rank,
    all queries), a back context (residues of every chunk, the rank's query slice) and the worker
    thread that overlaps the back stage of a batch with the front stage of the next one.  The same
    object drives N = 1 (no exchange: the packed candidates go straight to the back context)."""

    def __init__(self, front_ctx, back_ctx, n_queries: int, length: int, capacity: int, device,
                 dist, rank: int, world: int, n_chunks: int, bounds: np.ndarray,
                 stats_front=None, stats_back=None, threaded: bool = True):
        self.fc, self.bc = front_ctx, back_ctx
        self.n, self.length = n_queries, length
        self.dist, self.rank, self.world, self.n_chunks, self.bounds = dist, rank, world, n_chunks, bounds
        self.base, self.stop = int(bounds[rank]), int(bounds[rank + 1])
        self.front = GpuFront(front_ctx, n_queries, capacity, device, stats_front, n_buffers=2)
        self.back = GpuBack(back_ctx, stats_back)
        self.worker = BackWorker() if threaded else None
        self.timers = {}

    def _ready_event(self):
        if self.world == 1:
            return None              # gm_candidates_pack returned: the packed block is complete
        import torch
        ev = torch.cuda.Event()
        ev.record()                  # behind the all-to-all on torch's current stream
        return ev

    def submit(self, queries_ptr: int = 0, slice_ptr: int = 0, name_break_ptr: int = 0,
               slice_break_ptr: int = 0, hits_ptr: int = 0, counts_ptr: int = 0) -> None:
        """One query batch.  With host addresses: queries H2D into both contexts and the slice's
        hit lists D2H (gm_results_download) ride in the pipeline; without: the resident queries
        are aligned again (hit lists cleared first)."""
        import time
        n_slice = self.stop - self.base
        if queries_ptr:
            t0 = time.perf_counter()
            self.fc.query_upload_ptr(queries_ptr, self.n, self.length, name_break_ptr)
            self.timers["upload"] = self.timers.get("upload", 0.0) + time.perf_counter() - t0

        def first():
            if n_slice == 0:
                return
            t0 = time.perf_counter()
            if slice_ptr:
                self.bc.query_upload_ptr(slice_ptr, n_slice, self.length, slice_break_ptr)
            else:
                self.bc.results_clear()
            self.timers["back_upload"] = self.timers.get("back_upload", 0.0) + time.perf_counter() - t0

        def after():
            if hits_ptr and n_slice:
                t0 = time.perf_counter()
                self.bc.results_download_ptr(hits_ptr, counts_ptr)
                self.timers["back_download"] = self.timers.get("back_download", 0.0) + time.perf_counter() - t0

        shard_step(self.front, self.back if n_slice else None, self.dist, self.rank, self.world,
                   self.n_chunks, self.bounds, before_back=self._ready_event, timers=self.timers,
                   worker=self.worker, after_back=after, before_first_back=first,
                   max_outstanding=len(self.front.bufs) - 1)

    def drain(self) -> None:
        if self.worker is not None:
            self.worker.drain()

    def close(self) -> None:
        if self.worker is not None:
            self.worker.close()
            self.worker = None


def hit_checksum(hits: np.ndarray, counts: np.ndarray, base: int) -> int:
    """Order-sensitive 64-bit checksum of a slice's hit lists (gm_hit[n][cap] as uint32[n][cap][9],
    counts[n]); the sum over the slices of a batch does not depend on how the batch was sliced:
    the slice-local query ids are replaced by base + row."""
    n, cap = counts.shape[0], hits.shape[1]
    h = hits.reshape(n, cap, 9).astype(np.uint64)
    k = np.arange(cap, dtype=np.uint64)[None, :]
    valid = k < counts.astype(np.uint64)[:, None]
    gq = (np.arange(n, dtype=np.uint64) + np.uint64(base))[:, None]
    mult = np.array([0, 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63,
                     0x27D4EB2F165667C5, 0xD6E8FEB86659FD93, 0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53],
                    dtype=np.uint64)          # field 0 (query_id) is slice-local: left out
    with np.errstate(over="ignore"):
        rec = ((h + np.uint64(1)) * mult[None, None, :]).sum(axis=2, dtype=np.uint64)
        pos = (gq * np.uint64(cap) + k + np.uint64(1)) * np.uint64(0x9FB21C651E98DF25)
        mixed = (rec ^ (rec >> np.uint64(29))) * pos
        total = int(mixed[valid].sum(dtype=np.uint64)) + int(counts.astype(np.uint64).sum()) * 0x2545F4914F6CDD1D
    return total & 0xFFFFFFFFFFFFFFFF
