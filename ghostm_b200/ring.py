"""db-chunk sharding across GPUs with an exact, order-preserving merge.

The reference applies Merge once per (db chunk, candidate chunk) in ascending db order and
carries the per-query hit lists between calls (aligner.cpp:114-174); the carried hits take part
in the unstable sort, so the result depends on that order.  With the db chunks spread over the
ranks (rank r owns a contiguous block of chunks) the search and SW extension of every chunk are
independent, and only the small hit lists (<= best records per query) have to travel: rank r
receives the lists from rank r-1, merges its own chunks into them in ascending order, and sends
them on to rank r+1.  The last rank ends up with exactly the lists a single process would have.
Every rank blocks only on its predecessor, so consecutive query batches pipeline through the
ranks: in steady state each rank is busy with (search + SW + merge) of its own chunks.

The transport is torch.distributed send/recv: NCCL on the device buffers (NVLink peer-to-peer)
on GPUs, gloo on CPU tensors in the CPU tests.  The engine is abstract so that the host logic is
testable without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> List[int]:
    """Contiguous block of db chunks owned by `rank` (ascending db order across ranks)."""
    per, extra = divmod(n_chunks, world)
    start = rank * per + min(rank, extra)
    return list(range(start, start + per + (1 if rank < extra else 0)))


class Engine:
    """What the ring needs from a device (or, in tests, from the oracle)."""

    def prepare(self, chunk_id: int) -> None:      # seed search + SW extension of one db chunk
        raise NotImplementedError

    def merge(self) -> None:                        # Merge (+TraceBack) of the prepared chunk
        raise NotImplementedError

    def list_tensors(self):                         # (hits, counts) torch tensors, current lists
        raise NotImplementedError

    def lists_received(self) -> None:               # called after the tensors were overwritten
        pass


def ring_step(engine: Engine, dist, rank: int, world: int, my_chunks: Sequence[int]) -> bool:
    """One query batch through this rank.  Returns True on the rank that holds the final lists.
    `dist` is torch.distributed (or None when world == 1)."""
    first = True
    for c in my_chunks:
        engine.prepare(c)
        if first and rank > 0:
            hits, counts = engine.list_tensors()
            dist.recv(hits, src=rank - 1)
            dist.recv(counts, src=rank - 1)
            engine.lists_received()
        engine.merge()
        first = False
    if not my_chunks and rank > 0:                  # a rank without chunks just forwards
        hits, counts = engine.list_tensors()
        dist.recv(hits, src=rank - 1)
        dist.recv(counts, src=rank - 1)
        engine.lists_received()
    if rank < world - 1:
        hits, counts = engine.list_tensors()
        dist.send(hits, dst=rank + 1)
        dist.send(counts, dst=rank + 1)
        return False
    return True
