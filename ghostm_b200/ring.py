"""db-chunk sharding across GPUs with an exact, order-preserving merge.

The reference applies Merge once per (db chunk, candidate chunk) in ascending db order and
carries the per-query hit lists between calls (aligner.cpp:114-174); the carried hits take part
in the unstable sort, so the result depends on that order.  With the db chunks spread over the
ranks (chunk c lives on rank c % N) the search and SW extension of every chunk are independent,
and only the small hit lists (<= best records per query) have to travel: the owner of chunk c
receives the lists from the owner of chunk c-1, merges chunk c into them and sends them on to
the owner of chunk c+1.  The owner of the last chunk ends up with exactly the lists a single
process would have.  Every rank blocks only on its predecessor and does its search + extension
before it needs the lists, so the ranks work in parallel and consecutive query batches pipeline
through them: in steady state each rank is busy with (search + SW + merge) of its own chunks.

The transport is torch.distributed send/recv: NCCL on the device buffers (NVLink peer-to-peer)
on GPUs, gloo on CPU tensors in the CPU tests.  The engine is abstract so that the host logic is
testable without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> List[int]:
    """db chunks owned by `rank`: round robin, chunk c lives on rank c % world.  Consecutive
    chunks sit on consecutive ranks, so the hit lists hop one rank per chunk and no rank waits
    for a whole block of foreign chunks before it can merge."""
    return list(range(rank, n_chunks, world))


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


def ring_step(engine: Engine, dist, rank: int, world: int, n_chunks: int) -> bool:
    """One query batch through this rank: for every owned chunk c (ascending) search + extend,
    receive the hit lists from the owner of chunk c-1, merge, pass them to the owner of chunk
    c+1.  Returns True on the rank that ends up with the final lists (owner of the last chunk).
    `dist` is torch.distributed (unused when world == 1)."""
    for c in chunks_of_rank(n_chunks, rank, world):
        engine.prepare(c)
        if c > 0 and world > 1:
            hits, counts = engine.list_tensors()
            dist.recv(hits, src=(c - 1) % world)
            dist.recv(counts, src=(c - 1) % world)
            engine.lists_received()
        engine.merge()
        if c < n_chunks - 1 and world > 1:
            hits, counts = engine.list_tensors()
            dist.send(hits, dst=(c + 1) % world)
            dist.send(counts, dst=(c + 1) % world)
    return rank == (n_chunks - 1) % world
