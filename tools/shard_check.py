#!/usr/bin/env python
"""Multi-GPU parity check of ghostm_b200.shard (test infrastructure; uses the oracle).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29621 tools/shard_check.py [workload ...]

Every rank owns the db chunks c % N == rank (front context, with index) and the Merge/TraceBack of
its query slice (back context, sequence-only chunks); candidates travel over NCCL all-to-all.
Each rank compares its slice with the oracle's single-process hit lists, record by record."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from ghostm_b200 import capi, shard
from oracle import oracle as O
from tests import helpers as H


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = sys.argv[1:] or ["small", "repeats", "options", "frames6"]
    for name in names:
        db, qchunks, kw = H.workload(name)
        opt = O.Options(**kw)
        n_chunks = len(db.chunks)
        front_ctx, back_ctx = capi.Context(local), capi.Context(local)
        for ctx in (front_ctx, back_ctx):
            ctx.set_options(db.seed, opt.matrix, shift=opt.shift, log_region=opt.log_region,
                            threshold=opt.threshold, extend=opt.extend, best=opt.best,
                            max_list_length=opt.max_list_length, open_gap=opt.open_gap,
                            extend_gap=opt.extend_gap)
            ctx.set_candidate_capacity(1 << 22)
        for c in shard.chunks_of_rank(n_chunks, rank, world):
            front_ctx.db_upload(c, db.chunks[c])
        for c, ch in enumerate(db.chunks):
            back_ctx.db_upload_seq(c, ch.seq, ch.seq_starts)
        n_hits = 0
        for qc in qchunks:
            bounds = shard.slice_bounds(qc.name_breaks(), qc.n, world)
            base, stop = int(bounds[rank]), int(bounds[rank + 1])
            front_ctx.query_upload(qc.seqs, qc.name_breaks())
            front = shard.GpuFront(front_ctx, qc.n, 1 << 22, f"cuda:{local}")
            back = None
            if stop > base:
                sl = H.slice_query_chunk(qc, base, stop)
                back_ctx.query_upload(sl.seqs, sl.name_breaks())
                back = shard.GpuBack(back_ctx)
            shard.shard_step(front, back, dist, rank, world, n_chunks, bounds,
                             before_back=torch.cuda.synchronize)
            if back is not None:
                ref = O.align_chunk(qc, db, opt)
                hits, counts = back_ctx.results()
                assert np.array_equal(counts, ref.counts[base:stop]), (name, rank)
                for i in range(base, stop):
                    got = hits[i - base, :counts[i - base]].copy()
                    got["query_id"] += base
                    ok, field = H.hits_equal(got, ref.hits[i, :ref.counts[i]])
                    assert ok, (name, rank, i, field)
                n_hits += int(counts.sum())
            # the same query chunk twice through the pipelined driver bench.py uses (back stage on its
            # own thread, overlapped with the next batch's front stage; host buffers in, host buffers out)
            n_slice = stop - base
            q_pin = torch.from_numpy(np.ascontiguousarray(qc.seqs)).pin_memory()
            nb_pin = torch.from_numpy(np.ascontiguousarray(qc.name_breaks().astype(np.uint8))).pin_memory()
            hits_pin = torch.zeros((max(n_slice, 1) * back_ctx.cap * 9,), dtype=torch.int32).pin_memory()
            counts_pin = torch.zeros((max(n_slice, 1),), dtype=torch.int32).pin_memory()
            pipe = shard.GpuPipeline(front_ctx, back_ctx, qc.n, qc.seqs.shape[1], 1 << 22, f"cuda:{local}", dist,
                                     rank, world, n_chunks, bounds)
            for _ in range(2):
                pipe.submit(queries_ptr=q_pin.data_ptr(), slice_ptr=q_pin[base:stop].data_ptr() if n_slice else 0,
                            name_break_ptr=nb_pin.data_ptr(),
                            slice_break_ptr=nb_pin[base:stop].data_ptr() if n_slice else 0,
                            hits_ptr=hits_pin.data_ptr(), counts_ptr=counts_pin.data_ptr())
            pipe.drain()
            pipe.close()
            if n_slice:
                got_c = counts_pin.numpy().view(np.uint32)[:n_slice]
                got_h = hits_pin.numpy().view(capi.HIT_DTYPE)[: n_slice * back_ctx.cap].reshape(n_slice, back_ctx.cap)
                assert np.array_equal(got_c, ref.counts[base:stop]), (name, rank, "pipelined")
                for i in range(base, stop):
                    got = got_h[i - base, :got_c[i - base]].copy()
                    got["query_id"] += base
                    ok, field = H.hits_equal(got, ref.hits[i, :ref.counts[i]])
                    assert ok, (name, rank, i, field, "pipelined")
        front_ctx.close()
        back_ctx.close()
        print(f"SHARD_GPU_OK {name} rank {rank}/{world} hits {n_hits}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
