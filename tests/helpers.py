"""Shared workload builders and comparison helpers for the test-suite (test infrastructure)."""
from __future__ import annotations

import functools

import numpy as np

from ghostm_b200 import formats, synth
from oracle import oracle as O


@functools.lru_cache(maxsize=None)
def workload(name: str):
    """-> (Db, [QueryChunk], oracle Options kwargs).  Sizes: the oracle finishes in seconds."""
    if name == "small":            # several db chunks, same-name runs of 3, ragged query lengths
        dbs, dbn = synth.protein_db(11, 1_200_000)
        qs, qn = synth.queries_from_db(12, dbs, 600, 75, group=3, min_length=30)
        return formats.make_db(dbs, dbn, 4, 0.5), formats.make_query_chunks(qs, qn, 75, 0.03), {}
    if name == "frames6":          # 6 equal names per read, like translated DNA (config 2 stand-in)
        dbs, dbn = synth.protein_db(21, 1_000_000)
        qs, qn = synth.queries_from_db(20261018, dbs, 900, 25, group=6, sub_rate=0.10)
        return formats.make_db(dbs, dbn, 4, 1), formats.make_query_chunks(qs, qn, 25, 128), {}
    if name == "repeats":          # config 5 flavour: large intervals, many candidates, chunk cuts
        dbs, dbn = synth.repeat_db(5, 400_000)
        qs, qn = synth.repeat_queries(6, 80, 75)
        return (formats.make_db(dbs, dbn, 4, 1), formats.make_query_chunks(qs, qn, 75, 128),
                dict(max_list_length=3000))
    if name == "long":             # config 4 flavour: strips (L > 80 rows), cap-raised oracle
        dbs, dbn = synth.protein_db(3, 400_000)
        qs, qn = synth.queries_from_db(4, dbs, 24, 300, min_length=120)
        return formats.make_db(dbs, dbn, 4, 1), formats.make_query_chunks(qs, qn, 300, 128), {}
    if name == "options":          # non-default -r -s -e -t -b
        dbs, dbn = synth.protein_db(31, 600_000)
        qs, qn = synth.queries_from_db(32, dbs, 300, 60, group=2)
        return (formats.make_db(dbs, dbn, 4, 1), formats.make_query_chunks(qs, qn, 60, 128),
                dict(log_region=3, shift=1, extend=5, threshold=3, best=24))
    raise KeyError(name)


def setup_context(ctx, db, opt: O.Options, capacity: int = 1 << 22):
    ctx.set_options(db.seed, opt.matrix, shift=opt.shift, log_region=opt.log_region,
                    threshold=opt.threshold, extend=opt.extend, best=opt.best,
                    max_list_length=opt.max_list_length, open_gap=opt.open_gap,
                    extend_gap=opt.extend_gap)
    ctx.set_candidate_capacity(capacity)
    for i, ch in enumerate(db.chunks):
        ctx.db_upload(i, ch)


def gpu_stage_chunks(ctx, chunk_id: int, max_list_length: int):
    """Drive search -> chunk rule -> score like Aligner::Execute (aligner.cpp:131-171) and yield
    (first, end, ids, starts, scores, ends) per candidate chunk."""
    from ghostm_b200 import capi
    counts, total = ctx.search(chunk_id)
    first = 0
    while True:
        end, n, last = capi.chunk_rule(counts, first, max_list_length)
        if n == 0:
            return
        ids, starts = ctx.candidates(first, end, n)
        scores, ends = ctx.score(first, end, n)
        yield first, end, ids, starts, scores, ends
        if last:
            return
        first = end


def hits_equal(a: np.ndarray, b: np.ndarray):
    for f in ("query_id", "db_id", "db_chunk", "score", "db_start", "db_end", "aln_len", "aln_match"):
        if not np.array_equal(a[f], b[f]):
            return False, f
    if not np.array_equal(a["seq_id"].view(np.uint32), b["seq_id"].view(np.uint32)):
        return False, "seq_id"
    return True, ""
