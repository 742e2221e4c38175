"""-m gpu: the BASELINE.json shapes at full chunk size.

The oracle cannot run whole configs 3-5 in a test (config 3 is ~36 core-hours), so every case
combines, on ONE full-size db chunk built on the device:
  * cross-kernel identity on all queries: the five seed-search kernels (tile, hash, bucket, sweep, generic)
    are different algorithms and must agree candidate for candidate;
  * bit-exact comparison with the oracle on a seeded SAMPLE of the queries: candidates, SW
    scores/ends and the final hit lists (queries have unique names, so a query's list depends on
    nothing but its own candidates - the sample's lists must equal the full run's rows);
  * size-independent properties: candidates ascending per query, hit scores descending per list,
    at most `best` hits, one hit per db sequence within a chunk, TraceBack start <= end.
"""
import numpy as np
import pytest

from ghostm_b200 import capi, formats, workloads
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _device_chunk(ctx, seq, starts, seed=0xF):
    ctx.db_build_index(0, seq, starts, seed)
    kc, pos = ctx.db_download_index(0, 32 ** 4 + 1, seq.shape[0])
    names = [f"s{i}" for i in range(starts.shape[0])]
    chunk = formats.DbChunk(seq, starts.astype(np.uint32), names, seed, kc, pos)
    return formats.Db(seed=seed, max_chunk_len=1 << 27,
                      sum_residues=int(seq.shape[0] - starts.shape[0]), chunks=[chunk])


def _search_all_variants(ctx, n_q, variants):
    ref = None
    for v in variants:
        ctx.set_search_variant(v)
        counts, total = ctx.search(0)
        ids, cand = ctx.candidates(0, n_q, total)
        if ref is None:
            ref = (counts, ids, cand)
        else:
            assert np.array_equal(ref[0], counts), f"variant {v}: counts differ"
            assert np.array_equal(ref[1], ids) and np.array_equal(ref[2], cand), f"variant {v}"
    ctx.set_search_variant(capi.DEFAULT_SEARCH_VARIANT)
    return ref


def _check_properties(counts, ids, cand, hits, hit_counts, best):
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    assert np.all(np.diff(ids.astype(np.int64)) >= 0)
    same_q = ids[1:] == ids[:-1]
    assert np.all(cand[1:][same_q] > cand[:-1][same_q]), "candidates must ascend inside a query"
    assert int(hit_counts.max()) <= best
    for q in range(hits.shape[0]):
        h = hits[q, :hit_counts[q]]
        if h.shape[0] == 0:
            continue
        assert np.all(np.diff(h["score"].astype(np.int64)) <= 0)
        assert np.unique(h["db_id"]).shape[0] == h.shape[0], "one hit per db sequence and chunk"
        assert np.all(h["db_start"] <= h["db_end"]) and np.all(h["aln_match"] <= h["aln_len"])
    return off


def _oracle_sample(ctx, db, queries, sample, opt, counts, ids, cand, hits, hit_counts, max_sw=4000,
                   sw_sample=None):
    """Exact comparison of the sampled queries against the oracle."""
    chunk = db.chunks[0]
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    for q in sample:
        ref = O.search_query(queries[q], chunk, opt)
        got = cand[off[q]:off[q + 1]]
        assert np.array_equal(ref, got), (q, ref.shape, got.shape)
    # SW scores/ends of (a bounded number of) the sample's candidates
    scores, ends = ctx.score(0, queries.shape[0], int(counts.sum()))
    for q in (sample if sw_sample is None else sw_sample):
        n = min(int(counts[q]), max_sw)
        if n == 0:
            continue
        qi = np.full(n, q, dtype=np.uint32)
        st = cand[off[q]:off[q] + n]
        rs, re = O.calculate_score(queries, chunk, qi, st, opt)
        assert np.array_equal(rs, scores[off[q]:off[q] + n]), q
        assert np.array_equal(re, ends[off[q]:off[q] + n]), q
    return scores, ends


def _oracle_hit_lists(db, queries, sample, opt, hits, hit_counts):
    qc = formats.QueryChunk(np.ascontiguousarray(queries[sample]), [f"q{int(i)}" for i in sample])
    ref = O.align_chunk(qc, db, opt)
    for k, q in enumerate(sample):
        assert ref.counts[k] == hit_counts[q], (q, ref.counts[k], hit_counts[q])
        a = ref.hits[k, :ref.counts[k]].copy()
        b = hits[q, :hit_counts[q]].copy()
        a["query_id"] = q
        assert a.tobytes() == b.tobytes(), q


def _context(opt, capacity):
    ctx = capi.Context(0)
    ctx.set_options(0xF, opt.matrix, shift=opt.shift, log_region=opt.log_region,
                    threshold=opt.threshold, extend=opt.extend, best=opt.best,
                    max_list_length=opt.max_list_length, open_gap=opt.open_gap,
                    extend_gap=opt.extend_gap)
    ctx.set_candidate_capacity(capacity)
    return ctx


def test_config3_full_chunk():
    """Config 3 shape: 75-aa reads against one full 120 MiB chunk of the 1 G-residue db."""
    opt = O.Options()
    n_q = 4096
    ctx = _context(opt, 1 << 24)
    seq, starts = workloads.synth_chunk(1, 0, 120 << 20)
    db = _device_chunk(ctx, seq, starts)
    queries = workloads.synth_queries(2, seq[:4 << 20].copy(), n_q, 75)
    ctx.query_upload(queries)
    counts, ids, cand = _search_all_variants(ctx, n_q, (4, 5, 6, 3, 2, 1, 0))
    assert counts.mean() > 300
    ctx.align_chunk(0)
    hits, hit_counts = ctx.results()
    _check_properties(counts, ids, cand, hits, hit_counts, opt.best)
    # candidates and hit lists of 256 sampled queries against the oracle on the FULL chunk; SW scores
    # and ends of a sub-sample (every candidate of 32 queries)
    sample = np.sort(np.random.default_rng(33).choice(n_q, size=256, replace=False))
    ctx.search(0)
    _oracle_sample(ctx, db, queries, sample, opt, counts, ids, cand, hits, hit_counts, sw_sample=sample[::8])
    _oracle_hit_lists(db, queries, sample, opt, hits, hit_counts)
    ctx.close()


def test_config4_long_queries_full_chunk():
    """Config 4 shape: long queries (strip-mined SW, generic seed search: list_len > 64) against
    one full 128 MiB chunk of the 256 M-residue db.  The oracle is the cap-raised build's
    algorithm (MAX_COLUMN_LENGTH 1024); it is compared on a bounded number of
    candidates per sampled query."""
    opt = O.Options()
    n_q, L = 24, 400
    ctx = _context(opt, 1 << 25)
    seq, starts = workloads.synth_chunk(3, 0, 128 << 20)
    db = _device_chunk(ctx, seq, starts)
    queries = workloads.synth_queries(4, seq[:4 << 20].copy(), n_q, L, frac_db=1.0)
    ctx.query_upload(queries)
    counts, ids, cand = _search_all_variants(ctx, n_q, (4, 3, 0))
    ctx.align_chunk(0)
    hits, hit_counts = ctx.results()
    _check_properties(counts, ids, cand, hits, hit_counts, opt.best)
    assert int(hit_counts.min()) > 0
    # every query is a mutated db substring: its best hit must be the source region
    sample = np.array([0, n_q // 2, n_q - 1])
    ctx.search(0)
    _oracle_sample(ctx, db, queries, sample, opt, counts, ids, cand, hits, hit_counts, max_sw=300)
    # the final hit lists of two queries (all their candidates through SW, Merge and TraceBack)
    _oracle_hit_lists(db, queries, sample[:2], opt, hits, hit_counts)
    ctx.close()


def test_config5_repeats_full_chunk():
    """Config 5 shape: low-complexity db (30 % tandem repeats) and queries carrying the same
    repeats - huge index intervals, dense tiles: exercises the bucket kernel's hand-over of
    over-capacity queries to the sweep kernel."""
    opt = O.Options()
    n_q = 512
    ctx = _context(opt, 1 << 27)
    seq, starts = workloads.synth_chunk(5, 0, 32 << 20, repeats=True)
    db = _device_chunk(ctx, seq, starts)
    from ghostm_b200 import synth
    qs, _ = synth.repeat_queries(6, n_q, 75)
    queries = np.ascontiguousarray(np.stack(qs))
    ctx.query_upload(queries)
    counts, ids, cand = _search_all_variants(ctx, n_q, (4, 5, 6, 3, 2, 1, 0))
    ctx.align_chunk(0)
    hits, hit_counts = ctx.results()
    _check_properties(counts, ids, cand, hits, hit_counts, opt.best)
    order = np.argsort(counts)
    sample = np.sort(np.array([order[0], order[n_q // 4], order[n_q // 2]]))
    ctx.search(0)
    _oracle_sample(ctx, db, queries, sample, opt, counts, ids, cand, hits, hit_counts, max_sw=2000)
    _oracle_hit_lists(db, queries, sample[:2], opt, hits, hit_counts)
    ctx.close()
