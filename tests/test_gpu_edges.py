"""-m gpu: edge cases of the path against the oracle (empty and ragged inputs, extreme options,
clipped windows at chunk ends, single-sequence db, all-X queries) and the device index build."""
import numpy as np
import pytest

from ghostm_b200 import formats, synth
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run(ctx, db, qc, opt):
    H.setup_context(ctx, db, opt)
    ctx.query_upload(qc.seqs, qc.name_breaks())
    for ci in range(len(db.chunks)):
        ctx.align_chunk(ci)
    return ctx.results()


def _check(ctx, db, qc, opt):
    ref = O.align_chunk(qc, db, opt)
    hits, counts = _run(ctx, db, qc, opt)
    assert np.array_equal(counts, ref.counts)
    for i in range(qc.n):
        assert hits[i, :counts[i]].tobytes() == ref.hits[i, :counts[i]].tobytes(), i
    return int(counts.sum())


def _small_db(seed=7, n=60_000):
    dbs, dbn = synth.protein_db(seed, n, min_len=30, max_len=200)
    return dbs, dbn


@pytest.mark.parametrize("kw", [dict(threshold=1), dict(threshold=4), dict(threshold=0), dict(best=1),
                                dict(threshold=5), dict(threshold=7, log_region=6), dict(threshold=47),
                                dict(threshold=48), dict(threshold=500),
                                dict(best=0), dict(best=40), dict(shift=1), dict(shift=7),
                                dict(log_region=1), dict(log_region=9), dict(extend=0), dict(extend=40),
                                dict(open_gap=-1, extend_gap=-3), dict(open_gap=-20, extend_gap=-5)])
def test_option_extremes(gpu_ctx, kw):
    dbs, dbn = _small_db()
    qs, qn = synth.queries_from_db(8, dbs, 64, 50, group=2, min_length=12)
    db = formats.make_db(dbs, dbn, 4, 1)
    qc = formats.make_query_chunks(qs, qn, 50, 128)[0]
    _check(gpu_ctx, db, qc, O.Options(**kw))


def test_degenerate_queries_and_db(gpu_ctx):
    dbs, dbn = _small_db(9, 20_000)
    # one query only; all-X query; query of one repeated residue; query shorter than the seed padded
    qs = [np.full(40, formats.BASE_X, dtype=np.uint8), np.zeros(40, dtype=np.uint8), dbs[0][:40].copy(),
          dbs[1][:3].copy(), np.full(40, 24, dtype=np.uint8)]
    qn = ["allx", "polyA", "exact", "tiny", "stops"]
    db = formats.make_db(dbs, dbn, 4, 1)
    qc = formats.make_query_chunks(qs, qn, 40, 128)[0]
    assert _check(gpu_ctx, db, qc, O.Options()) > 0
    one = formats.make_query_chunks(qs[2:3], qn[2:3], 40, 128)[0]
    assert _check(gpu_ctx, db, one, O.Options()) > 0
    # single-sequence db, and a db whose sequences are all shorter than the seed (empty index)
    single = formats.make_db([dbs[0]], ["only"], 4, 1)
    _check(gpu_ctx, single, qc, O.Options())
    short = formats.make_db([dbs[0][:3], dbs[1][:4], dbs[2][:2]], ["a", "b", "c"], 4, 1)
    assert short.chunks[0].positions.shape[0] == 0
    assert _check(gpu_ctx, short, qc, O.Options()) == 0


def test_windows_clipped_at_chunk_ends(gpu_ctx):
    """Hits at the very first and very last residues of a chunk: db_offset < 0 clamps to 0 and the
    window is cut at the chunk length (aligner.cpp:576-583); region 0 and its virtual rule."""
    rng = np.random.default_rng(3)
    dbs = [synth.random_residues(rng, 90) for _ in range(40)]
    dbn = [f"s{i}" for i in range(40)]
    qs = [dbs[0][:60].copy(), dbs[0][10:70].copy(), dbs[-1][-60:].copy(), dbs[-1][-70:-10].copy(),
          np.concatenate([dbs[0][17:40], dbs[0][17:54]])]
    qn = [f"q{i}" for i in range(len(qs))]
    db = formats.make_db(dbs, dbn, 4, 1)
    qc = formats.make_query_chunks(qs, qn, 60, 128)[0]
    for kw in (dict(), dict(extend=30), dict(log_region=2, threshold=1)):
        assert _check(gpu_ctx, db, qc, O.Options(**kw)) > 0


def test_spaced_seed_and_k5(gpu_ctx):
    """Seed masks other than 1111: contiguous k=5 (`db -k 5`) and a spaced mask 1101011."""
    dbs, dbn = _small_db(11, 80_000)
    qs, qn = synth.queries_from_db(12, dbs, 48, 60, sub_rate=0.05)
    qc = formats.make_query_chunks(qs, qn, 60, 128)[0]
    db5 = formats.make_db(dbs, dbn, 5, 1)
    _check(gpu_ctx, db5, qc, O.Options())
    seed = 0b1101011
    chunk = formats.make_db_chunk(dbs, dbn, seed)
    db = formats.Db(seed=seed, max_chunk_len=1 << 20, sum_residues=sum(len(s) for s in dbs), chunks=[chunk])
    assert _check(gpu_ctx, db, qc, O.Options()) > 0


def test_device_index_build_matches_reference_layout(gpu_ctx):
    """gm_db_build_index (bench path) against the counting-sort index of db_creator.cpp:167-241."""
    dbs, dbn = _small_db(13, 300_000)
    dbs[5][3:9] = formats.BASE_X
    dbs.insert(2, dbs[2][:4].copy())
    dbn.insert(2, "seedlen")
    for k in (4, 5):
        db = formats.make_db(dbs, dbn, k, 1)
        ch = db.chunks[0]
        gpu_ctx.db_build_index(3, ch.seq, ch.seq_starts, ch.seed)
        kc, pos = gpu_ctx.db_download_index(3, ch.keys_count.shape[0], ch.seq.shape[0])
        assert np.array_equal(kc, ch.keys_count)
        assert np.array_equal(pos, ch.positions)
    gpu_ctx.db_release(3)


def test_capacity_error_is_loud(gpu_ctx):
    from ghostm_b200 import capi
    db, qchunks, kw = H.workload("repeats")
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt, capacity=1000)
    gpu_ctx.query_upload(qchunks[0].seqs, qchunks[0].name_breaks())
    with pytest.raises(capi.GhostmError, match="candidate buffer"):
        gpu_ctx.align_chunk(0)


@pytest.mark.parametrize("length", [81, 100, 128, 200, 257, 520, 1000, 1023])
def test_query_lengths_across_the_traceback_kernels(gpu_ctx, length):
    """L <= 80: register TraceBack; above: the warp-cooperative wavefront kernel with 4 / 8 / 16 / 32
    rows per lane (and strip-mined SW, generic seed search once list_len > 64)."""
    dbs, dbn = _small_db(17, 80_000)
    qs, qn = synth.queries_from_db(18, dbs, 40, length, group=2, min_length=length // 2)
    db = formats.make_db(dbs, dbn, 4, 1)
    qc = formats.make_query_chunks(qs, qn, length, 128)[0]
    assert _check(gpu_ctx, db, qc, O.Options()) > 0


@pytest.mark.parametrize("kw", [dict(), dict(shift=1), dict(shift=5, log_region=2), dict(log_region=7),
                                dict(log_region=1, shift=3)])
@pytest.mark.parametrize("nw", [256, 640])
def test_tile_search_many_tiles_and_carry_zones(gpu_ctx, kw, nw, monkeypatch):
    """The tile seed-search kernel with its bitmap capped to a few hundred words (GM_TILE_NW), so a
    1 M-residue chunk spans dozens of tiles: slices, carry zones between tiles (their width depends
    on shift and region size), the split table and the scanner's range cuts all get exercised with
    non-default geometry.  Candidates must equal the oracle's, query by query."""
    monkeypatch.setenv("GM_TILE_NW", str(nw))
    dbs, dbn = synth.protein_db(91, 1_000_000)
    qs, qn = synth.queries_from_db(92, dbs, 96, 75, min_length=20)
    db = formats.make_db(dbs, dbn, 4, 1)
    qc = formats.make_query_chunks(qs, qn, 75, 128)[0]
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    gpu_ctx.query_upload(qc.seqs, qc.name_breaks())
    total_ref = 0
    for ci, chunk in enumerate(db.chunks):
        counts, total = gpu_ctx.search(ci)
        ids, cand = gpu_ctx.candidates(0, qc.n, total)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        for q in range(qc.n):
            ref = O.search_query(qc.seqs[q], chunk, opt)
            assert np.array_equal(ref, cand[off[q]:off[q + 1]]), (kw, nw, ci, q)
            total_ref += ref.shape[0]
    assert total_ref > 0
