"""-m gpu: the CUDA path (through the C ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

WORKLOADS = ["small", "frames6", "repeats", "long", "options"]


@pytest.mark.parametrize("fast_search", [True, False])
@pytest.mark.parametrize("name", WORKLOADS)
def test_stage_parity(gpu_ctx, name, fast_search):
    """(query_id, db_start) after search and (score, db_end) after SW, per candidate chunk;
    both seed-search kernels (register-window and generic)."""
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    gpu_ctx.set_search_variant(fast_search)
    n_stages = 0
    for qc in qchunks:
        gpu_ctx.query_upload(qc.seqs, qc.name_breaks())
        for ci, chunk in enumerate(db.chunks):
            ref = list(O.search_chunks(qc.seqs, chunk, opt))
            got = list(H.gpu_stage_chunks(gpu_ctx, ci, opt.max_list_length))
            assert len(ref) == len(got), (name, ci, len(ref), len(got))
            for (ids, starts), (_, _, gids, gstarts, gscores, gends) in zip(ref, got):
                assert np.array_equal(ids, gids)
                assert np.array_equal(starts, gstarts)
                scores, ends = O.calculate_score(qc.seqs, chunk, ids, starts, opt)
                assert np.array_equal(scores, gscores), np.flatnonzero(scores != gscores)[:10]
                assert np.array_equal(ends, gends), np.flatnonzero(ends != gends)[:10]
                n_stages += 1
    assert n_stages > 0
    gpu_ctx.set_search_variant(True)


@pytest.mark.parametrize("deferred,fast", [(True, True), (False, True), (False, False)])
@pytest.mark.parametrize("name", WORKLOADS)
def test_hit_lists(gpu_ctx, name, deferred, fast):
    """Final hit lists (Merge + TraceBack on the device), with TraceBack deferred to the
    survivors (default) and in the reference order (inside every Merge)."""
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    gpu_ctx.set_deferred_traceback(deferred)
    gpu_ctx.set_search_variant(fast)     # fast = register-resident search / TraceBack kernels
    for qc in qchunks:
        ref = O.align_chunk(qc, db, opt)
        gpu_ctx.query_upload(qc.seqs, qc.name_breaks())
        for ci in range(len(db.chunks)):
            gpu_ctx.align_chunk(ci)
        hits, counts = gpu_ctx.results()
        assert np.array_equal(counts, ref.counts)
        for i in range(qc.n):
            ok, field = H.hits_equal(hits[i, :counts[i]], ref.hits[i, :counts[i]])
            assert ok, (name, i, field, hits[i, :counts[i]], ref.hits[i, :counts[i]])
    assert int(ref.counts.sum()) > 0
    gpu_ctx.set_deferred_traceback(True)
    gpu_ctx.set_search_variant(True)
