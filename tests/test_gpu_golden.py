"""-m gpu: the CUDA path against the committed golden vectors of the unmodified reference
(stage dumps, hit lists and output text), independent of the oracle's arithmetic."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", H.GOLDEN_CASES)
def test_gpu_matches_reference_golden(gpu_ctx, name):
    db, qchunks, kw, stages, results, texts = H.golden(name)
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    si, text = 0, ""
    for qi, qc in enumerate(qchunks):
        gpu_ctx.query_upload(qc.seqs, qc.name_breaks())
        for ci in range(len(db.chunks)):
            for (_, _, ids, starts, scores, ends) in H.gpu_stage_chunks(gpu_ctx, ci, opt.max_list_length):
                r = stages[si]
                si += 1
                assert (r[0], r[1]) == (qi, ci)
                assert np.array_equal(r[3], ids) and np.array_equal(r[4], starts)
                assert np.array_equal(r[5], scores) and np.array_equal(r[6], ends)
        gpu_ctx.query_upload(qc.seqs, qc.name_breaks())     # fresh hit lists, whole path
        for ci in range(len(db.chunks)):
            gpu_ctx.align_chunk(ci)
        hits, counts = gpu_ctx.results()
        res = O.ResultLists(qc.n, opt.best)
        res.hits[:] = hits
        res.counts[:] = counts
        text += O.format_output(res, qc, db, opt)          # formatting only (statistics.cpp restated)
        if 1 in texts and len(qchunks) == 1:
            assert H.format_v1(res.lists(), qc, db) == texts[1]
    assert si == len(stages)
    assert text == texts[0]
