"""-m gpu: the CUDA path (through the C ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

from ghostm_b200 import capi
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

WORKLOADS = ["small", "frames6", "repeats", "long", "options"]


@pytest.mark.parametrize("fast_search", [6, 5, 4, 3, 2, 1, 0])
@pytest.mark.parametrize("name", WORKLOADS)
def test_stage_parity(gpu_ctx, name, fast_search):
    """(query_id, db_start) after search and (score, db_end) after SW, per candidate chunk;
    all five seed-search kernels (tile, hash, bucket, register-window sweep, generic)."""
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
    gpu_ctx.set_search_variant(capi.DEFAULT_SEARCH_VARIANT)


@pytest.mark.parametrize("deferred,fast", [(True, 4), (True, 3), (True, 2), (False, 1), (False, 0)])
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
    gpu_ctx.set_search_variant(capi.DEFAULT_SEARCH_VARIANT)


@pytest.mark.parametrize("name,world", [("small", 3), ("repeats", 2), ("options", 2), ("frames6", 4),
                                        ("long", 2)])
def test_shard_front_back_on_one_gpu(name, world):
    """ghostm_b200.shard with `world` simulated ranks on one device: chunk-parallel front
    (search + SW + gm_candidates_pack), exchange by query slice, query-sliced back
    (gm_candidates_import + Merge + TraceBack on sequence-only chunks).  Every slice must hold
    exactly the oracle's single-process hit lists."""
    from ghostm_b200 import capi, shard
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    front_ctx = capi.Context(0)
    H.setup_context(front_ctx, db, opt)
    n_chunks = len(db.chunks)
    total_hits = 0
    for qc in qchunks:
        ref = O.align_chunk(qc, db, opt)
        bounds = shard.slice_bounds(qc.name_breaks(), qc.n, world)
        front_ctx.query_upload(qc.seqs, qc.name_breaks())
        fronts = [shard.GpuFront(front_ctx, qc.n, 1 << 22, "cuda:0") for _ in range(world)]
        backs = []
        for r in range(world):
            base, stop = int(bounds[r]), int(bounds[r + 1])
            if stop == base:
                backs.append(None)
                continue
            ctx = capi.Context(0)
            ctx.set_options(db.seed, opt.matrix, shift=opt.shift, log_region=opt.log_region,
                            threshold=opt.threshold, extend=opt.extend, best=opt.best,
                            max_list_length=opt.max_list_length, open_gap=opt.open_gap,
                            extend_gap=opt.extend_gap)
            ctx.set_candidate_capacity(1 << 22)
            for i, ch in enumerate(db.chunks):
                ctx.db_upload_seq(i, ch.seq, ch.seq_starts)
            sl = H.slice_query_chunk(qc, base, stop)
            ctx.query_upload(sl.seqs, sl.name_breaks())
            backs.append(shard.GpuBack(ctx))
        for round0 in range(0, n_chunks, world):
            outboxes = [shard.front_round(fronts[r], round0 + r, n_chunks, bounds) for r in range(world)]
            inboxes = shard.exchange_local(outboxes, bounds)
            for r in range(world):
                if backs[r] is not None:
                    shard.back_round(backs[r], inboxes[r], round0, n_chunks, int(bounds[r]),
                                     int(bounds[r + 1]))
        for r in range(world):
            if backs[r] is None:
                continue
            base, stop = int(bounds[r]), int(bounds[r + 1])
            backs[r].finish()
            hits, counts = backs[r].ctx.results()
            assert np.array_equal(counts, ref.counts[base:stop])
            for i in range(base, stop):
                got = hits[i - base, :counts[i - base]].copy()
                got["query_id"] += base
                ok, field = H.hits_equal(got, ref.hits[i, :ref.counts[i]])
                assert ok, (name, r, i, field)
            total_hits += int(counts.sum())
            backs[r].ctx.close()
    front_ctx.close()
    assert total_hits > 0


def test_sequence_only_chunk_refuses_search(gpu_ctx):
    from ghostm_b200 import capi
    db, qchunks, kw = H.workload("small")
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    gpu_ctx.db_upload_seq(0, db.chunks[0].seq, db.chunks[0].seq_starts)
    gpu_ctx.query_upload(qchunks[0].seqs, qchunks[0].name_breaks())
    with pytest.raises(capi.GhostmError, match="sequence-only"):
        gpu_ctx.search(0)
    gpu_ctx.db_upload(0, db.chunks[0])


def test_shard_over_nccl_all_visible_gpus(tmp_path):
    """The same front/back split with one process per GPU and the exchange over NCCL
    (tools/shard_check.py); needs at least two visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    n = min(n, 4)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port",
                          "29621", os.path.join(root, "tools", "shard_check.py"), "small", "repeats"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("SHARD_GPU_OK") == 2 * n


@pytest.mark.parametrize("name,world", [("small", 2), ("repeats", 3)])
def test_candidates_transfer_between_contexts(name, world):
    """gm_candidates_transfer (pack + peer copy + import in one call, the C++ driver's multi-device
    exchange): per slice the same hit lists as the oracle's single-process run."""
    from ghostm_b200 import capi, shard
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    qc = qchunks[0]
    ref = O.align_chunk(qc, db, opt)
    bounds = shard.slice_bounds(qc.name_breaks(), qc.n, world)
    front_ctx = capi.Context(0)
    H.setup_context(front_ctx, db, opt)
    front_ctx.query_upload(qc.seqs, qc.name_breaks())
    front = shard.GpuFront(front_ctx, qc.n, 1 << 22, "cuda:0")
    backs = []
    for r in range(world):
        base, stop = int(bounds[r]), int(bounds[r + 1])
        ctx = capi.Context(0)
        ctx.set_options(db.seed, opt.matrix, shift=opt.shift, log_region=opt.log_region,
                        threshold=opt.threshold, extend=opt.extend, best=opt.best,
                        max_list_length=opt.max_list_length, open_gap=opt.open_gap,
                        extend_gap=opt.extend_gap)
        ctx.set_candidate_capacity(1 << 22)
        for i, ch in enumerate(db.chunks):
            ctx.db_upload_seq(i, ch.seq, ch.seq_starts)
        sl = H.slice_query_chunk(qc, base, stop)
        ctx.query_upload(sl.seqs, sl.name_breaks())
        backs.append(ctx)
    for c in range(len(db.chunks)):
        segs = front.prepare(c)
        if not segs:
            continue
        for r, ctx in enumerate(backs):
            base, stop = int(bounds[r]), int(bounds[r + 1])
            front_ctx.candidates_transfer_to(ctx, c, base, stop)
            for f, e in segs:
                f2, e2 = min(max(f, base), stop), max(min(e, stop), base)
                if f2 >= e2:
                    f2 = e2 = base
                ctx.merge(f2 - base, e2 - base)
    for r, ctx in enumerate(backs):
        base, stop = int(bounds[r]), int(bounds[r + 1])
        hits, counts = ctx.results()
        assert np.array_equal(counts, ref.counts[base:stop])
        for i in range(base, stop):
            got = hits[i - base, :counts[i - base]].copy()
            got["query_id"] += base
            ok, field = H.hits_equal(got, ref.hits[i, :ref.counts[i]])
            assert ok, (name, r, i, field)
        ctx.close()
    front_ctx.close()


@pytest.mark.parametrize("name", WORKLOADS)
def test_async_layer_matches_the_oracle(gpu_ctx, name):
    """gm_query_upload_async / gm_align_chunk_async / gm_results_download_async / gm_wait: the whole
    batch enqueued without a host round trip.  "repeats" (candidate budget 3000: the chunk rule cuts)
    trips the device-side flag and is redone synchronously inside gm_wait; "options" (best 24, -t 3)
    is not eligible and runs synchronously inside the async call; the others stay asynchronous.
    Every variant must give the oracle's hit lists."""
    import torch
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    H.setup_context(gpu_ctx, db, opt)
    for qc in qchunks:
        ref = O.align_chunk(qc, db, opt)
        q = torch.from_numpy(np.ascontiguousarray(qc.seqs)).pin_memory()
        nb = torch.from_numpy(np.ascontiguousarray(qc.name_breaks().astype(np.uint8))).pin_memory()
        hits_t = torch.zeros((qc.n * gpu_ctx.cap * 9,), dtype=torch.int32).pin_memory()
        counts_t = torch.zeros((qc.n,), dtype=torch.int32).pin_memory()
        for rep in range(2):      # twice: the second batch starts while nothing was waited for explicitly
            gpu_ctx.query_upload_async_ptr(q.data_ptr(), qc.n, qc.seqs.shape[1], nb.data_ptr())
            for ci in range(len(db.chunks)):
                gpu_ctx.align_chunk_async(ci)
            gpu_ctx.results_download_async_ptr(hits_t.data_ptr(), counts_t.data_ptr())
        st = capi.GmStats()
        gpu_ctx.wait(st)
        counts = counts_t.numpy().view(np.uint32)
        hits = hits_t.numpy().view(capi.HIT_DTYPE).reshape(qc.n, gpu_ctx.cap)
        assert np.array_equal(counts, ref.counts), name
        for i in range(qc.n):
            ok, field = H.hits_equal(hits[i, :counts[i]], ref.hits[i, :counts[i]])
            assert ok, (name, i, field)
        if name != "options":      # the synchronous stand-in inside the async call reports no stats
            assert st.cells > 0 and st.candidates > 0
    assert int(ref.counts.sum()) > 0
