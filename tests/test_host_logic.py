"""CPU: formats, the C ABI surface, the chunk rule, the introsort restatement, the sharded driver
over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ghostm_b200 import capi, formats, shard, synth
from oracle import oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/ghostm_b200.h is the contract: every function it declares must be exported."""
    header = open(os.path.join(ROOT, "include", "ghostm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", header))
    names -= {"defined"}
    assert {"InitGpu", "SearchNextGpu", "CalculateScoreGpu", "gm_align_chunk"} <= names
    assert names == set(capi.LEGACY_SYMBOLS) | set(capi.EXTENDED_SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)      # loads without a GPU; no compute call is made here
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert b"sm_100a" in capi.load().gm_version()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under ghostm_b200/ may import or link it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ghostm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle)", ""), (dirpath, f)


@pytest.mark.parametrize("max_list", [0, 1, 7, 50, 10**9])
def test_chunk_rule_matches_search_next(max_list):
    """gm_chunk_rule (host side of the C ABI) against SearchNextCpu's carry logic restated in the
    oracle, on the candidate counts of a real search (aligner.cpp:383-389, 511-519)."""
    db, qchunks, _ = H.workload("repeats")
    qc, chunk = qchunks[0], db.chunks[0]
    opt = O.Options(max_list_length=max_list)
    counts = np.array([O.search_query(qc.seqs[i], chunk, opt).shape[0] for i in range(qc.n)],
                      dtype=np.uint32)
    ref = [(ids.min(), ids.max() + 1, ids.shape[0]) for ids, _ in O.search_chunks(qc.seqs, chunk, opt)]
    got, first = [], 0
    while True:
        end, n, last = capi.chunk_rule(counts, first, max_list)
        if n == 0:
            break
        nz = [q for q in range(first, end) if counts[q]]
        got.append((nz[0], nz[-1] + 1, n))
        if last:
            break
        first = end
    assert got == ref


def test_introsort_restatement_matches_libstdcxx(tmp_path):
    """gmo_std_sort_hits (C restatement) vs the real std::sort of this toolchain, tie-heavy input."""
    src = tmp_path / "s.cpp"
    src.write_text("""
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <vector>
struct R { uint32_t score, idx; };
int main() { uint32_t n; std::vector<R> v;
  while (fread(&n, 4, 1, stdin) == 1) { v.resize(n); if (n) fread(v.data(), 8, n, stdin);
    std::sort(v.begin(), v.end(), [](const R &a, const R &b) { return a.score > b.score; });
    for (auto &r : v) fwrite(&r.idx, 4, 1, stdout); }
  return 0; }
""")
    exe = tmp_path / "s"
    subprocess.check_call(["g++", "-O2", "-o", str(exe), str(src)])
    rng = np.random.default_rng(5)
    cases = [rng.integers(20, 20 + k, size=n).astype(np.uint32)
             for n in (0, 1, 2, 15, 16, 17, 33, 100, 700, 5000) for k in (1, 3, 40)]
    cases.append(np.arange(3000, dtype=np.uint32))            # sorted ascending: worst pivots
    cases.append(np.tile(np.array([5, 9], dtype=np.uint32), 800))
    inp = b""
    for c in cases:
        rec = np.stack([c, np.arange(c.shape[0], dtype=np.uint32)], axis=1)
        inp += np.uint32(c.shape[0]).tobytes() + rec.tobytes()
    out = np.frombuffer(subprocess.run([str(exe)], input=inp, stdout=subprocess.PIPE, check=True).stdout,
                        dtype=np.uint32)
    off = 0
    for c in cases:
        hits = np.zeros(c.shape[0], dtype=O.HIT_DTYPE)
        hits["score"] = c
        hits["query_id"] = np.arange(c.shape[0])
        O.lib().gmo_std_sort_hits(hits.ctypes.data, c.shape[0])
        assert np.array_equal(hits["query_id"], out[off:off + c.shape[0]])
        off += c.shape[0]


@pytest.mark.skipif(O.ref_bin() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_writers_byte_identical_to_reference(tmp_path):
    dbs, dbn = synth.protein_db(81, 1_200_000)
    dbs[3][5:9] = formats.BASE_X          # X inside k-mers, a sequence as short as the seed
    dbs.insert(7, dbs[7][:4].copy())
    dbn.insert(7, "tiny")
    qs, qn = synth.queries_from_db(82, dbs, 90, 75, min_length=10)
    formats.write_fasta(str(tmp_path / "db.fa"), dbn, dbs, 70)
    formats.write_fasta(str(tmp_path / "q.fa"), qn, qs)
    for cmd in (["db", "-i", str(tmp_path / "db.fa"), "-o", str(tmp_path / "rdb"), "-l", "1"],
                ["qry", "-i", str(tmp_path / "q.fa"), "-o", str(tmp_path / "rq"), "-l", "60"]):
        subprocess.check_call([O.ref_bin()] + cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    formats.write_db(str(tmp_path / "mdb"), formats.make_db(dbs, dbn, 4, 1))
    formats.write_queries(str(tmp_path / "mq"), formats.make_query_chunks(qs, qn, 60, 128))
    n = 0
    for f in sorted(os.listdir(tmp_path)):
        if f.startswith(("rdb", "rq")):
            assert (tmp_path / f).read_bytes() == (tmp_path / ("m" + f[1:])).read_bytes(), f
            n += 1
    assert n >= 12
    db = formats.read_db(str(tmp_path / "rdb"))
    for ch in db.chunks:   # the oracle's index restatement too
        kc = np.zeros_like(ch.keys_count)
        pos = np.zeros(ch.seq.shape[0], dtype=np.uint32)
        k = O.lib().gmo_build_index(ch.seq, ch.seq.shape[0], ch.seq_starts, ch.n_seqs, ch.seed, kc, pos)
        assert k == ch.positions.shape[0]
        assert np.array_equal(kc, ch.keys_count) and np.array_equal(pos[:k], ch.positions)


def test_chunks_of_rank_partition():
    for n in (1, 3, 8, 9):
        for w in (1, 2, 4, 8):
            parts = [shard.chunks_of_rank(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert all(c % w == r for r, p in enumerate(parts) for c in p)


# ---- ghostm_b200.shard: chunk-parallel front, query-sliced back ------------------------------


def test_slice_bounds_respect_runs():
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            nb = (rng.random(n) < 0.4).astype(np.uint8)
            b = shard.slice_bounds(nb, n, world)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b.astype(np.int64)) >= 0)
            for x in b[1:-1]:
                assert x == n or x == 0 or nb[x], (n, world, b)
            b2 = shard.slice_bounds(None, n, world)
            assert b2[0] == 0 and b2[-1] == n
            assert np.diff(b2.astype(np.int64)).max() - np.diff(b2.astype(np.int64)).min() <= 1


def _assert_slices_equal_single(qc, db, opt, backs, bounds):
    single = O.align_chunk(qc, db, opt)
    for r, back in enumerate(backs):
        base, stop = int(bounds[r]), int(bounds[r + 1])
        assert np.array_equal(single.counts[base:stop], back.res.counts)
        for i in range(base, stop):
            a = single.hits[i, :single.counts[i]].copy()
            b = back.res.hits[i - base, :single.counts[i]].copy()
            b["query_id"] += base
            assert a.tobytes() == b.tobytes(), (r, i)
    return int(single.counts.sum())


@pytest.mark.parametrize("name,world", [("small", 3), ("repeats", 2), ("options", 4), ("frames6", 2)])
def test_shard_local_simulation_matches_single(name, world):
    """All ranks simulated in one process with the oracle as engine: every query slice must end
    with exactly the single-process hit lists (several db chunks, same-name runs, candidate-chunk
    cuts, best > 16 where Merge re-sorts the carried lists on every call)."""
    db, qchunks, kw = H.workload(name)
    opt = O.Options(**kw)
    qc = qchunks[0]
    bounds = shard.slice_bounds(qc.name_breaks(), qc.n, world)
    fronts = [H.OracleFront(qc, db, opt) for _ in range(world)]
    backs = [H.OracleBack(H.slice_query_chunk(qc, int(bounds[r]), int(bounds[r + 1])), db, opt)
             for r in range(world)]
    n_chunks = len(db.chunks)
    for round0 in range(0, n_chunks, world):
        outboxes = [shard.front_round(fronts[r], round0 + r, n_chunks, bounds) for r in range(world)]
        inboxes = shard.exchange_local(outboxes, bounds)
        for r in range(world):
            shard.back_round(backs[r], inboxes[r], round0, n_chunks, int(bounds[r]), int(bounds[r + 1]))
    assert _assert_slices_equal_single(qc, db, opt, backs, bounds) > 0


_SHARD_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from ghostm_b200 import shard
from oracle import oracle as O
from tests import helpers as H
from tests.test_host_logic import _assert_slices_equal_single

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
db, qchunks, kw = H.workload({workload!r})
kw.update({extra!r})
opt = O.Options(**kw)
qc = qchunks[0]
bounds = shard.slice_bounds(qc.name_breaks(), qc.n, world)
front = H.OracleFront(qc, db, opt)
back = H.OracleBack(H.slice_query_chunk(qc, int(bounds[rank]), int(bounds[rank + 1])), db, opt)
worker = shard.BackWorker() if {threaded!r} else None
done = []
for batch in range(2):      # two batches: the second front overlaps the first back on the worker
    if batch:
        if worker: worker.drain()
        back.reset()
    shard.shard_step(front, back, dist, rank, world, len(db.chunks), bounds, worker=worker,
                     after_back=lambda: done.append(1))
if worker:
    worker.drain()
    worker.close()
assert len(done) == 2
single = O.align_chunk(qc, db, opt)
base, stop = int(bounds[rank]), int(bounds[rank + 1])
assert np.array_equal(single.counts[base:stop], back.res.counts)
for i in range(base, stop):
    a = single.hits[i, :single.counts[i]].copy()
    b = back.res.hits[i - base, :single.counts[i]].copy()
    b["query_id"] += base
    assert a.tobytes() == b.tobytes(), (rank, i)
print("SHARD_OK", rank, int(back.res.counts.sum()))
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.parametrize("workload,extra,threaded", [("small", {}, False), ("small", {}, True),
                                                     ("repeats", {"max_list_length": 8}, True)])
def test_shard_over_gloo_world2(tmp_path, workload, extra, threaded):
    """db chunks sharded over 2 ranks, candidates exchanged by query slice over gloo all-to-all:
    each rank must hold exactly the single-process lists of its slice - inline and with the back
    stage on its worker thread, and with more candidate chunks per db chunk (-l 8) than the first
    meta exchange carries."""
    script = tmp_path / "w.py"
    script.write_text(_SHARD_WORKER.format(root=ROOT, workload=workload, extra=extra, threaded=threaded))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29613",
                          str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.count("SHARD_OK") == 2


def test_parallel_partition_identity():
    """The identity the device Merge uses to run libstdc++'s __unguarded_partition on a whole warp
    (merge_traceback.cu, WarpLazySort): with L_k the k-th position from the left whose key is <=
    the pivot's and R_k the k-th from the right whose key is >= it, the sequential loop performs
    exactly the swaps L_k <-> R_k for k <= K = #{k: L_k < R_k} and returns L_1 if K == 0 else
    min(L_{K+1}, R_K).  Checked against the sequential loop on tie-heavy inputs."""
    rng = np.random.default_rng(0)

    def sequential(a, f, l):
        a = list(a)
        pk, first, last = a[f][0], f + 1, l
        while True:
            while a[first][0] > pk:
                first += 1
            last -= 1
            while pk > a[last][0]:
                last -= 1
            if not first < last:
                return first, a
            a[first], a[last] = a[last], a[first]
            first += 1

    def parallel(a, f, l):
        a = list(a)
        pk = a[f][0]
        L = [i for i in range(f + 1, l) if a[i][0] <= pk]
        rasc = [i for i in range(f + 1, l) if a[i][0] >= pk]
        R = rasc[::-1]
        K = 0
        while K < min(len(L), len(R)) and L[K] < R[K]:
            K += 1
        for k in range(K):
            a[L[k]], a[R[k]] = a[R[k]], a[L[k]]
        cut = L[0] if K == 0 else min(L[K] if K < len(L) else 1 << 60, R[K - 1])
        return cut, a

    for _ in range(3000):
        n, spread = int(rng.integers(17, 300)), int(rng.integers(1, 9))
        a = [(int(k), i) for i, k in enumerate(rng.integers(0, spread, size=n))]
        f, l = 0, n
        x, y, z = f + 1, f + (l - f) // 2, l - 1          # __move_median_to_first
        gt = lambda i, j: a[i][0] > a[j][0]
        if gt(x, y):
            m = y if gt(y, z) else (z if gt(x, z) else x)
        else:
            m = x if gt(x, z) else (z if gt(y, z) else y)
        a[f], a[m] = a[m], a[f]
        assert sequential(a, f, l) == parallel(a, f, l)


def test_hit_checksum_is_slice_invariant_and_order_sensitive():
    """bench.py's per-N proof: the checksum of a batch's hit lists summed over query slices does not
    depend on where the slices are cut, and changes when two hits of a list swap places."""
    rng = np.random.default_rng(5)
    n, cap = 97, 10
    hits = rng.integers(0, 2 ** 32, size=(n, cap, 9), dtype=np.uint64).astype(np.uint32)
    counts = rng.integers(0, cap + 1, size=n).astype(np.uint32)
    full = shard.hit_checksum(hits, counts, 0)
    for cuts in ([40], [1, 2, 96], [13, 50, 51, 80]):
        b = [0] + cuts + [n]
        parts = sum(shard.hit_checksum(hits[b[i]:b[i + 1]], counts[b[i]:b[i + 1]], b[i]) for i in range(len(b) - 1))
        assert parts & (2 ** 64 - 1) == full
    q = int(np.flatnonzero(counts >= 2)[0])
    swapped = hits.copy()
    swapped[q, 0], swapped[q, 1] = hits[q, 1].copy(), hits[q, 0].copy()
    assert shard.hit_checksum(swapped, counts, 0) != full
    beyond = hits.copy()
    beyond[q, counts[q]:] += 1          # records past the count are not part of the list
    assert shard.hit_checksum(beyond, counts, 0) == full


def test_back_worker_runs_in_order_and_surfaces_errors():
    w = shard.BackWorker()
    seen = []
    tickets = [w.submit(lambda i=i: seen.append(i)) for i in range(50)]
    w.drain()
    assert seen == list(range(50)) and all(t.is_set() for t in tickets)

    def boom():
        raise ValueError("back stage failed")
    t = w.submit(boom)
    later = w.submit(lambda: seen.append("skipped"))
    t.wait(); later.wait()
    with pytest.raises(ValueError):
        w.drain()
    assert "skipped" not in seen        # jobs behind a failure do not run
    w.close()
