"""TEST INFRASTRUCTURE ONLY - ctypes binding of oracle/liboracle.so (ghostm_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (ghostm_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

HIT_DTYPE = np.dtype([("query_id", "<u4"), ("db_id", "<u4"), ("db_chunk", "<u4"), ("score", "<u4"),
                      ("db_start", "<u4"), ("db_end", "<u4"), ("aln_len", "<u4"),
                      ("aln_match", "<u4"), ("seq_id", "<f4")])


def build() -> str:
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ghostm_oracle.c")
    if (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.gmo_get_key.restype = C.c_uint32
        L.gmo_get_key.argtypes = [u8p, C.c_uint32]
        L.gmo_build_index.restype = C.c_uint32
        L.gmo_build_index.argtypes = [u8p, C.c_uint32, u32p, C.c_uint32, C.c_uint32, u32p, u32p]
        L.gmo_search_query.restype = C.c_uint64
        L.gmo_search_query.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, u32p, u32p, u32p, C.c_uint64]
        L.gmo_search_begin.restype = C.c_void_p
        L.gmo_search_free.argtypes = [C.c_void_p]
        L.gmo_search_next.restype = C.c_uint64
        L.gmo_search_next.argtypes = [C.c_void_p, u8p, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p,
                                      C.POINTER(C.POINTER(C.c_uint32)),
                                      C.POINTER(C.POINTER(C.c_uint32))]
        L.gmo_calculate_score.restype = None
        L.gmo_calculate_score.argtypes = [u8p, C.c_uint32, u8p, C.c_uint32, C.c_uint64, u32p, u32p,
                                          i32p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, u32p, u32p]
        L.gmo_traceback.restype = None
        L.gmo_traceback.argtypes = [u8p, u8p, C.c_uint32, C.c_uint32, i32p, C.c_int, C.c_int,
                                    C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_float)]
        L.gmo_db_get_id.restype = C.c_uint32
        L.gmo_db_get_id.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.gmo_std_sort_hits.restype = None
        L.gmo_std_sort_hits.argtypes = [C.c_void_p, C.c_size_t]
        L.gmo_merge.restype = None
        L.gmo_merge.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_uint64, u32p, u32p, u32p, u32p,
                                u8p, C.c_uint32, C.c_uint32, u8p, u8p, C.c_uint32, u32p, C.c_uint32,
                                C.c_uint32, i32p, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                C.c_uint32]
        L.gmo_blosum62.restype = None
        L.gmo_blosum62.argtypes = [i32p]
        L.gmo_format_row.restype = C.c_int
        L.gmo_format_row.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_void_p,
                                     C.c_uint32, C.c_uint64, C.c_float, C.c_float]
        L.gmo_query_length.restype = C.c_uint32
        L.gmo_query_length.argtypes = [u8p, C.c_uint32]
        _LIB = L
    return _LIB


def blosum62() -> np.ndarray:
    m = np.zeros(32 * 32, dtype=np.int32)
    lib().gmo_blosum62(m)
    return m


class Options:
    """aligner.cpp:225-245 defaults (AlignerOption)."""

    def __init__(self, **kw):
        self.log_region = 4
        self.shift = 2
        self.threshold = 2
        self.max_list_length = 1 << 27
        self.open_gap = -11
        self.extend_gap = -1
        self.extend = 2
        self.best = 10
        self.matrix = None
        self.lam = np.float32(0.267)   # statistics.cpp:135-138 BLOSUM62 (11,1)
        self.K = np.float32(0.041)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(k)
            setattr(self, k, v)
        if self.matrix is None:
            self.matrix = blosum62()


def search_query(query: np.ndarray, chunk, opt: Options) -> np.ndarray:
    cap = 1 << 12
    while True:
        out = np.zeros(cap, dtype=np.uint32)
        n = lib().gmo_search_query(np.ascontiguousarray(query), query.shape[0], chunk.seed, opt.shift,
                                   opt.log_region, opt.threshold, chunk.keys_count, chunk.positions,
                                   out, cap)
        if n <= cap:
            return out[:n].copy()
        cap = int(n)


def search_chunks(queries: np.ndarray, chunk, opt: Options):
    """Generator over the candidate chunks of one (query chunk, db chunk) pair:
    yields (query_ids, starts) exactly as SearchNextCpu hands them to CalculateScore."""
    L = lib()
    st = L.gmo_search_begin()
    try:
        q = np.ascontiguousarray(queries, dtype=np.uint8)
        while True:
            pid = C.POINTER(C.c_uint32)()
            pst = C.POINTER(C.c_uint32)()
            n = L.gmo_search_next(st, q.reshape(-1), q.shape[0], q.shape[1], chunk.seed, opt.shift,
                                  opt.log_region, opt.threshold, opt.max_list_length,
                                  chunk.keys_count, chunk.positions, C.byref(pid), C.byref(pst))
            if n == 0:
                return
            ids = np.ctypeslib.as_array(pid, shape=(n,)).copy()
            starts = np.ctypeslib.as_array(pst, shape=(n,)).copy()
            yield ids, starts
    finally:
        L.gmo_search_free(st)


def calculate_score(queries: np.ndarray, chunk, ids: np.ndarray, starts: np.ndarray, opt: Options):
    n = ids.shape[0]
    scores = np.zeros(n, dtype=np.uint32)
    ends = np.zeros(n, dtype=np.uint32)
    q = np.ascontiguousarray(queries, dtype=np.uint8)
    lib().gmo_calculate_score(chunk.seq, chunk.seq.shape[0], q.reshape(-1), q.shape[1], n,
                              np.ascontiguousarray(ids), np.ascontiguousarray(starts), opt.matrix,
                              opt.open_gap, opt.extend_gap, opt.extend, opt.log_region, scores, ends)
    return scores, ends


def traceback(query: np.ndarray, chunk, db_end: int, opt: Options):
    a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_float()
    lib().gmo_traceback(chunk.seq, np.ascontiguousarray(query), query.shape[0], int(db_end),
                        opt.matrix, opt.open_gap, opt.extend_gap, opt.extend, opt.log_region,
                        C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return a.value, b.value, c.value, np.float32(d.value)


class ResultLists:
    """vector<vector<Alignment>> result_list of one query chunk (aligner.cpp:114)."""

    def __init__(self, n_queries: int, best: int):
        self.cap = max(best, 1)
        self.hits = np.zeros((n_queries, self.cap), dtype=HIT_DTYPE)
        self.counts = np.zeros(n_queries, dtype=np.uint32)

    def lists(self) -> List[np.ndarray]:
        return [self.hits[i, :self.counts[i]] for i in range(self.hits.shape[0])]


def merge(res: ResultLists, qchunk, chunk, db_chunk_id: int, ids, starts, scores, ends,
          opt: Options) -> None:
    q = np.ascontiguousarray(qchunk.seqs, dtype=np.uint8)
    lib().gmo_merge(res.hits.ctypes.data, res.counts, res.cap, ids.shape[0],
                    np.ascontiguousarray(ids), np.ascontiguousarray(starts),
                    np.ascontiguousarray(scores), np.ascontiguousarray(ends), q.reshape(-1),
                    q.shape[0], q.shape[1], qchunk.name_breaks(), chunk.seq, chunk.seq.shape[0],
                    chunk.seq_starts, chunk.n_seqs, db_chunk_id, opt.matrix, opt.open_gap,
                    opt.extend_gap, opt.extend, opt.log_region, opt.best)


def align_chunk(qchunk, db, opt: Options, stages=None) -> ResultLists:
    """Aligner::Execute's body for one query chunk (aligner.cpp:114-174), CPU mode.
    `stages`, if a list, receives (db_chunk, cand_chunk, ids, starts, scores, ends)."""
    res = ResultLists(qchunk.n, opt.best)
    for ci, chunk in enumerate(db.chunks):
        for cc, (ids, starts) in enumerate(search_chunks(qchunk.seqs, chunk, opt)):
            scores, ends = calculate_score(qchunk.seqs, chunk, ids, starts, opt)
            if stages is not None:
                stages.append((ci, cc, ids, starts, scores, ends))
            merge(res, qchunk, chunk, ci, ids, starts, scores, ends, opt)
    return res


def format_output(res: ResultLists, qchunk, db, opt: Options) -> str:
    """Aligner::WriteOutput (aligner.cpp:951-976), style 0."""
    L = lib()
    buf = C.create_string_buffer(4096)
    out = []
    for i in range(qchunk.n):
        qlen = L.gmo_query_length(np.ascontiguousarray(qchunk.seqs[i]), qchunk.length)
        for h in res.hits[i, :res.counts[i]]:
            rec = np.array([h], dtype=HIT_DTYPE)
            name = db.chunks[int(h["db_chunk"])].names[int(h["db_id"])]
            n = L.gmo_format_row(buf, len(buf), qchunk.names[i].encode("latin-1"),
                                 name.encode("latin-1"), rec.ctypes.data, qlen, db.sum_length_u32,
                                 C.c_float(float(opt.lam)), C.c_float(float(opt.K)))
            out.append(buf.raw[:n].decode("latin-1"))
    return "".join(out)


# ---- reference binaries (oracle/_ref), when built -------------------------------------------

def ref_bin(name: str = "ghostm"):
    p = os.path.join(_HERE, "_ref", name)
    return p if os.path.exists(p) else None


def read_probe_dump(path: str):
    """Parse oracle/ref_probe.cpp's dump -> (stages, results).
    stages: list of (qchunk, dbchunk, cchunk, ids, starts, scores, ends)
    results: {qchunk: list over queries of structured arrays (HIT_DTYPE-like + name)}"""
    data = open(path, "rb").read()
    assert data[:8] == b"GMPROBE1"
    off = 8
    stages, results = [], {}

    def u32(n=1):
        nonlocal off
        v = np.frombuffer(data, dtype="<u4", count=n, offset=off)
        off += 4 * n
        return v

    while True:
        tag = int(u32()[0])
        if tag == 0:
            break
        if tag == 1:
            qc, dc, cc, n = (int(x) for x in u32(4))
            rec = u32(4 * n).reshape(n, 4)
            stages.append((qc, dc, cc, rec[:, 0].copy(), rec[:, 1].copy(), rec[:, 2].copy(),
                           rec[:, 3].copy()))
        elif tag == 2:
            qc, nq = (int(x) for x in u32(2))
            per_query = []
            for _ in range(nq):
                nh = int(u32()[0])
                hits = []
                for _ in range(nh):
                    f = u32(8)
                    ln = int(f[7])
                    name = data[off:off + ln].decode("latin-1")
                    off += ln
                    hits.append(dict(db_id=int(f[0]), score=int(f[1]), db_start=int(f[2]),
                                     db_end=int(f[3]), aln_len=int(f[4]), aln_match=int(f[5]),
                                     seq_id=np.frombuffer(f[6:7].tobytes(), dtype="<f4")[0],
                                     db_name=name))
                per_query.append(hits)
            results[qc] = per_query
        else:
            raise ValueError(f"bad tag {tag}")
    return stages, results
