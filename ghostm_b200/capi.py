"""ctypes binding of libghostm_b200.so (include/ghostm_b200.h).

This is plumbing for the test-suite and bench.py: the product is the shared library and the
C++ host driver built on it.  There is NO fallback: if the CUDA library is missing or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GHOSTM_B200_LIB") or os.path.join(_HERE, "libghostm_b200.so")   # override: A/B builds

DEFAULT_SEARCH_VARIANT = 4

LEGACY_SYMBOLS = ["InitGpu", "GetNeededGPUMemorySize", "CheckGpuMemory", "SetOptionGpu",
                  "printGpuInfo", "SetQueryGpu", "SetDbGpu", "SearchNextGpu", "CalculateScoreGpu",
                  "FreeGpu"]
EXTENDED_SYMBOLS = ["gm_version", "gm_last_error", "gm_device_count", "gm_create", "gm_destroy",
                    "gm_set_options", "gm_set_candidate_capacity", "gm_db_upload", "gm_db_release",
                    "gm_query_upload", "gm_align_chunk", "gm_results_download", "gm_results_upload",
                    "gm_search", "gm_chunk_rule", "gm_candidates_download", "gm_score", "gm_merge",
                    "gm_db_build_index", "gm_db_download_index", "gm_results_clear",
                    "gm_results_device", "gm_stream", "gm_measure_dpx_peak",
                    "gm_set_deferred_traceback", "gm_traceback_pending", "gm_set_search_variant",
                    "gm_align_prepare", "gm_align_merge", "gm_db_upload_seq", "gm_candidates_pack",
                    "gm_candidates_import", "gm_candidates_transfer", "gm_device_memory",
                    "gm_query_upload_async", "gm_align_chunk_async", "gm_results_download_async", "gm_wait"]

HIT_DTYPE = np.dtype([("query_id", "<u4"), ("db_id", "<u4"), ("db_chunk", "<u4"), ("score", "<u4"),
                      ("db_start", "<u4"), ("db_end", "<u4"), ("aln_len", "<u4"),
                      ("aln_match", "<u4"), ("seq_id", "<f4")])


class GmOptions(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("shift", C.c_uint32), ("log_region", C.c_uint32),
                ("threshold", C.c_uint32), ("extend", C.c_uint32), ("best", C.c_uint32),
                ("max_list_length", C.c_uint32), ("open_gap", C.c_int32), ("extend_gap", C.c_int32),
                ("score_matrix", C.c_int32 * 1024)]


class GmStats(C.Structure):
    _fields_ = [("candidates", C.c_uint64), ("cells", C.c_uint64), ("seed_positions", C.c_uint64),
                ("tracebacks", C.c_uint64), ("candidate_chunks", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("ms_search", C.c_float), ("ms_score", C.c_float),
                ("ms_merge", C.c_float), ("ms_traceback", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class GhostmError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GhostmError(f"{LIB_PATH} is missing: build it with `make -C ghostm_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.gm_version.restype = C.c_char_p
    L.gm_last_error.restype = C.c_char_p
    L.gm_device_count.restype = C.c_int
    L.gm_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.gm_destroy.argtypes = [vp]
    L.gm_destroy.restype = None
    L.gm_set_options.argtypes = [vp, C.POINTER(GmOptions)]
    L.gm_set_candidate_capacity.argtypes = [vp, C.c_uint64]
    L.gm_db_upload.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp,
                               C.c_uint32]
    L.gm_db_upload_seq.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.gm_candidates_pack.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint64, vp]
    L.gm_candidates_import.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint64]
    L.gm_candidates_transfer.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.gm_db_release.argtypes = [vp, C.c_uint32]
    L.gm_query_upload.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
    L.gm_align_chunk.argtypes = [vp, C.c_uint32, C.POINTER(GmStats)]
    L.gm_align_prepare.argtypes = [vp, C.c_uint32, C.POINTER(GmStats)]
    L.gm_align_merge.argtypes = [vp, C.POINTER(GmStats)]
    L.gm_results_download.argtypes = [vp, vp, vp]
    L.gm_results_download_async.argtypes = [vp, vp, vp]
    L.gm_query_upload_async.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
    L.gm_align_chunk_async.argtypes = [vp, C.c_uint32]
    L.gm_wait.argtypes = [vp, C.POINTER(GmStats)]
    L.gm_device_memory.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.gm_results_upload.argtypes = [vp, vp, vp]
    L.gm_search.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint64), C.POINTER(GmStats)]
    L.gm_chunk_rule.restype = C.c_uint32
    L.gm_chunk_rule.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64),
                                C.POINTER(C.c_int)]
    L.gm_candidates_download.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp]
    L.gm_score.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(GmStats)]
    L.gm_merge.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(GmStats)]
    L.gm_db_build_index.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.gm_db_download_index.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint32)]
    L.gm_set_deferred_traceback.argtypes = [vp, C.c_int]
    L.gm_set_search_variant.argtypes = [vp, C.c_int]
    L.gm_traceback_pending.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(GmStats)]
    L.gm_results_clear.argtypes = [vp]
    L.gm_results_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.gm_stream.restype = vp
    L.gm_stream.argtypes = [vp]
    L.gm_measure_dpx_peak.argtypes = [vp, C.POINTER(C.c_double)]
    # legacy (reference aligner_gpu.h:32-117)
    L.GetNeededGPUMemorySize.restype = C.c_size_t
    L.GetNeededGPUMemorySize.argtypes = [C.c_uint32] * 6
    L.CheckGpuMemory.argtypes = [C.c_uint32] * 6
    L.SetOptionGpu.argtypes = [C.c_uint32, vp, C.c_int]
    L.printGpuInfo.argtypes = [C.c_int]
    L.printGpuInfo.restype = None
    L.SetQueryGpu.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.SetDbGpu.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.SearchNextGpu.restype = C.c_uint32
    L.SearchNextGpu.argtypes = [C.c_uint32] * 8 + [vp, vp]
    L.CalculateScoreGpu.restype = None
    L.CalculateScoreGpu.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32,
                                    C.c_uint32, C.c_int, C.c_int]
    _lib = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def chunk_rule(counts: np.ndarray, first_query: int, max_list_length: int) -> Tuple[int, int, bool]:
    """Reference candidate-chunk rule (aligner.cpp:383-389, 511-519) -> (end, n_cand, last)."""
    L = load()
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    n = C.c_uint64()
    last = C.c_int()
    end = L.gm_chunk_rule(_ptr(counts), counts.shape[0], first_query, max_list_length,
                          C.byref(n), C.byref(last))
    return int(end), int(n.value), bool(last.value)


class Context:
    """One device context (gm_context)."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        self._check(self.L.gm_create(device, C.byref(h)))
        self.h = h
        self.n_queries = 0
        self.cap = 1
        self.opt = None

    def _check(self, rc: int):
        if rc != 0:
            raise GhostmError(f"ghostm_b200 error {rc}: {self.L.gm_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.gm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, seed: int, matrix: np.ndarray, shift=2, log_region=4, threshold=2, extend=2,
                    best=10, max_list_length=1 << 27, open_gap=-11, extend_gap=-1):
        o = GmOptions()
        o.seed, o.shift, o.log_region, o.threshold = seed, shift, log_region, threshold
        o.extend, o.best, o.max_list_length = extend, best, max_list_length
        o.open_gap, o.extend_gap = open_gap, extend_gap
        m = np.ascontiguousarray(matrix, dtype=np.int32).reshape(-1)
        assert m.shape[0] == 1024
        C.memmove(o.score_matrix, m.ctypes.data, 4096)
        self._check(self.L.gm_set_options(self.h, C.byref(o)))
        self.opt = o
        self.cap = max(best, 1)

    def set_candidate_capacity(self, n: int):
        self._check(self.L.gm_set_candidate_capacity(self.h, n))

    def db_upload(self, chunk_id: int, chunk):
        seq = np.ascontiguousarray(chunk.seq, dtype=np.uint8)
        kc = np.ascontiguousarray(chunk.keys_count, dtype=np.uint32)
        ps = np.ascontiguousarray(chunk.positions, dtype=np.uint32)
        st = np.ascontiguousarray(chunk.seq_starts, dtype=np.uint32)
        self._check(self.L.gm_db_upload(self.h, chunk_id, _ptr(seq), seq.shape[0], _ptr(kc),
                                        kc.shape[0], _ptr(ps), ps.shape[0], _ptr(st), st.shape[0]))

    def db_upload_seq(self, chunk_id: int, seq: np.ndarray, seq_starts: np.ndarray):
        """Sequence-only chunk (Merge/TraceBack side of a db-sharded run)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        st = np.ascontiguousarray(seq_starts, dtype=np.uint32)
        self._check(self.L.gm_db_upload_seq(self.h, chunk_id, _ptr(seq), seq.shape[0], _ptr(st),
                                            st.shape[0]))

    def candidates_pack(self, bounds, counts_ptr: int, data_ptr: int, capacity_words: int) -> np.ndarray:
        """gm_candidates_pack into DEVICE buffers (raw addresses) -> per-part totals (host)."""
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        totals = np.zeros(b.shape[0] - 1, dtype=np.uint64)
        self._check(self.L.gm_candidates_pack(self.h, b.shape[0] - 1, _ptr(b), counts_ptr or None,
                                              data_ptr or None, capacity_words, _ptr(totals)))
        return totals

    def candidates_import(self, chunk_id: int, counts_ptr: int, data_ptr: int, total: int):
        self._check(self.L.gm_candidates_import(self.h, chunk_id, counts_ptr, data_ptr or None, total))

    def candidates_transfer_to(self, dst: "Context", chunk_id: int, first: int, end: int):
        """gm_candidates_transfer: my scored candidates of queries [first, end) -> dst."""
        self._check(self.L.gm_candidates_transfer(self.h, dst.h, chunk_id, first, end))

    def db_build_index(self, chunk_id: int, seq: np.ndarray, seq_starts: np.ndarray, seed: int):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        st = np.ascontiguousarray(seq_starts, dtype=np.uint32)
        self._check(self.L.gm_db_build_index(self.h, chunk_id, _ptr(seq), seq.shape[0], _ptr(st),
                                             st.shape[0], seed))

    def db_download_index(self, chunk_id: int, n_keys_plus_1: int, max_positions: int):
        kc = np.zeros(n_keys_plus_1, dtype=np.uint32)
        ps = np.zeros(max_positions, dtype=np.uint32)
        n = C.c_uint32()
        self._check(self.L.gm_db_download_index(self.h, chunk_id, _ptr(kc), _ptr(ps), C.byref(n)))
        return kc, ps[:n.value].copy()

    def db_release(self, chunk_id: int):
        self._check(self.L.gm_db_release(self.h, chunk_id))

    def query_upload(self, seqs: np.ndarray, name_break: Optional[np.ndarray] = None):
        q = np.ascontiguousarray(seqs, dtype=np.uint8)
        nb = None if name_break is None else np.ascontiguousarray(name_break, dtype=np.uint8)
        self._check(self.L.gm_query_upload(self.h, _ptr(q), q.shape[0], q.shape[1], _ptr(nb)))
        self.n_queries = q.shape[0]

    def search(self, chunk_id: int, stats: Optional[GmStats] = None):
        counts = np.zeros(self.n_queries, dtype=np.uint32)
        total = C.c_uint64()
        self._check(self.L.gm_search(self.h, chunk_id, _ptr(counts), C.byref(total),
                                     C.byref(stats) if stats is not None else None))
        return counts, int(total.value)

    def candidates(self, first: int, end: int, n: int):
        ids = np.zeros(n, dtype=np.uint32)
        starts = np.zeros(n, dtype=np.uint32)
        self._check(self.L.gm_candidates_download(self.h, first, end, _ptr(ids), _ptr(starts)))
        return ids, starts

    def score(self, first: int, end: int, n: int, stats: Optional[GmStats] = None, fetch=True):
        scores = np.zeros(n, dtype=np.uint32) if fetch else None
        ends = np.zeros(n, dtype=np.uint32) if fetch else None
        self._check(self.L.gm_score(self.h, first, end, _ptr(scores), _ptr(ends),
                                    C.byref(stats) if stats is not None else None))
        return scores, ends

    def merge(self, first: int, end: int, stats: Optional[GmStats] = None):
        self._check(self.L.gm_merge(self.h, first, end, C.byref(stats) if stats is not None else None))

    def align_chunk(self, chunk_id: int, stats: Optional[GmStats] = None):
        self._check(self.L.gm_align_chunk(self.h, chunk_id,
                                          C.byref(stats) if stats is not None else None))

    def align_prepare(self, chunk_id: int, stats: Optional[GmStats] = None):
        self._check(self.L.gm_align_prepare(self.h, chunk_id,
                                            C.byref(stats) if stats is not None else None))

    def align_merge(self, stats: Optional[GmStats] = None):
        self._check(self.L.gm_align_merge(self.h, C.byref(stats) if stats is not None else None))

    def align_chunk_async(self, chunk_id: int):
        self._check(self.L.gm_align_chunk_async(self.h, chunk_id))

    def wait(self, stats: Optional[GmStats] = None):
        self._check(self.L.gm_wait(self.h, C.byref(stats) if stats is not None else None))

    def query_upload_async_ptr(self, ptr: int, n: int, length: int, name_break_ptr: int = 0):
        self._check(self.L.gm_query_upload_async(self.h, ptr, n, length, name_break_ptr or None))
        self.n_queries = n

    def results_download_async_ptr(self, hits_ptr: int, counts_ptr: int):
        self._check(self.L.gm_results_download_async(self.h, hits_ptr, counts_ptr))

    def results(self):
        hits = np.zeros((self.n_queries, self.cap), dtype=HIT_DTYPE)
        counts = np.zeros(self.n_queries, dtype=np.uint32)
        self._check(self.L.gm_results_download(self.h, _ptr(hits), _ptr(counts)))
        return hits, counts

    def set_search_variant(self, variant: int):
        """4 = tile kernel (default), 2 = bucket kernel, 3 = hash kernel, 1 = sweep kernel,
        0 = generic kernels."""
        self._check(self.L.gm_set_search_variant(self.h, int(variant)))

    def set_deferred_traceback(self, on: bool):
        self._check(self.L.gm_set_deferred_traceback(self.h, int(on)))

    def traceback_pending(self, stats: Optional[GmStats] = None) -> int:
        n = C.c_uint64()
        self._check(self.L.gm_traceback_pending(self.h, C.byref(n),
                                                C.byref(stats) if stats is not None else None))
        return int(n.value)

    def results_clear(self):
        self._check(self.L.gm_results_clear(self.h))

    def results_device(self):
        """-> (hits_ptr, counts_ptr) device addresses of the current hit lists."""
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.L.gm_results_device(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream(self) -> int:
        return self.L.gm_stream(self.h)

    def measure_dpx_peak(self) -> float:
        v = C.c_double()
        self._check(self.L.gm_measure_dpx_peak(self.h, C.byref(v)))
        return v.value

    def query_upload_ptr(self, ptr: int, n: int, length: int, name_break_ptr: int = 0):
        """gm_query_upload from a raw host address (e.g. a pinned torch tensor)."""
        self._check(self.L.gm_query_upload(self.h, ptr, n, length, name_break_ptr or None))
        self.n_queries = n

    def results_download_ptr(self, hits_ptr: int, counts_ptr: int):
        self._check(self.L.gm_results_download(self.h, hits_ptr, counts_ptr))

    def results_upload(self, hits: np.ndarray, counts: np.ndarray):
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self._check(self.L.gm_results_upload(self.h, _ptr(hits), _ptr(counts)))
