// Merge (exact top-`best` selection) and TraceBack on the device, sm_100a.
//
// Merge semantics: reference Aligner::Merge, aligner.cpp:687-769.  For every run of equally
// named queries the list l = [new candidates of the run's queries in (query, region) order] ++
// [hits carried from earlier calls] is sorted with std::sort and the comparator "score
// descending" (aligner.cpp:52-63), then walked: carried hits pass, new hits pass once per db
// sequence (DB::GetID of the forward end, db.h:94-120), the walk stops at `best` hits; the list
// lives at the last query of the run.  std::sort is unstable, so which of several equal-score
// candidates survives depends on the exact algorithm: libstdc++'s introsort (bits/stl_algo.h,
// threshold 16, median-of-3 to first, unguarded Hoare partition, heapsort at depth 2*lg n,
// final insertion sort) is REPLAYED here move for move on (score, list index) records.  One warp
// owns one run; lane 0 executes the replay in shared memory, the warp builds the list and
// copies records cooperatively.  Selection never depends on TraceBack's output, so accepted new
// hits are queued and traced by a second kernel.
//
// TraceBack semantics: reference Aligner::TraceBack, aligner.cpp:771-949: reverse affine SW from
// the forward end over at most L + 2*extend*2*2^r columns, stop at the first SEQUENCE_END,
// track match count and alignment length along the argmax path (diagonal first, then strictly
// greater insertion, then strictly greater deletion), keep the FIRST strict maximum.
#include <stdlib.h>

#include "gm_common.cuh"

namespace gm {

namespace {

constexpr int kMergeThreads = 512;
constexpr int kMergeWarps = kMergeThreads / 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kCarriedFlag = 0x8000u;  // payload bit: record refers to a carried hit

// ---- libstdc++ 13 std::sort, replayed.  Rec is u32 (score<<16 | ref) or u64 (score<<32 | ref);
// only the score takes part in comparisons, exactly like AlignmentComp.
template <typename Rec> struct RecKey;
template <> struct RecKey<uint32_t> {
  static __device__ __forceinline__ uint32_t key(uint32_t r) { return r >> 16; }
};
template <> struct RecKey<unsigned long long> {
  static __device__ __forceinline__ uint32_t key(unsigned long long r) { return (uint32_t)(r >> 32); }
};

template <typename Rec>
__device__ __forceinline__ bool comp(Rec a, Rec b) {  // aligner.cpp:52-63: a.score > b.score
  return RecKey<Rec>::key(a) > RecKey<Rec>::key(b);
}

template <typename Rec>
__device__ void push_heap_(Rec *first, long hole, long top, Rec value) {
  long parent = (hole - 1) / 2;
  while (hole > top && comp(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

template <typename Rec>
__device__ void adjust_heap_(Rec *first, long hole, long len, Rec value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (comp(first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  push_heap_(first, hole, top, value);
}

template <typename Rec>
__device__ void heap_sort_(Rec *first, Rec *last) {  // __partial_sort(first, last, last)
  const long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    while (true) {
      const Rec value = first[parent];
      adjust_heap_(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    const Rec value = *last;
    *last = *first;
    adjust_heap_(first, 0L, (long)(last - first), value);
  }
}

template <typename Rec>
__device__ __forceinline__ void swap_(Rec *a, Rec *b) {
  const Rec t = *a;
  *a = *b;
  *b = t;
}

template <typename Rec>
__device__ void unguarded_linear_insert_(Rec *last) {
  const Rec val = *last;
  Rec *next = last - 1;
  while (comp(val, *next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

template <typename Rec>
__device__ void insertion_sort_(Rec *first, Rec *last) {
  if (first == last) return;
  for (Rec *i = first + 1; i != last; ++i) {
    if (comp(*i, *first)) {
      const Rec val = *i;
      for (Rec *p = i; p != first; --p) *p = *(p - 1);  // move_backward(first, i, i + 1)
      *first = val;
    } else {
      unguarded_linear_insert_(i);
    }
  }
}

// std::sort(first, first + n, comp), replayed LAZILY and WARP-COOPERATIVELY.  Only a prefix of
// the sorted array is ever read by the Merge walk (it stops after `best` accepted hits), and the
// prefix can be produced without touching most of the array, move for move identical to the
// full sort:
//  * __introsort_loop recurses into [cut, last) and loops on [first, cut); the two halves are
//    disjoint and never exchange elements afterwards, so the right halves can wait on a stack
//    until the walk actually needs positions inside them;
//  * after the loop phase the array is a sequence of blocks of <= 16 elements (or heap-sorted
//    ranges), each block >= the next one under the comparator; __final_insertion_sort inserts
//    elements left to right and an element never crosses into an earlier block (the comparison
//    is strict), so positions [0, e) are final as soon as every element before the block
//    boundary e has been inserted;
//  * __unguarded_partition(f+1, l, pivot = *f) is a fixed function of the array, not of the
//    order in which a machine evaluates it: with L_1 < L_2 < ... the positions whose key is
//    <= the pivot's ("left stoppers") and R_1 > R_2 > ... those whose key is >= it ("right
//    stoppers"), the scan pointers only ever rest on untouched positions or on the partner of
//    the previous swap, hence the sequential loop performs exactly the swaps L_k <-> R_k for
//    k = 1..K, K = #{k : L_k < R_k} (monotone), and returns L_1 if K = 0, else
//    min(L_{K+1}, R_K).  The warp compacts both stopper lists with ballots, finds K and does
//    the K swaps in parallel (checked against the sequential loop in
//    tests/test_host_logic.py::test_parallel_partition_identity).
// All lanes run the same control flow on replicated scalars; lane 0 alone executes the short
// sequential pieces (median of three, insertion of <= 16-element blocks, heapsort fallback).
template <typename Rec, typename Pos>
struct WarpLazySort {
  struct Frame { int first, last, depth; };
  Rec *a;
  Pos *lpos, *rasc;   // scratch: left stoppers ascending, right stoppers ascending (n entries each)
  int n;
  int looped;     // [0, looped) went through the introsort loop phase (block boundary)
  int final_end;  // [0, final_end) is in its final sorted place
  int sp;
  Frame stack[64];

  __device__ void init(Rec *arr, Pos *scratch, int count) {
    a = arr;
    lpos = scratch;
    rasc = scratch + count;
    n = count;
    looped = 0;
    final_end = 0;
    sp = 0;
    int lg = 0;
    for (int m = count; m > 1; m >>= 1) ++lg;
    if (count > 0) stack[sp++] = Frame{0, count, lg * 2};
  }

  __device__ void loop_next_block(uint32_t lane) {  // __introsort_loop on the leftmost pending range
    if (sp == 0) { looped = n; return; }   // cannot happen: the frames always cover [looped, n)
    const Frame fr = stack[--sp];
    int f = fr.first, l = fr.last, depth = fr.depth;
    while (l - f > 16) {
      if (depth == 0) {
        if (lane == 0) heap_sort_(a + f, a + l);
        __syncwarp();
        break;
      }
      --depth;
      if (lane == 0) {
        Rec *pf = a + f, *x = pf + 1, *y = pf + (l - f) / 2, *z = a + l - 1;  // __move_median_to_first
        if (comp(*x, *y)) {
          if (comp(*y, *z)) swap_(pf, y);
          else if (comp(*x, *z)) swap_(pf, z);
          else swap_(pf, x);
        } else if (comp(*x, *z)) swap_(pf, x);
        else if (comp(*y, *z)) swap_(pf, z);
        else swap_(pf, y);
      }
      __syncwarp();
      const uint32_t pk = RecKey<Rec>::key(a[f]);
      const uint32_t lt = (1u << lane) - 1u;
      uint32_t nA = 0, nB = 0;
      for (int base = f + 1; base < l; base += 32) {     // __unguarded_partition(f + 1, l, f)
        const int i = base + (int)lane;
        const bool valid = i < l;
        const uint32_t k = valid ? RecKey<Rec>::key(a[i]) : 0u;
        const bool isA = valid && k <= pk, isB = valid && k >= pk;
        const uint32_t ba = __ballot_sync(kFull, isA), bb = __ballot_sync(kFull, isB);
        if (isA) lpos[nA + __popc(ba & lt)] = (Pos)i;
        if (isB) rasc[nB + __popc(bb & lt)] = (Pos)i;
        nA += __popc(ba);
        nB += __popc(bb);
      }
      __syncwarp();
      const uint32_t mn = nA < nB ? nA : nB;
      uint32_t K = 0;
      for (uint32_t k0 = 0; k0 < mn; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool ok = k < mn && (uint32_t)lpos[k] < (uint32_t)rasc[nB - 1 - k];
        const uint32_t b = __ballot_sync(kFull, ok);
        K += __popc(b);
        if (b != kFull) break;
      }
      for (uint32_t k = lane; k < K; k += 32) swap_(a + lpos[k], a + rasc[nB - 1 - k]);
      int cut;
      if (K == 0) {
        cut = (int)lpos[0];
      } else {
        const int c2 = (int)rasc[nB - K];
        cut = (K < nA && (int)lpos[K] < c2) ? (int)lpos[K] : c2;
      }
      __syncwarp();
      stack[sp++] = Frame{cut, l, depth};
      l = cut;
    }
    looped = l;
  }

  // make position i final; returns false if i >= n
  __device__ bool ensure(int i, uint32_t lane) {
    if (i >= n) return false;
    while (final_end <= i) {
      if (looped == final_end) loop_next_block(lane);
      if (lane == 0) {
        for (int k = final_end; k < looped; ++k) {   // __final_insertion_sort, element k
          if (k == 0) continue;
          Rec *e = a + k;
          if ((n <= 16 || k < 16) && comp(*e, *a)) {  // __insertion_sort's guarded branch
            const Rec val = *e;
            for (Rec *q = e; q != a; --q) *q = *(q - 1);
            *a = val;
          } else {
            unguarded_linear_insert_(e);
          }
        }
      }
      __syncwarp();
      final_end = looped;
    }
    return true;
  }
};

// DB::GetID, db.h:94-120: the id with pos[id] <= position < pos[id + 1] (the last sequence ends
// at seq_len); the reference's binary search returns UINT_MAX when there is none.  Same answer
// from a warp-wide 32-ary search (4 rounds instead of 19 dependent loads for 350 k sequences).
__device__ uint32_t db_get_id(const uint32_t *pos, uint32_t n_seqs, uint32_t seq_len,
                              uint32_t position, uint32_t lane) {
  if (position >= seq_len || pos[0] > position) return kNoId;
  if (pos[n_seqs - 1] <= position) return n_seqs - 1;
  uint32_t lo = 0, hi = n_seqs - 1;       // answer in [lo, hi), pos[lo] <= position < pos[hi]
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + 31) / 32;
    const uint32_t idx = lo + lane * step;
    const bool ok = idx < hi && pos[idx] <= position;
    const uint32_t b = __ballot_sync(kFull, ok);
    if (b == 0) return kNoId;              // cannot happen: pos[lo] <= position is invariant
    const uint32_t nlo = lo + (31 - __clz(b)) * step;
    hi = hi < nlo + step ? hi : nlo + step;
    lo = nlo;
  }
  return lo;
}

template <typename Rec, typename Pos>
__device__ void merge_run(const MergeParams &p, Rec *list, Pos *scratch, uint32_t n_new,
                          uint32_t n_old, uint32_t qf, uint32_t ql, uint32_t lane) {
  constexpr bool kWide = sizeof(Rec) == 8;
  const uint32_t n = n_new + n_old;
  // ---- build the list in reference order (aligner.cpp:732-740)
  uint32_t fill = 0;
  for (uint32_t q = qf; q <= ql; ++q) {
    if (q < p.first_query || q >= p.end_query) continue;
    const uint32_t cnt = p.cand_cnt[q], off = p.cand_off[q];
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint32_t score = p.cand_score[off + i];
      if (kWide) list[fill + i] = (Rec)(((unsigned long long)score << 32) | (fill + i));
      else list[fill + i] = (Rec)((score << 16) | (fill + i));
    }
    fill += cnt;
  }
  for (uint32_t i = lane; i < n_old; i += 32) {
    const uint32_t score = p.old_hits[(size_t)ql * p.cap + i].score;
    if (kWide) list[n_new + i] = (Rec)(((unsigned long long)score << 32) | 0x80000000u | i);
    else list[n_new + i] = (Rec)((score << 16) | kCarriedFlag | i);
  }
  __syncwarp();
  WarpLazySort<Rec, Pos> sorter;
  sorter.init(list, scratch, (int)n);
  // ---- walk (aligner.cpp:703-725 / :746-768) over the lazily sorted list, warp-uniform
  uint32_t out = 0;
  gm_hit *dst = p.new_hits + (size_t)ql * p.cap;
  const gm_hit *old = p.old_hits + (size_t)ql * p.cap;
  for (uint32_t it = 0; it < n; ++it) {
    sorter.ensure((int)it, lane);
    const Rec r = list[it];
    const uint32_t ref = kWide ? (uint32_t)r : ((uint32_t)r & 0xFFFFu);
    const bool carried = kWide ? (ref & 0x80000000u) != 0 : (ref & kCarriedFlag) != 0;
    if (carried) {
      const uint32_t idx = kWide ? (ref & 0x7FFFFFFFu) : (ref & 0x7FFFu);
      if (out < p.cap && lane < sizeof(gm_hit) / 4)
        reinterpret_cast<uint32_t *>(dst + out)[lane] = reinterpret_cast<const uint32_t *>(old + idx)[lane];
      ++out;
    } else {
      // locate (query, i) of list position ref
      uint32_t q = qf, rem = ref;
      for (;; ++q) {
        if (q < p.first_query || q >= p.end_query) continue;
        const uint32_t cnt = p.cand_cnt[q];
        if (rem < cnt) break;
        rem -= cnt;
      }
      const uint32_t g = p.cand_off[q] + rem;
      const uint32_t end = p.cand_end[g];
      const uint32_t db_id = db_get_id(p.seq_starts, p.n_seqs, p.db_len, end, lane);
      // overlap[db_id] == id (aligner.cpp:707): accepted earlier in this call
      const uint32_t have = out < p.cap ? out : p.cap;
      bool mine = false;
      for (uint32_t k = lane; k < have; k += 32)
        mine |= dst[k].db_chunk == p.db_chunk && dst[k].db_id == db_id &&
                dst[k].aln_match == kNoId && dst[k].aln_len == p.serial;
      if (!__any_sync(kFull, mine)) {
        if (out < p.cap && lane == 0) {
          gm_hit h;
          h.query_id = q;
          h.db_id = db_id;
          h.db_chunk = p.db_chunk;
          h.score = p.cand_score[g];
          h.db_start = p.cand_start ? p.cand_start[g] : 0u;   // placeholder until TraceBack (:941)
          h.db_end = end;          // absolute; TraceBack makes both sequence-relative
          h.aln_len = p.serial;    // accepted by this call ...
          h.aln_match = kNoId;     // ... TraceBack pending
          h.seq_id = 0.f;
          dst[out] = h;
          if (!p.deferred) p.jobs[atomicAdd(p.n_jobs, 1u)] = ql * p.cap + out;
        }
        ++out;
      }
    }
    __syncwarp();
    if (out >= p.best) break;  // aligner.cpp:722-724
  }
  if (lane == 0) p.new_cnt[ql] = out < p.cap ? out : p.cap;
  __syncwarp();
}

__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint32_t lists[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: smem_elems u32 records, then 2 x smem_elems u16 stopper positions
  uint32_t *my_list = lists + (size_t)warp * p.smem_elems * 2;
  uint16_t *my_scratch = reinterpret_cast<uint16_t *>(my_list + p.smem_elems);
  while (true) {
    uint32_t run = 0;
    if (lane == 0) run = atomicAdd(p.run_counter, 1u);
    run = __shfl_sync(kFull, run, 0);
    if (run >= p.n_runs) break;
    const uint32_t qf = p.run_first[run], ql = p.run_last[run];
    uint32_t n_new = 0;
    for (uint32_t q = qf + lane; q <= ql; q += 32) {
      if (q >= p.first_query && q < p.end_query) n_new += p.cand_cnt[q];
      if (q != ql) p.new_cnt[q] = 0;               // aligner.cpp:741
    }
    n_new = __reduce_add_sync(kFull, n_new);
    const uint32_t n_old = p.old_cnt[ql];
    const uint32_t n = n_new + n_old;
    if (n_new == 0 && n_old <= 16) {
      // sorted input of <= 16 records goes through insertion sort only, which is stable:
      // the carried list is unchanged (aligner.cpp:702 on an already sorted list).
      for (uint32_t i = lane; i < n_old; i += 32)
        p.new_hits[(size_t)ql * p.cap + i] = p.old_hits[(size_t)ql * p.cap + i];
      if (lane == 0) p.new_cnt[ql] = n_old;
      continue;
    }
    if (n <= p.smem_elems && n < kCarriedFlag && !p.wide_scores) {
      merge_run<uint32_t, uint16_t>(p, my_list, my_scratch, n_new, n_old, qf, ql, lane);
    } else {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(p.big_cursor, 2ull * n);   // n records + 2 n u32 positions
      base = __shfl_sync(kFull, base, 0);
      if (base + 2ull * n > p.big_capacity) {
        if (lane == 0) { atomicExch(p.error, 1); p.new_cnt[ql] = 0; }
        continue;
      }
      merge_run<unsigned long long, uint32_t>(p, p.big_scratch + base,
                                              reinterpret_cast<uint32_t *>(p.big_scratch + base + n),
                                              n_new, n_old, qf, ql, lane);
    }
  }
}

// One thread per accepted hit; column state in global scratch ([4][L+1] ints per thread,
// interleaved by thread so that neighbouring threads touch neighbouring words).
__global__ void __launch_bounds__(128) traceback_kernel(const TracebackParams p) {
  const uint32_t n_jobs = *p.n_jobs;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = (int)p.query_len;
  int *dp = p.work + tid;                        // element k of array a: dp[(a*(L+1) + k) * stride]
#define GM_COL(a, k) dp[((size_t)(a) * (L + 1) + (k)) * stride]
  for (uint32_t job = tid; job < n_jobs; job += stride) {
    gm_hit h = p.hits[p.jobs[job]];
    const ChunkRef chunk = p.chunks[h.db_chunk];
    const uint8_t *query = p.queries + (size_t)h.query_id * L;
    const uint32_t db_offset = h.db_end;
    uint32_t len = p.base_len;
    if (db_offset < len) len = db_offset + 1;                         // aligner.cpp:802-805
    for (int k = 0; k <= L; ++k) { GM_COL(0, k) = 0; GM_COL(1, k) = 0; GM_COL(2, k) = 0; GM_COL(3, k) = 0; }
    int max_score = 0;
    uint32_t max_start = 0, max_match = 0, max_len = 0;
    for (uint32_t j = 0; j < len; ++j) {
      const uint8_t c = chunk.seq[db_offset - j];
      if (c == kSeqEnd) break;                                        // aligner.cpp:927-929
      const int32_t *row = p.matrix + c * kAlphabet;
      int temp_score = 0, del = 0;
      uint32_t temp_match = 0, temp_len = 0;
      // the row below (k+1) of this column, carried in registers
      int below_h = GM_COL(0, L);
      uint32_t below_match = (uint32_t)GM_COL(2, L), below_len = (uint32_t)GM_COL(3, L);
      for (int k = L - 1; k >= 0; --k) {
        const int old_h = GM_COL(0, k);
        int ins = GM_COL(1, k);
        const uint32_t old_match = (uint32_t)GM_COL(2, k), old_len = (uint32_t)GM_COL(3, k);
        int local = 0;
        uint32_t nm = 0, nl = 0;
        const int s = temp_score + row[query[k]];
        if (s > 0) {
          local = s;
          nm = temp_match + (c == query[k] ? 1u : 0u);
          nl = temp_len + 1;
        }
        ins = (ins + p.extend_gap < old_h + p.open_gap) ? old_h + p.open_gap : ins + p.extend_gap;
        if (ins > local) { local = ins; nm = old_match; nl = old_len + 1; }
        del = (del + p.extend_gap < below_h + p.open_gap) ? below_h + p.open_gap : del + p.extend_gap;
        if (del > local) { local = del; nm = below_match; nl = below_len + 1; }
        temp_score = old_h;
        temp_match = old_match;
        temp_len = old_len;
        GM_COL(0, k) = local;
        GM_COL(1, k) = ins;
        GM_COL(2, k) = (int)nm;
        GM_COL(3, k) = (int)nl;
        below_h = local;
        below_match = nm;
        below_len = nl;
        if (local > max_score) {                                      // ">": first maximum wins
          max_score = local;
          max_start = j;
          max_match = nm;
          max_len = nl;
        }
      }
    }
    const uint32_t seq_pos = chunk.seq_starts[h.db_id];
    h.db_start = db_offset - max_start - seq_pos;                     // aligner.cpp:941, :715
    h.db_end = db_offset - seq_pos;                                   // :716
    h.seq_id = (float)max_match / (float)(int)max_len;                // :945
    h.aln_len = max_len;
    h.aln_match = max_match;
    p.hits[p.jobs[job]] = h;
  }
#undef GM_COL
}

// Register-resident TraceBack for L <= R rows: one thread per hit, the column state of row k is
// two packed words, (H << 16 | E) and (matches << 16 | alignment length), so the reverse DP of a
// hit never leaves the register file; the score matrix (16-bit) and the thread's own query
// residues (transposed, one byte column per thread) sit in shared memory.
template <int R>
__global__ void __launch_bounds__(128, 2) traceback_reg_kernel(const TracebackParams p) {
  __shared__ int16_t mat[kAlphabet * kAlphabet];
  __shared__ uint8_t qs[R][128];
  for (int i = threadIdx.x; i < kAlphabet * kAlphabet; i += blockDim.x) mat[i] = (int16_t)p.matrix[i];
  __syncthreads();
  const uint32_t n_jobs = *p.n_jobs;
  const int L = (int)p.query_len;
  const int go = p.open_gap, ge = p.extend_gap;
  for (uint32_t job = blockIdx.x * blockDim.x + threadIdx.x; job < n_jobs;
       job += gridDim.x * blockDim.x) {
    gm_hit h = p.hits[p.jobs[job]];
    const ChunkRef chunk = p.chunks[h.db_chunk];
    const uint8_t *query = p.queries + (size_t)h.query_id * L;
#pragma unroll 5
    for (int k = 0; k < R; ++k) qs[k][threadIdx.x] = k < L ? query[k] : (uint8_t)kBaseX;
    const uint32_t db_offset = h.db_end;
    uint32_t len = p.base_len;
    if (db_offset < len) len = db_offset + 1;                         // aligner.cpp:802-805
    uint32_t he[R], pay[R];
#pragma unroll
    for (int k = 0; k < R; ++k) { he[k] = 0; pay[k] = 0; }
    int max_score = 0;
    uint32_t max_start = 0, max_pay = 0;
    for (uint32_t j = 0; j < len; ++j) {
      const uint32_t c = chunk.seq[db_offset - j];
      if (c == kSeqEnd) break;                                        // aligner.cpp:927-929
      const int16_t *row = mat + c * kAlphabet;
      int temp_score = 0, del = 0, below_h = 0;
      uint32_t temp_pay = 0, below_pay = 0;
#pragma unroll
      for (int k = R - 1; k >= 0; --k) {
        if (k < L) {
          const uint32_t old = he[k], old_pay = pay[k];
          const int old_h = (int)old >> 16;
          int ins = (int)(int16_t)(old & 0xFFFFu);
          const uint32_t qk = qs[k][threadIdx.x];
          const int s = temp_score + row[qk];
          int local = 0;
          uint32_t np = 0;
          if (s > 0) { local = s; np = temp_pay + (c == qk ? 0x10001u : 0x1u); }
          ins = max(ins + ge, old_h + go);
          if (ins > local) { local = ins; np = old_pay + 1; }
          del = max(del + ge, below_h + go);
          if (del > local) { local = del; np = below_pay + 1; }
          temp_score = old_h;
          temp_pay = old_pay;
          he[k] = ((uint32_t)local << 16) | ((uint32_t)ins & 0xFFFFu);
          pay[k] = np;
          below_h = local;
          below_pay = np;
          if (local > max_score) { max_score = local; max_start = j; max_pay = np; }  // first max
        }
      }
    }
    const uint32_t seq_pos = chunk.seq_starts[h.db_id];
    const uint32_t max_match = max_pay >> 16, max_len = max_pay & 0xFFFFu;
    h.db_start = db_offset - max_start - seq_pos;                     // aligner.cpp:941, :715
    h.db_end = db_offset - seq_pos;                                   // :716
    h.seq_id = (float)max_match / (float)(int)max_len;                // :945
    h.aln_len = max_len;
    h.aln_match = max_match;
    p.hits[p.jobs[job]] = h;
  }
}

// Warp-cooperative TraceBack for long queries (80 < L <= 1024; config 4): one WARP per hit, the
// query rows are split over the lanes in processing order (i = L-1-k: lane l owns rows
// i in [l*RL, (l+1)*RL)), the reverse DP runs as a systolic wavefront: at step t lane l computes
// its rows of column j = t - l and hands its last row (H, payload, running deletion) to lane l+1,
// which is one column behind; the column state of a lane's rows stays in registers ((H<<16|E) and
// (matches<<16|length) per row).  The db residue of a column travels down the lanes with the
// wavefront.  Every lane keeps the first strict maximum of its own cells in scan order; the warp
// then takes the largest score and, among equals, the earliest (column, row) - exactly the cell
// the sequential scan (aligner.cpp:898) keeps.  One thread per hit needs L+1 columns of global
// scratch at these lengths and ran 234 ms for 480 hits of 1000 residues.
template <int RL>
__global__ void __launch_bounds__(256) traceback_warp_kernel(const TracebackParams p) {
  __shared__ int16_t mat[kAlphabet * kAlphabet];
  for (int i = threadIdx.x; i < kAlphabet * kAlphabet; i += blockDim.x) mat[i] = (int16_t)p.matrix[i];
  __syncthreads();
  const uint32_t n_jobs = *p.n_jobs;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const int L = (int)p.query_len;
  const int go = p.open_gap, ge = p.extend_gap;
  for (uint32_t job = gwarp; job < n_jobs; job += n_warps) {
    gm_hit h = p.hits[p.jobs[job]];
    const ChunkRef chunk = p.chunks[h.db_chunk];
    const uint8_t *query = p.queries + (size_t)h.query_id * L;
    const uint32_t db_offset = h.db_end;
    uint32_t len = p.base_len;
    if (db_offset < len) len = db_offset + 1;                         // aligner.cpp:802-805
    uint32_t qk[RL], he[RL], pay[RL];
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      const int i = (int)lane * RL + r;
      qk[r] = i < L ? query[L - 1 - i] : 0xFFu;                       // 0xFF: no such row
      he[r] = 0;
      pay[r] = 0;
    }
    int best = 0;
    uint32_t best_j = 0, best_i = 0, best_pay = 0;
    int out_h = 0, out_del = 0, diag_h = 0;       // last row of this lane / boundary of the previous column
    uint32_t out_pay = 0, diag_pay = 0;
    uint32_t cbuf = kSeqEnd, c_cur = kSeqEnd;
    bool stopped = false;
    const uint32_t steps = len + 31;
    for (uint32_t t = 0; t < steps; ++t) {
      if ((t & 31) == 0) {                                            // next 32 columns, one per lane
        const uint32_t idx = t + lane;
        cbuf = idx < len ? chunk.seq[db_offset - idx] : (uint32_t)kSeqEnd;
      }
      const uint32_t c0 = __shfl_sync(kFull, cbuf, t & 31);
      c_cur = __shfl_up_sync(kFull, c_cur, 1);
      if (lane == 0) c_cur = c0;
      int in_h = __shfl_up_sync(kFull, out_h, 1), in_del = __shfl_up_sync(kFull, out_del, 1);
      uint32_t in_pay = __shfl_up_sync(kFull, out_pay, 1);
      if (lane == 0) { in_h = 0; in_del = 0; in_pay = 0; }            // below row L-1: zeros (:791-797)
      const uint32_t j = t - lane;
      bool active = t >= lane && j < len && !stopped;
      if (active && c_cur == kSeqEnd) { stopped = true; active = false; }   // aligner.cpp:927-929
      if (active) {
        const uint32_t c = c_cur;
        const int16_t *row = mat + c * kAlphabet;
        int temp_score = diag_h, below_h = in_h, del = in_del;
        uint32_t temp_pay = diag_pay, below_pay = in_pay;
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          if (qk[r] != 0xFFu) {
            const uint32_t old = he[r], old_pay = pay[r];
            const int old_h = (int)old >> 16;
            int ins = (int)(int16_t)(old & 0xFFFFu);
            const int s = temp_score + row[qk[r]];
            int local = 0;
            uint32_t np = 0;
            if (s > 0) { local = s; np = temp_pay + (c == qk[r] ? 0x10001u : 0x1u); }
            ins = max(ins + ge, old_h + go);
            if (ins > local) { local = ins; np = old_pay + 1; }
            del = max(del + ge, below_h + go);
            if (del > local) { local = del; np = below_pay + 1; }
            temp_score = old_h;
            temp_pay = old_pay;
            he[r] = ((uint32_t)local << 16) | ((uint32_t)ins & 0xFFFFu);
            pay[r] = np;
            below_h = local;
            below_pay = np;
            if (local > best) { best = local; best_j = j; best_i = lane * RL + r; best_pay = np; }  // first max
          }
        }
        out_h = below_h;
        out_pay = below_pay;
        out_del = del;
      }
      diag_h = in_h;          // boundary of column j becomes the diagonal of column j+1
      diag_pay = in_pay;
    }
    // largest score; among equals the earliest cell in scan order (column, then row)
    unsigned long long key = ((unsigned long long)(uint32_t)best << 32) |
                             (0xFFFFFFFFu - ((best_j << 11) | best_i));
    uint32_t bp = best_pay;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long k2 = __shfl_xor_sync(kFull, key, o);
      const uint32_t p2 = __shfl_xor_sync(kFull, bp, o);
      if (k2 > key) { key = k2; bp = p2; }
    }
    if (lane == 0) {
      const uint32_t max_start = (0xFFFFFFFFu - (uint32_t)key) >> 11;
      const uint32_t seq_pos = chunk.seq_starts[h.db_id];
      const uint32_t max_match = bp >> 16, max_len = bp & 0xFFFFu;
      h.db_start = db_offset - max_start - seq_pos;                   // aligner.cpp:941, :715
      h.db_end = db_offset - seq_pos;                                 // :716
      h.seq_id = (float)max_match / (float)(int)max_len;              // :945
      h.aln_len = max_len;
      h.aln_match = max_match;
      p.hits[p.jobs[job]] = h;
    }
    __syncwarp();
  }
}

// Group TraceBack for short queries (L <= G * RL <= 80): the systolic wavefront of
// traceback_warp_kernel on G lanes per hit instead of 32, so a warp traces 32 / G hits at once and
// no lane idles (at L = 75 the full-warp kernel keeps 19 of 32 lanes busy).  Against one thread per
// hit (traceback_reg_kernel: 150 state registers, 8 warps per SM, one long dependent chain) a lane
// holds RL = ceil(L / G) rows, so ~3x the warps are resident and G chains run per hit.
template <int G, int RL>
__global__ void __launch_bounds__(256) traceback_group_kernel(const TracebackParams p) {
  __shared__ int16_t mat[kAlphabet * kAlphabet];
  for (int i = threadIdx.x; i < kAlphabet * kAlphabet; i += blockDim.x) mat[i] = (int16_t)p.matrix[i];
  __syncthreads();
  constexpr uint32_t kGroups = 32 / G;
  const uint32_t n_jobs = *p.n_jobs;
  const uint32_t lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const int L = (int)p.query_len;
  const int go = p.open_gap, ge = p.extend_gap;
  for (uint32_t job0 = gwarp * kGroups; job0 < n_jobs; job0 += n_warps * kGroups) {
    const uint32_t job = job0 + grp;
    const bool have = job < n_jobs;
    gm_hit h = {};
    ChunkRef chunk = {nullptr, nullptr};
    const uint8_t *query = p.queries;
    uint32_t db_offset = 0, len = 0;
    if (have) {
      h = p.hits[p.jobs[job]];
      chunk = p.chunks[h.db_chunk];
      query = p.queries + (size_t)h.query_id * L;
      db_offset = h.db_end;
      len = p.base_len;
      if (db_offset < len) len = db_offset + 1;                       // aligner.cpp:802-805
    }
    uint32_t qk[RL], he[RL], pay[RL];
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      const int i = (int)gl * RL + r;
      qk[r] = (have && i < L) ? query[L - 1 - i] : 0xFFu;             // 0xFF: no such row
      he[r] = 0;
      pay[r] = 0;
    }
    int best = 0;
    uint32_t best_j = 0, best_i = 0, best_pay = 0;
    int out_h = 0, out_del = 0, diag_h = 0;       // last row of this lane / boundary of the previous column
    uint32_t out_pay = 0, diag_pay = 0;
    uint32_t cbuf = kSeqEnd, c_cur = kSeqEnd;
    bool stopped = false;
    const uint32_t steps = __reduce_max_sync(kFull, len) + G - 1;      // the longest window of the warp's hits
    for (uint32_t t = 0; t < steps; ++t) {
      if ((t % G) == 0) {                                             // next G columns, one per lane of the group
        const uint32_t idx = t + gl;
        cbuf = idx < len ? chunk.seq[db_offset - idx] : (uint32_t)kSeqEnd;
      }
      const uint32_t c0 = __shfl_sync(kFull, cbuf, t % G, G);
      c_cur = __shfl_up_sync(kFull, c_cur, 1, G);
      if (gl == 0) c_cur = c0;
      int in_h = __shfl_up_sync(kFull, out_h, 1, G), in_del = __shfl_up_sync(kFull, out_del, 1, G);
      uint32_t in_pay = __shfl_up_sync(kFull, out_pay, 1, G);
      if (gl == 0) { in_h = 0; in_del = 0; in_pay = 0; }              // below row L-1: zeros (:791-797)
      const uint32_t j = t - gl;
      bool active = t >= gl && j < len && !stopped;
      if (active && c_cur == kSeqEnd) { stopped = true; active = false; }   // aligner.cpp:927-929
      if (active) {
        const uint32_t c = c_cur;
        const int16_t *row = mat + c * kAlphabet;
        int temp_score = diag_h, below_h = in_h, del = in_del;
        uint32_t temp_pay = diag_pay, below_pay = in_pay;
#pragma unroll
        for (int r = 0; r < RL; ++r) {
          if (qk[r] != 0xFFu) {
            const uint32_t old = he[r], old_pay = pay[r];
            const int old_h = (int)old >> 16;
            int ins = (int)(int16_t)(old & 0xFFFFu);
            const int s = temp_score + row[qk[r]];
            int local = 0;
            uint32_t np = 0;
            if (s > 0) { local = s; np = temp_pay + (c == qk[r] ? 0x10001u : 0x1u); }
            ins = max(ins + ge, old_h + go);
            if (ins > local) { local = ins; np = old_pay + 1; }
            del = max(del + ge, below_h + go);
            if (del > local) { local = del; np = below_pay + 1; }
            temp_score = old_h;
            temp_pay = old_pay;
            he[r] = ((uint32_t)local << 16) | ((uint32_t)ins & 0xFFFFu);
            pay[r] = np;
            below_h = local;
            below_pay = np;
            if (local > best) { best = local; best_j = j; best_i = gl * RL + r; best_pay = np; }  // first max
          }
        }
        out_h = below_h;
        out_pay = below_pay;
        out_del = del;
      }
      diag_h = in_h;          // boundary of column j becomes the diagonal of column j+1
      diag_pay = in_pay;
    }
    // largest score; among equals the earliest cell in scan order (column, then row)
    unsigned long long key = ((unsigned long long)(uint32_t)best << 32) |
                             (0xFFFFFFFFu - ((best_j << 11) | best_i));
    uint32_t bp = best_pay;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const unsigned long long k2 = __shfl_xor_sync(kFull, key, o, G);
      const uint32_t p2 = __shfl_xor_sync(kFull, bp, o, G);
      if (k2 > key) { key = k2; bp = p2; }
    }
    if (gl == 0 && have) {
      const uint32_t max_start = (0xFFFFFFFFu - (uint32_t)key) >> 11;
      const uint32_t seq_pos = chunk.seq_starts[h.db_id];
      const uint32_t max_match = bp >> 16, max_len = bp & 0xFFFFu;
      h.db_start = db_offset - max_start - seq_pos;                   // aligner.cpp:941, :715
      h.db_end = db_offset - seq_pos;                                 // :716
      h.seq_id = (float)max_match / (float)(int)max_len;              // :945
      h.aln_len = max_len;
      h.aln_match = max_match;
      p.hits[p.jobs[job]] = h;
    }
    __syncwarp();
  }
}

// Collect the result slots whose TraceBack is pending and whose db chunk is resident here.
__global__ void collect_pending_kernel(const gm_hit *hits, const uint32_t *counts, uint32_t n_queries,
                                       uint32_t cap, const ChunkRef *chunks, uint32_t *jobs,
                                       uint32_t *n_jobs, uint32_t *n_left) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_queries * cap;
       i += gridDim.x * blockDim.x) {
    const uint32_t q = i / cap, k = i - q * cap;
    if (k >= counts[q]) continue;
    const gm_hit &h = hits[i];
    if (h.aln_match != kNoId) continue;
    if (h.db_chunk < GM_MAX_DB_CHUNKS && chunks[h.db_chunk].seq != nullptr)
      jobs[atomicAdd(n_jobs, 1u)] = i;
    else
      atomicAdd(n_left, 1u);   // its db chunk is not resident here: stays pending
  }
}

}  // namespace

cudaError_t collect_pending_launch(const gm_hit *hits, const uint32_t *counts, uint32_t n_queries,
                                   uint32_t cap, const ChunkRef *chunks, uint32_t *jobs,
                                   uint32_t *n_jobs, uint32_t *n_left, int sm_count,
                                   cudaStream_t stream) {
  collect_pending_kernel<<<sm_count * 4, 256, 0, stream>>>(hits, counts, n_queries, cap, chunks, jobs,
                                                           n_jobs, n_left);
  return cudaGetLastError();
}

size_t merge_smem_bytes(uint32_t elems_per_warp) { return (size_t)kMergeWarps * elems_per_warp * 8; }

cudaError_t merge_launch(const MergeParams &p, int sm_count, cudaStream_t stream) {
  const size_t smem = merge_smem_bytes(p.smem_elems);
  cudaError_t err = allow_max_dynamic_smem(merge_kernel);
  if (err != cudaSuccess) return err;
  merge_kernel<<<sm_count, kMergeThreads, smem, stream>>>(p);   // 192 KB of lists: one CTA per SM
  return cudaGetLastError();
}

int traceback_grid(int sm_count) { return sm_count * 8; }
int traceback_threads() { return 128; }

// True when the register-resident kernel covers this query length and gap/score range.
bool traceback_fast_ok(uint32_t query_len, int open_gap, int extend_gap) {
  return query_len <= 80 && open_gap > -8000 && extend_gap > -8000 && open_gap <= 0 && extend_gap <= 0;
}

// True when the warp-cooperative kernel covers this query length and gap/score range (16-bit H, E).
bool traceback_warp_ok(uint32_t query_len, int open_gap, int extend_gap) {
  return query_len <= 1024 && open_gap > -8000 && extend_gap > -8000 && open_gap <= 0 && extend_gap <= 0;
}

cudaError_t traceback_launch(const TracebackParams &p, int sm_count, cudaStream_t stream, bool fast) {
  if (fast && !traceback_fast_ok(p.query_len, p.open_gap, p.extend_gap) &&
      traceback_warp_ok(p.query_len, p.open_gap, p.extend_gap) && p.base_len < (1u << 20)) {
    const int grid = sm_count * 4;
    if (p.query_len <= 128) traceback_warp_kernel<4><<<grid, 256, 0, stream>>>(p);
    else if (p.query_len <= 256) traceback_warp_kernel<8><<<grid, 256, 0, stream>>>(p);
    else if (p.query_len <= 512) traceback_warp_kernel<16><<<grid, 256, 0, stream>>>(p);
    else traceback_warp_kernel<32><<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
  }
  if (fast && traceback_fast_ok(p.query_len, p.open_gap, p.extend_gap)) {
    static const int group = [] { const char *e = getenv("GM_TB_GROUP"); return e ? atoi(e) : 4; }();
    if (group == 4) {       // 4 lanes per hit (8 hits per warp)
      const int grid = sm_count * 8;
      if (p.query_len <= 40) traceback_group_kernel<4, 10><<<grid, 256, 0, stream>>>(p);
      else if (p.query_len <= 64) traceback_group_kernel<4, 16><<<grid, 256, 0, stream>>>(p);
      else if (p.query_len <= 76) traceback_group_kernel<4, 19><<<grid, 256, 0, stream>>>(p);
      else traceback_group_kernel<4, 20><<<grid, 256, 0, stream>>>(p);
      return cudaGetLastError();
    }
    if (group == 8) {       // 8 lanes per hit (4 hits per warp)
      const int grid = sm_count * 8;
      if (p.query_len <= 40) traceback_group_kernel<8, 5><<<grid, 256, 0, stream>>>(p);
      else if (p.query_len <= 64) traceback_group_kernel<8, 8><<<grid, 256, 0, stream>>>(p);
      else traceback_group_kernel<8, 10><<<grid, 256, 0, stream>>>(p);
      return cudaGetLastError();
    }
    const int grid = sm_count * 2;
    if (p.query_len <= 32) traceback_reg_kernel<32><<<grid, 128, 0, stream>>>(p);
    else if (p.query_len <= 64) traceback_reg_kernel<64><<<grid, 128, 0, stream>>>(p);
    else if (p.query_len <= 75) traceback_reg_kernel<75><<<grid, 128, 0, stream>>>(p);
    else traceback_reg_kernel<80><<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
  }
  traceback_kernel<<<traceback_grid(sm_count), traceback_threads(), 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace gm
