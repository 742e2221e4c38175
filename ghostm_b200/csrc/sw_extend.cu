// Windowed affine-gap Smith-Waterman extension (the dominant kernel), sm_100a.
//
// Semantics: reference CalculateScoreCpu, aligner.cpp:545-685 (GPU twin aligner_gpu.cu:369-504):
// for every candidate, local alignment of the L query rows against the db window
// [start-extend, start-extend + L + 2*extend + 2*2^r) clipped to the chunk, a gap of length n
// costing |open| + (n-1)*|extend|, both DP columns reset to 0 at SEQUENCE_END while the running
// maximum survives, result = (max score, db offset of the LAST column that attains it).
//
// Design (DESIGN.md "SW extension"):
//  * inter-candidate SIMD: every thread owns TWO candidates of the same query, one per half of
//    a packed s16x2 word, and runs the whole DP for them with the column state (H+open, E)
//    of up to R=80 query rows in registers.  No shuffles, no shared-memory DP state.
//  * per cell (x2 candidates) 6 ALU-pipe DPX ops: VIADDMNMX.S16x2 (x4, one with .RELU),
//    VIADD.16x2, VIMNMX.S16x2; the vertical (F) recurrence is restated so that its loop-carried
//    dependency is ONE VIADDMNMX per row.
//  * one warp task = 64 candidates of one query, so the query profile T[row][db residue]
//    (16-bit, in shared memory, per warp) is read at bank = residue: conflict free, and the two
//    halves are packed with one IMAD (FMA pipe), off the ALU pipe that bounds the kernel.
//  * SEQUENCE_END columns and clipped windows are handled by a warp-uniform slow path that
//    masks the affected half; queries longer than R rows run as horizontal strips whose
//    boundary row (H+open, F, column max) goes through an L2-resident scratch.
//  * scores fit s16: the host refuses option sets where L * max(matrix) could overflow.
#include <stdlib.h>

#include <type_traits>

#include "gm_common.cuh"

namespace gm {

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t pack2(int v) { return (uint32_t)(v & 0xFFFF) * 0x10001u; }

template <int R>
__global__ void __launch_bounds__(kSwThreads, (R <= 40 ? 2 : 1)) sw_extend_dpx_kernel(const SwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t *matT = reinterpret_cast<int16_t *>(smem_raw);                 // [query residue][db residue]
  uint16_t *prof_all = reinterpret_cast<uint16_t *>(smem_raw + 2048);    // per warp [R][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint16_t *prof = prof_all + warp * (R * 32);

  for (int i = threadIdx.x; i < kAlphabet * kAlphabet; i += blockDim.x) {
    const int c = i >> 5, q = i & 31;  // matrix[db residue * 32 + query residue] (aligner.cpp:612-618)
    matT[q * 32 + c] = (int16_t)p.matrix[i];
  }
  __syncthreads();

  const int go = p.open_gap, ge = p.extend_gap;
  const int gef = go > ge ? go : ge;  // F_{k+1} = max(F_k + max(ge,go), m_k + go), see header
  const uint32_t go_pk = pack2(go), ge_pk = pack2(ge), gef_pk = pack2(gef);
  const uint32_t total_tasks = p.task_prefix[p.n_q];
  const uint32_t gwarp = blockIdx.x * kSwWarps + warp;
  const uint32_t L = p.query_len;

  while (true) {
    uint32_t task = 0;
    if (lane == 0) task = atomicAdd(p.task_counter, 1u);
    task = __shfl_sync(kFull, task, 0);
    if (task >= total_tasks) break;
    uint32_t lo = 0, hi = p.n_q;  // task_prefix[lo] <= task < task_prefix[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (p.task_prefix[mid] <= task) lo = mid; else hi = mid;
    }
    const uint32_t q = p.first_query + lo;
    const uint32_t blk = task - p.task_prefix[lo];
    const uint32_t cnt = p.cand_cnt[q], off = p.cand_off[q];
    const uint8_t *query = p.queries + (size_t)q * L;

    // the two candidates of this lane
    const uint32_t ia = blk * kSwCandPerTask + lane, ib = ia + 32;
    uint32_t wa = 0, wb = 0, offa = 0, offb = 0;
    if (ia < cnt) {
      const uint32_t st = p.cand_start[off + ia];
      offa = st >= p.extend ? st - p.extend : 0u;                       // aligner.cpp:576-579
      wa = min(p.base_len, p.db_len - offa);                            // aligner.cpp:580-583
    }
    if (ib < cnt) {
      const uint32_t st = p.cand_start[off + ib];
      offb = st >= p.extend ? st - p.extend : 0u;
      wb = min(p.base_len, p.db_len - offb);
    }
    const uint8_t *pa = p.db + offa, *pb = p.db + offb;
    const uint32_t wmax = __reduce_max_sync(kFull, wa > wb ? wa : wb);

    uint32_t best = go_pk;           // running max of H+open, per half (max_score = 0)
    uint32_t enda = 0, endb = 0;     // column of the last maximum (aligner.cpp:650-653)

    for (uint32_t strip = 0; strip < p.n_strips; ++strip) {
      // ---- per-warp query profile of this strip: T[k][c] = matrix[c][query[row]] - open
      __syncwarp();
#pragma unroll 4
      for (int k = 0; k < R; ++k) {
        const uint32_t row = strip * R + k;
        int v = -16384 - go;  // padding rows below the query never score
        if (row < L) v = (int)matT[(int)query[row] * 32 + lane] - go;
        prof[k * 32 + lane] = (uint16_t)v;
      }
      __syncwarp();

      const bool first_strip = strip == 0, last_strip = strip + 1 == p.n_strips;
      uint32_t *scr = p.strip_scratch + (size_t)gwarp * p.base_len * 96 + lane;

      uint32_t hgo[R], e[R];
#pragma unroll
      for (int k = 0; k < R; ++k) { hgo[k] = go_pk; e[k] = 0u; }        // aligner.cpp:587-590
      uint32_t top_prev = go_pk;      // H+open of the row above the strip, previous column
      uint32_t ca = wa > 0 ? pa[0] : kSeqEnd, cb = wb > 0 ? pb[0] : kSeqEnd;

      for (uint32_t j = 0; j < wmax; ++j) {
        // prefetch the next column's residues; columns beyond a window read as SEQUENCE_END
        const uint32_t na = (j + 1 < wa) ? pa[j + 1] : (uint32_t)kSeqEnd;
        const uint32_t nb = (j + 1 < wb) ? pb[j + 1] : (uint32_t)kSeqEnd;
        uint32_t top = go_pk, f = gef_pk, cmax = 0x80008000u;
        if (!first_strip) {
          top = scr[(j * 3 + 0) * 32];
          f = scr[(j * 3 + 1) * 32];
          cmax = scr[(j * 3 + 2) * 32];
        }
        const uint16_t *ta = prof + ca, *tb = prof + cb;
        // Row k+1's insertion/diagonal part is issued before row k's H is written, so that the
        // previous column's H+open of row k is dead by then and hgo[k] is updated in place
        // (no register rotation at the loop back-edge).
        e[0] = __viaddmax_s16x2(e[0], ge_pk, hgo[0]);                    // aligner.cpp:623-627
        uint32_t m = __viaddmax_s16x2_relu(                               // :617-620,:629-631
            top_prev, (uint32_t)tb[0] * 65536u + (uint32_t)ta[0], e[0]);
#pragma unroll
        for (int k = 0; k < R; ++k) {
          uint32_t m_next = 0;
          if (k + 1 < R) {
            const uint32_t s = (uint32_t)tb[(k + 1) * 32] * 65536u + (uint32_t)ta[(k + 1) * 32];
            e[k + 1] = __viaddmax_s16x2(e[k + 1], ge_pk, hgo[k + 1]);
            m_next = __viaddmax_s16x2_relu(hgo[k], s, e[k + 1]);
          }
          const uint32_t mgo = __vadd2(m, go_pk);
          hgo[k] = __viaddmax_s16x2(f, go_pk, mgo);                      // H = max(m, F)  :641-643
          f = __viaddmax_s16x2(f, gef_pk, mgo);                          // :634-639
          cmax = __vmaxs2(cmax, hgo[k]);
          m = m_next;
        }
        const bool xa = ca == kSeqEnd, xb = cb == kSeqEnd;
        if (__any_sync(kFull, xa | xb)) {                                // aligner.cpp:664-669
          const uint32_t keep = (xa ? 0u : 0x0000FFFFu) | (xb ? 0u : 0xFFFF0000u);
          const uint32_t rst = go_pk & ~keep;
#pragma unroll
          for (int k = 0; k < R; ++k) {
            hgo[k] = (hgo[k] & keep) | rst;
            e[k] &= keep;
          }
          cmax = (cmax & keep) | (0x80008000u & ~keep);
        }
        top_prev = top;
        if (!last_strip) {
          scr[(j * 3 + 0) * 32] = hgo[R - 1];
          scr[(j * 3 + 1) * 32] = f;
          scr[(j * 3 + 2) * 32] = cmax;
        } else {
          bool ph, pl;
          best = __vibmax_s16x2(cmax, best, &ph, &pl);                  // ">=": last maximum wins
          if (pl) enda = j;
          if (ph) endb = j;
        }
        ca = na;
        cb = nb;
      }
    }
    if (ia < cnt) {
      p.cand_score[off + ia] = (uint32_t)((int)(int16_t)(best & 0xFFFFu) - go);
      p.cand_end[off + ia] = offa + enda;
    }
    if (ib < cnt) {
      p.cand_score[off + ib] = (uint32_t)((int)(int16_t)(best >> 16) - go);
      p.cand_end[off + ib] = offb + endb;
    }
    const unsigned long long cells =
        (unsigned long long)__reduce_add_sync(kFull, wa + wb) * (unsigned long long)L;
    if (lane == 0) atomicAdd(p.cells, cells);
  }
}

// (max, last column) of two disjoint sets of DP cells, per packed half: the larger maximum, on a tie
// the later column (aligner.cpp:650-653 keeps the LAST column that attains the maximum).
__device__ __forceinline__ void merge_max(uint32_t &best, uint32_t &enda, uint32_t &endb, uint32_t obest,
                                          uint32_t oenda, uint32_t oendb) {
  const int ma = (int)(int16_t)(best & 0xFFFFu), oa = (int)(int16_t)(obest & 0xFFFFu);
  const int mb = (int)(int16_t)(best >> 16), ob = (int)(int16_t)(obest >> 16);
  enda = ma > oa ? enda : (oa > ma ? oenda : max(enda, oenda));
  endb = mb > ob ? endb : (ob > mb ? oendb : max(endb, oendb));
  best = ((uint32_t)(mb > ob ? mb : ob) << 16) | ((uint32_t)(ma > oa ? ma : oa) & 0xFFFFu);
}

// Pair variant for queries of 41 residues and more: TWO adjacent lanes share a pair of candidates, the even
// lane holds rows [0, RH) and the odd lane rows [RH, 2 RH) of the same DP, one column behind (a
// two-stage systolic array): what leaves the even lane's last row at column j - H+open and the
// vertical F - reaches the odd lane by one shuffle each and is consumed at its column j one step
// later.  Every thread keeps RH <= 40 rows (128 registers: 4 warps per scheduler instead of 2 for a
// full-height column), the boundary row never leaves the register file (the strip-mined kernel
// sends it through an L2 scratch and walks the window twice), and each lane keeps its own running
// maximum: (max, last column) of the candidate is the larger of the two, on a tie the later column.
// One warp task = 32 candidates of one query.
template <int RH, bool ONE_STRIP>
__global__ void __launch_bounds__(kSwThreads, 2) sw_extend_pair_kernel(const SwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t *matT = reinterpret_cast<int16_t *>(smem_raw);                 // [query residue][db residue]
  uint16_t *prof_all = reinterpret_cast<uint16_t *>(smem_raw + 2048);    // per warp [2 RH][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t odd = lane & 1;
  uint16_t *prof = prof_all + warp * ((2 * RH + 1) * 32);

  for (int i = threadIdx.x; i < kAlphabet * kAlphabet; i += blockDim.x) {
    const int c = i >> 5, q = i & 31;  // matrix[db residue * 32 + query residue] (aligner.cpp:612-618)
    matT[q * 32 + c] = (int16_t)p.matrix[i];
  }
  __syncthreads();

  const int go = p.open_gap, ge = p.extend_gap;
  const int gef = go > ge ? go : ge;  // F_{k+1} = max(F_k + max(ge,go), m_k + go), see header
  const uint32_t go_pk = pack2(go), ge_pk = pack2(ge), gef_pk = pack2(gef);
  const uint32_t total_tasks = p.task_prefix[p.n_q];
  const uint32_t L = p.query_len;

  while (true) {
    uint32_t task = 0;
    if (lane == 0) task = atomicAdd(p.task_counter, 1u);
    task = __shfl_sync(kFull, task, 0);
    if (task >= total_tasks) break;
    uint32_t lo = 0, hi = p.n_q;  // task_prefix[lo] <= task < task_prefix[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (p.task_prefix[mid] <= task) lo = mid; else hi = mid;
    }
    const uint32_t q = p.first_query + lo;
    const uint32_t blk = task - p.task_prefix[lo];
    const uint32_t cnt = p.cand_cnt[q], off = p.cand_off[q];
    const uint8_t *query = p.queries + (size_t)q * L;

    // the two candidates of this lane PAIR
    const uint32_t ia = blk * 32u + (lane >> 1), ib = ia + 16;
    uint32_t wa = 0, wb = 0, offa = 0, offb = 0;
    if (ia < cnt) {
      const uint32_t st = p.cand_start[off + ia];
      offa = st >= p.extend ? st - p.extend : 0u;                       // aligner.cpp:576-579
      wa = min(p.base_len, p.db_len - offa);                            // aligner.cpp:580-583
    }
    if (ib < cnt) {
      const uint32_t st = p.cand_start[off + ib];
      offb = st >= p.extend ? st - p.extend : 0u;
      wb = min(p.base_len, p.db_len - offb);
    }
    const uint8_t *pa = p.db + offa, *pb = p.db + offb;
    const uint32_t wmax = __reduce_max_sync(kFull, wa > wb ? wa : wb);

    uint32_t best = go_pk;           // running max of H+open over this lane's rows (all strips), per half
    uint32_t enda = 0, endb = 0;     // column of the last maximum (aligner.cpp:650-653)
    // the boundary row between two strips of 2 RH rows: (H+open, F) per column and lane pair, L2 resident
    uint32_t *scr = ONE_STRIP ? nullptr
                              : p.strip_scratch + (size_t)(blockIdx.x * kSwWarps + warp) * p.base_len * 32 + (lane >> 1);
    constexpr int kOddRow = RH | 1;

    const uint32_t n_strips = ONE_STRIP ? 1u : p.n_strips;
    for (uint32_t strip = 0; strip < n_strips; ++strip) {
      // per-warp query profile T[k][c] = matrix[c][query[row]] - open for the 2 RH rows of the strip.
      // A table row is 64 bytes = 16 banks; the odd lanes' half starts at an ODD table row, so that in
      // every load the even lanes (row k) and the odd lanes (row k + RH) hit different halves of the
      // 32 banks
      __syncwarp();
#pragma unroll 4
      for (int k = 0; k < 2 * RH; ++k) {
        const uint32_t row = strip * (2 * RH) + k;
        int v = -16384 - go;  // padding rows below the query, and the SEQUENCE_END column, never score
        if (row < L && lane != kSeqEnd) v = (int)matT[(int)query[row] * 32 + lane] - go;
        prof[(k < RH ? k : k - RH + kOddRow) * 32 + lane] = (uint16_t)v;
      }
      __syncwarp();
      const uint16_t *myprof = prof + odd * (kOddRow * 32);
      const bool first_strip = ONE_STRIP || strip == 0, last_strip = ONE_STRIP || strip + 1 == n_strips;

      uint32_t hgo[RH], e[RH];
#pragma unroll
      for (int k = 0; k < RH; ++k) { hgo[k] = go_pk; e[k] = 0u; }        // aligner.cpp:587-590
      uint32_t top_prev = go_pk;       // H+open of the row above this lane's rows, previous column
      uint32_t send_top = go_pk, send_f = gef_pk;
      // maximum of THIS strip's rows and its last column (">=" in column order); merged below, because
      // a later strip may reach the same maximum in an earlier column
      uint32_t sbest = go_pk, sea = 0, seb = 0;
      // the odd lane is one column behind: at step t it works on column t - 1 (nothing at t = 0)
      uint32_t ca = (!odd && wa > 0) ? pa[0] : (uint32_t)kSeqEnd, cb = (!odd && wb > 0) ? pb[0] : (uint32_t)kSeqEnd;

      for (uint32_t t = 0; t <= wmax; ++t) {
        const uint32_t j = t - odd;                     // this lane's column (wraps for the odd lane at t = 0)
        const uint32_t jn = j + 1;                      // next column; columns beyond a window read as SEQUENCE_END
        const uint32_t na = (jn < wa) ? pa[jn] : (uint32_t)kSeqEnd;
        const uint32_t nb = (jn < wb) ? pb[jn] : (uint32_t)kSeqEnd;
        // what left the even lane's last row at this column, one step ago
        const uint32_t recv_top = __shfl_up_sync(kFull, send_top, 1);
        const uint32_t recv_f = __shfl_up_sync(kFull, send_f, 1);
        uint32_t top = go_pk, f = gef_pk;
        if (odd) {
          top = recv_top;
          f = recv_f;
        } else if (!first_strip && j < wmax) {          // the strip above, same column
          top = scr[(j * 2 + 0) * 16];
          f = scr[(j * 2 + 1) * 16];
        }
        uint32_t cmax = 0x80008000u;
        const uint16_t *ta = myprof + ca, *tb = myprof + cb;
        const bool xa = ca == kSeqEnd, xb = cb == kSeqEnd;
        const uint32_t keep = (xa ? 0u : 0x0000FFFFu) | (xb ? 0u : 0xFFFF0000u);
        // One column.  END_COL: a half whose db residue is SEQUENCE_END (aligner.cpp:664-669: both
        // columns reset to 0, the running maximum survives).  Its profile row is -16384 - open, so
        // the diagonal term is dead; with E masked to 0 BEFORE it is used, m = 0, F stays below
        // open and H + open comes out as `open` by itself: the reset costs one LOP3 per row inside
        // the column instead of two after it.
        auto column = [&](auto end_col) {
          constexpr bool END_COL = decltype(end_col)::value;
          e[0] = __viaddmax_s16x2(e[0], ge_pk, hgo[0]);                  // aligner.cpp:623-627
          if (END_COL) e[0] &= keep;
          uint32_t m = __viaddmax_s16x2_relu(                             // :617-620,:629-631
              top_prev, (uint32_t)tb[0] * 65536u + (uint32_t)ta[0], e[0]);
#pragma unroll
          for (int k = 0; k < RH; ++k) {
            uint32_t m_next = 0;
            if (k + 1 < RH) {
              const uint32_t s = (uint32_t)tb[(k + 1) * 32] * 65536u + (uint32_t)ta[(k + 1) * 32];
              e[k + 1] = __viaddmax_s16x2(e[k + 1], ge_pk, hgo[k + 1]);
              if (END_COL) e[k + 1] &= keep;
              m_next = __viaddmax_s16x2_relu(hgo[k], s, e[k + 1]);
            }
            const uint32_t mgo = __vadd2(m, go_pk);
            hgo[k] = __viaddmax_s16x2(f, go_pk, mgo);                    // H = max(m, F)  :641-643
            f = __viaddmax_s16x2(f, gef_pk, mgo);                        // :634-639
            cmax = __vmaxs2(cmax, hgo[k]);
            m = m_next;
          }
        };
        if (__any_sync(kFull, xa | xb)) {
          column(std::true_type());
          cmax = (cmax & keep) | (0x80008000u & ~keep);                  // an END column never holds the maximum
        } else {
          column(std::false_type());
        }
        top_prev = top;
        send_top = hgo[RH - 1];
        send_f = f;
        if (odd && !last_strip && j < wmax) {           // bottom row of the strip for the strip below
          scr[(j * 2 + 0) * 16] = hgo[RH - 1];
          scr[(j * 2 + 1) * 16] = f;
        }
        bool ph, pl;
        sbest = __vibmax_s16x2(cmax, sbest, &ph, &pl);                  // ">=": last maximum wins
        if (pl) sea = j;
        if (ph) seb = j;
        ca = na;
        cb = nb;
      }
      merge_max(best, enda, endb, sbest, sea, seb);
    }
    // the candidate's result: the larger of the two lanes' maxima, on a tie the later column
    {
      const uint32_t obest = __shfl_xor_sync(kFull, best, 1);
      const uint32_t oenda = __shfl_xor_sync(kFull, enda, 1), oendb = __shfl_xor_sync(kFull, endb, 1);
      merge_max(best, enda, endb, obest, oenda, oendb);
      if (!odd) {
        if (ia < cnt) {
          p.cand_score[off + ia] = (uint32_t)((int)(int16_t)(best & 0xFFFFu) - go);
          p.cand_end[off + ia] = offa + enda;
        }
        if (ib < cnt) {
          p.cand_score[off + ib] = (uint32_t)((int)(int16_t)(best >> 16) - go);
          p.cand_end[off + ib] = offb + endb;
        }
      }
    }
    const unsigned long long cells =
        (unsigned long long)__reduce_add_sync(kFull, odd ? 0u : wa + wb) * (unsigned long long)L;
    if (lane == 0) atomicAdd(p.cells, cells);
  }
}

// One thread per candidate, 32-bit scores, DP columns in local memory: the direct form of the
// reference loop.  Used when the packed s16 kernel's score range check fails (exotic matrices).
__global__ void __launch_bounds__(128) sw_extend_s32_kernel(const SwParams p) {
  const uint32_t L = p.query_len;
  for (uint32_t qi = blockIdx.x; qi < p.n_q; qi += gridDim.x) {
    const uint32_t q = p.first_query + qi, off = p.cand_off[q], cnt = p.cand_cnt[q];
    const uint8_t *query = p.queries + (size_t)q * L;
    for (uint32_t ci = threadIdx.x; ci < cnt; ci += blockDim.x) {
      const uint32_t i = off + ci;
      const uint32_t st = p.cand_start[i];
      const uint32_t dbo = st >= p.extend ? st - p.extend : 0u;
      const uint32_t w = min(p.base_len, p.db_len - dbo);
      int h[1024 + 1], ins[1024 + 1];
      for (uint32_t k = 0; k <= L; ++k) { h[k] = 0; ins[k] = 0; }
      int best = 0;
      uint32_t end = 0;
      for (uint32_t j = 0; j < w; ++j) {
        const uint8_t c = p.db[dbo + j];
        if (c != kSeqEnd) {
          const int32_t *row = p.matrix + c * kAlphabet;
          int diag = 0, del = 0;
          for (uint32_t k = 1; k <= L; ++k) {
            int v = max(0, diag + row[query[k - 1]]);
            ins[k] = max(ins[k] + p.extend_gap, h[k] + p.open_gap);
            v = max(v, ins[k]);
            del = max(del + p.extend_gap, h[k - 1] + p.open_gap);
            v = max(v, del);
            diag = h[k];
            h[k] = v;
            if (v >= best) { best = v; end = j; }
          }
        } else {
          for (uint32_t k = 0; k <= L; ++k) { h[k] = 0; ins[k] = 0; }
        }
      }
      p.cand_score[i] = (uint32_t)best;
      p.cand_end[i] = dbo + end;
      atomicAdd(p.cells, (unsigned long long)w * L);
    }
  }
}

template <int R>
cudaError_t launch_dpx(const SwParams &p, int sm_count, cudaStream_t stream) {
  const size_t smem = 2048 + (size_t)kSwWarps * R * 32 * sizeof(uint16_t);
  cudaError_t err = allow_max_dynamic_smem(sw_extend_dpx_kernel<R>);
  if (err != cudaSuccess) return err;
  sw_extend_dpx_kernel<R><<<sm_count * (R <= 40 ? 2 : 1), kSwThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int RH>
cudaError_t launch_pair(const SwParams &p, int sm_count, cudaStream_t stream) {
  const size_t smem = 2048 + (size_t)kSwWarps * (2 * RH + 1) * 32 * sizeof(uint16_t);
  cudaError_t err;
  if (p.n_strips == 1) {
    if ((err = allow_max_dynamic_smem(sw_extend_pair_kernel<RH, true>)) != cudaSuccess) return err;
    sw_extend_pair_kernel<RH, true><<<sm_count * 2, kSwThreads, smem, stream>>>(p);
  } else {
    if ((err = allow_max_dynamic_smem(sw_extend_pair_kernel<RH, false>)) != cudaSuccess) return err;
    sw_extend_pair_kernel<RH, false><<<sm_count * 2, kSwThreads, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace

// Rows per lane of the pair kernel (two lanes per candidate pair, 32 candidates per warp task) and the
// number of strips of 2 x rows for a query length; 0 when the single-lane kernel is to be used
// (queries of up to 40 residues already run at 128 registers there).
int sw_pair_rows(uint32_t query_len, uint32_t *n_strips) {
  static const int on = [] { const char *e = getenv("GM_SW_PAIR"); return e ? atoi(e) : 1; }();
  if (!on || query_len <= 40) return 0;
  static const int kHalf[] = {24, 28, 32, 36, 38, 40};
  int best = 0;
  uint32_t best_rows = 0, best_strips = 0;
  for (int r : kHalf) {      // fewest padded rows; among equals the taller strip (fewer boundary rows)
    const uint32_t strips = (query_len + 2 * r - 1) / (2 * r), rows = strips * 2 * r;
    if (best == 0 || rows < best_rows || (rows == best_rows && strips <= best_strips)) {
      best = r;
      best_rows = rows;
      best_strips = strips;
    }
  }
  *n_strips = best_strips;
  return best;
}

cudaError_t sw_extend_pair_launch(const SwParams &p, int rows, int sm_count, cudaStream_t stream) {
  switch (rows) {
    case 24: return launch_pair<24>(p, sm_count, stream);
    case 28: return launch_pair<28>(p, sm_count, stream);
    case 32: return launch_pair<32>(p, sm_count, stream);
    case 36: return launch_pair<36>(p, sm_count, stream);
    case 38: return launch_pair<38>(p, sm_count, stream);
    case 40: return launch_pair<40>(p, sm_count, stream);
    default: return cudaErrorInvalidValue;
  }
}

// rows per strip for a query length: the smallest instantiated R that covers L in
// ceil(L / kSwMaxRows) strips.
int sw_rows_per_strip(uint32_t query_len, uint32_t *n_strips) {
  static const int kRows[] = {16, 25, 32, 38, 40, 48, 64, 75, 80};
  if (const char *env = getenv("GM_SW_ROWS")) {  // tuning experiments: force rows per strip
    const int r = atoi(env);
    for (int k : kRows)
      if (k == r) { *n_strips = (query_len + r - 1) / r; return r; }
  }
  const uint32_t strips = (query_len + kSwMaxRows - 1) / kSwMaxRows;
  if (strips == 1 && query_len > 64) {
    // two half-height strips at 128 registers (4 warps per scheduler instead of 2) beat one full-height
    // strip when they fit the query almost exactly: L = 75 runs 2 x 38 rows at 4.58 instead of 4.47 TCUPS
    // in spite of the boundary row going through the L2 scratch
    const uint32_t half = (query_len + 1) / 2;
    for (int r : kRows)
      if ((uint32_t)r >= half && r <= 40 && 2u * r * 100u <= query_len * 102u) { *n_strips = 2; return r; }
  }
  const uint32_t need = (query_len + strips - 1) / strips;
  for (int r : kRows)
    if ((uint32_t)r >= need) { *n_strips = (query_len + r - 1) / r; return r; }
  *n_strips = strips;
  return kSwMaxRows;
}

cudaError_t sw_extend_launch(const SwParams &p, int rows, int sm_count, cudaStream_t stream) {
  switch (rows) {
    case 16: return launch_dpx<16>(p, sm_count, stream);
    case 25: return launch_dpx<25>(p, sm_count, stream);
    case 32: return launch_dpx<32>(p, sm_count, stream);
    case 38: return launch_dpx<38>(p, sm_count, stream);
    case 40: return launch_dpx<40>(p, sm_count, stream);
    case 48: return launch_dpx<48>(p, sm_count, stream);
    case 64: return launch_dpx<64>(p, sm_count, stream);
    case 75: return launch_dpx<75>(p, sm_count, stream);
    case 80: return launch_dpx<80>(p, sm_count, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t sw_extend_s32_launch(const SwParams &p, int sm_count, cudaStream_t stream) {
  if (p.query_len > 1024) return cudaErrorInvalidValue;
  sw_extend_s32_kernel<<<sm_count * 8, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace gm
