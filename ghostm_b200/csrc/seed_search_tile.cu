// Tiled seed search (threshold 2, list_len <= 64): the default seed-lookup + region-count filter.
//
// Semantics as in seed_search.cu (reference SearchNextCpu, aligner.cpp:418-509): with cnt(d) the
// number of lists (query k-mer offsets j) that have a position p >= j*shift in region
// d = (p - j*shift) >> r, every occupied region - and the virtual region 0, aligner.cpp:451 - emits
// the candidate d << r iff cnt(d) + cnt(d+1) >= 2, candidates ascending per query.
//
// Data layout and division of labour (every index position is touched ONCE, by ~1 warp instruction):
//
//   * the db chunk's position space is cut into a few large tiles of Wd = 31 * nw * 2^r positions
//     (nw = words of the occupancy bitmap, ~90 KB of shared memory per CTA, two CTAs per SM) and the
//     index carries, next to keys_count/positions (index.h:105-114), a SPLIT table
//     split[key * n_tiles + T] = first entry of key's list that is >= T * Wd (built once per chunk on
//     the device): the slice of list j that falls into tile T is two table reads, and a query's 36
//     rows of it are staged in shared memory once per query together with, per tile, the prefix sums
//     of the slices' lengths in STEPS of 31 positions;
//   * a CTA owns one query and walks the tiles in ascending order.  Within a tile the steps are
//     dealt in equal contiguous shares to the MARKER warps; each warp turns its share into step
//     descriptors itself (one lane per step: binary search in the prefix sums) - no serial table
//     build, no barrier - and streams them kTlUnroll loads deep.  Lane 0 of a step reads the
//     predecessor of the step's first position, so "first of its list in the region" (a MARK) is one
//     shuffle and one compare;
//   * a bitmap word holds 31 regions plus, in bit 31, a copy of the first region of the next word, so
//     ANY two adjacent regions share a word.  A mark does one atomicOr on its word (a second one on
//     the previous word when it sits in bit 0) and looks at the value the atomic returns: its own bit
//     already set = a second list in the region; the bit above = right neighbour occupied; the bit
//     below = left neighbour occupied.  Atomics on one word are totally ordered, so of two marks that
//     make a region emit the LATER one always sees the earlier one: every emission is detected exactly
//     where it happens, in one pass;
//   * lists of different tiles meet in the hc topmost words of a tile (a position p >= T * Wd of
//     list j lies up to j*shift positions below its tile in region space): those words are carried
//     into the next tile's bitmap instead of being cleared, as earlier arrivals;
//   * detected emissions are rare (~1.3 % of the marks).  They become bits of the CTA's EMIT bitmap
//     (absolute region space, global memory, L2 resident, all zero between queries - duplicates
//     vanish) and of a summary ring in shared memory (one bit per 16-byte group of emit words).  One
//     SCANNER warp follows the markers one tile behind: it walks the summary bits of the range that
//     has become final, fetches exactly the flagged groups, clears them and appends the regions in
//     ascending order to the query's staging area.  No capacity anywhere: a query with more
//     candidates than the staging area holds is counted first and then redone straight into its
//     slice of the output.  One atomicAdd on the global cursor per query; every query's candidates
//     are contiguous and ascending.
#include "gm_common.cuh"

#include <stdlib.h>

namespace gm {

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kTlLists = 64;
constexpr int kTlMaxTiles = 64;
constexpr int kTlUnroll = 4;
constexpr int kTlFlist = 1024;               // scanner: flagged groups per batch (one summary row at most)
constexpr uint32_t kTlMagic31 = 138547333u;  // ceil(2^32 / 31): x / 31 == umulhi(x, magic) for x < 2^27

struct TileShared {
  uint32_t lbeg[kTlLists];               // first usable entry of list j (aligner.cpp:430-431)
  uint32_t query, stage_n;
  uint32_t dummy[32];                    // per-lane target of the atomics of lanes that carry no mark
  unsigned long long base;
  unsigned long long visited;
};

__device__ __forceinline__ uint32_t tl_get_key(const uint8_t *s, uint32_t seed) {  // index.h:86-101
  uint32_t key = 0;
  for (uint32_t i = 0; seed != 0; ++i, seed >>= 1)
    if (seed & 1) key = (key << kCharBits) | s[i];
  return key;
}

__device__ __forceinline__ uint32_t tl_div31(uint32_t x) { return __umulhi(x, kTlMagic31); }

__device__ __forceinline__ uint32_t tl_smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// atomicOr on a shared-memory word given by its shared-window address
__device__ __forceinline__ uint32_t atomicOr_shared(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

__device__ __forceinline__ void tl_bar_markers(uint32_t n_threads) {   // barrier 1: the marker warps only
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// split[key * n_tiles + T] = lower bound of T * tile_pos in key's position list; one extra entry at
// the end (= positions_len), so that entry (key, n_tiles) is the end of key's list for every key.
__global__ void split_build_kernel(const uint32_t *__restrict__ keys_count, uint32_t n_keys,
                                   const uint32_t *__restrict__ positions, uint32_t n_tiles,
                                   uint32_t tile_pos, uint32_t *__restrict__ split) {
  const size_t total = (size_t)n_keys * n_tiles;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= total;
       i += (size_t)gridDim.x * blockDim.x) {
    if (i == total) { split[i] = keys_count[n_keys]; continue; }
    const uint32_t key = (uint32_t)(i / n_tiles), T = (uint32_t)(i - (size_t)key * n_tiles);
    uint32_t lo = keys_count[key], hi = keys_count[key + 1];
    const unsigned long long target = (unsigned long long)T * tile_pos;
    if (T == 0) hi = lo;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (positions[mid] < target) lo = mid + 1; else hi = mid;
    }
    split[i] = lo;
  }
}

// NW warps: warps 0 .. NW-2 mark, warp NW-1 scans.
template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) seed_search_tile_kernel(const SearchParams p) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ TileShared sh;
  constexpr uint32_t kThreads = NW * 32;
  constexpr uint32_t NM = NW - 1;                        // marker warps
  constexpr uint32_t kMarkers = NM * 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const uint32_t warp = __shfl_sync(kFull, tid >> 5, 0);   // provably warp-uniform for the compiler
  const uint32_t r = p.log_region, nw = p.tl_nw, hc = p.tl_hc, nT = p.tl_tiles, LL = p.list_len;
  const uint32_t rs_mask = p.tl_ring - 1;
  const uint32_t tile_pos = (31u * nw) << r;             // Wd
  const uint32_t occ_words = (nw + hc + 3u) & ~3u;
  uint32_t *occ = dyn;                                   // [nw + hc] 31 regions + 1 overlap bit per word
  uint32_t *sring = occ + occ_words;                     // [tl_ring] summary: bit g = emit group g is flagged
  uint4 *wtab = reinterpret_cast<uint4 *>(sring + p.tl_ring);   // [NM][32] step descriptors, per marker warp
  uint32_t *flist = reinterpret_cast<uint32_t *>(wtab + NM * 32);   // [kTlFlist] scanner: flagged groups
  uint32_t *bounds = flist + kTlFlist;                   // [LL][nT + 1] slices of this query
  uint32_t *pre = bounds + LL * (nT + 1);                // [nT][LL + 1] exclusive prefix of the slices' steps
  uint32_t *stage = p.staging + (size_t)blockIdx.x * p.staging_cap;
  uint32_t *emap = p.tl_emap + (size_t)blockIdx.x * p.tl_emap_stride;   // absolute emit words, all zero
  uint32_t occ_s = tl_smem_addr(occ);
  asm volatile("mov.u32 %0, %0;" : "+r"(occ_s));   // keep the window address in a register
  const uint32_t dummy_s = tl_smem_addr(&sh.dummy[lane]);
  const uint32_t *__restrict__ positions = p.positions;

  for (uint32_t i = tid; i < p.tl_ring; i += kThreads) sring[i] = 0;
  if (tid == 0) sh.visited = 0;
  __syncthreads();

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    const uint32_t q = sh.query;
    if (q >= p.n_queries) break;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    // ---- phase 0: the query's rows of the split table; leading positions < j*shift dropped
    if (tid < LL) {
      const uint32_t j = tid, off = j * p.shift;
      const uint32_t key = tl_get_key(query + off, p.seed);
      const uint32_t *row = p.split + (size_t)key * nT;
      uint32_t *bj = bounds + j * (nT + 1);
      for (uint32_t T = 0; T <= nT; ++T) bj[T] = __ldg(row + T);
      uint32_t b = bj[0];
      const uint32_t e1 = bj[1];
      while (b < e1 && positions[b] < off) ++b;                        // aligner.cpp:430-431
      bj[0] = b;
      sh.lbeg[j] = b;
      if (bj[nT] > b) atomicAdd(&sh.visited, (unsigned long long)(bj[nT] - b));
    }
    __syncthreads();
    // per tile: exclusive prefix sums of the slices' step counts (31 positions per step)
    for (uint32_t T = warp; T < nT; T += NW) {
      uint32_t carry = 0;
      uint32_t *pT = pre + T * (LL + 1);
#pragma unroll
      for (int jj = 0; jj < kTlLists / 32; ++jj) {
        const uint32_t j = lane + 32 * jj;
        uint32_t st = 0;
        if (j < LL) st = tl_div31(bounds[j * (nT + 1) + T + 1] - bounds[j * (nT + 1) + T] + 30u);
        uint32_t incl = st;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += t;
        }
        if (j < LL) pT[j] = carry + incl - st;
        carry += __shfl_sync(kFull, incl, 31);
      }
      if (lane == 0) pT[LL] = carry;
    }

    // Attempt 0 stages the candidates in the CTA's staging area; a query with more candidates than
    // it holds is redone with the exact count known, straight into its slice of the output.
    uint32_t *out = stage;
    uint32_t out_cap = p.staging_cap;
    bool direct = false;
    uint32_t n = 0;
    while (true) {
      if (tid == 0) sh.stage_n = 0;
      __syncthreads();   // bounds, pre, stage_n

      if (warp < NM) {
        // =============================== marker warps ===============================
        uint4 *wt = wtab + warp * 32;
        uint32_t pva[kTlUnroll], pvb[kTlUnroll];
        uint32_t s0 = 0, s_hi = 0, n_r = 0;

        // one lane per step of the round [s0, s0 + n_r) of tile T: which list, which step of the list
        auto build_round = [&](uint32_t T) {
          const uint32_t *pT = pre + T * (LL + 1);
          const uint32_t cb = T * tile_pos - ((31u * hc) << r);   // wraps for T == 0: only differences matter
          uint4 d = make_uint4(0, 0, 0, 0);     // null step: every lane reads positions[0], no mark
          if (lane < n_r) {
            const uint32_t s = s0 + lane;
            uint32_t lo = 0, hi = LL;           // pT[lo] <= s < pT[hi]
#pragma unroll
            for (int it = 0; it < 6; ++it) {
              const uint32_t mid = (lo + hi) >> 1;
              if (pT[mid] <= s) lo = mid; else hi = mid;
            }
            const uint32_t j = lo, k = s - pT[j];
            const uint32_t *bj = bounds + j * (nT + 1) + T;
            const uint32_t first = bj[0] + 31u * k;
            const uint32_t c = cb + j * p.shift;
            // lane i of the step reads entry first - 1 + i: lane 0 is the predecessor of the step's
            // first position; without one (start of the list) it gets a region no position can have
            d = make_uint4(first - 1u, bj[1] - 1u, c, first == sh.lbeg[j] ? c ^ 0x80000000u : c);
          }
          __syncwarp();
          wt[lane] = d;
          __syncwarp();
        };
        // this warp's contiguous share of tile T's steps; first round built, its first batch in flight
        auto open_tile = [&](uint32_t T) {
          const uint32_t S = pre[T * (LL + 1) + LL];
          const uint32_t base = S / NM, extra = S - base * NM;   // shares differ by one step at most
          s0 = warp * base + min(warp, extra);
          s_hi = s0 + base + (warp < extra ? 1u : 0u);
          n_r = min(32u, s_hi - s0);
          if (n_r) build_round(T);
        };
        auto issue = [&](uint32_t i, uint32_t (&pv)[kTlUnroll]) {
#pragma unroll
          for (int u = 0; u < kTlUnroll; ++u) {
            const uint2 ent = *reinterpret_cast<const uint2 *>(wt + i + u);
            // lanes past the end of the slice re-read its last entry (same region as their left
            // neighbour: no mark)
            pv[u] = __ldg(positions + min(ent.x + lane, ent.y));
          }
        };

        open_tile(0);
        if (n_r) issue(0, pva);
        for (uint32_t T = 0; T < nT; ++T) {
          // ---- bitmap of tile T: the top hc words of tile T-1 become the bottom ones, the rest is cleared
          {
            uint4 *o4 = reinterpret_cast<uint4 *>(occ);
            if (T == 0) {
              for (uint32_t i = tid; i < occ_words / 4; i += kMarkers) o4[i] = make_uint4(0, 0, 0, 0);
            } else {
              const uint32_t h4 = (hc + 3u) & ~3u;                  // nw % 4 == 0
              uint4 *ptr = o4 + h4 / 4 + tid;
              const uint32_t n4 = nw / 4 - h4 / 4, full = n4 / kMarkers;
#pragma unroll 4
              for (uint32_t k = 0; k < full; ++k) ptr[k * kMarkers] = make_uint4(0, 0, 0, 0);
              if (full * kMarkers + tid < n4) ptr[full * kMarkers] = make_uint4(0, 0, 0, 0);
              if (tid < h4) {
                if (tid < hc) {
                  const uint32_t v = occ[nw + tid];
                  occ[nw + tid] = 0;
                  occ[tid] = v;
                } else {
                  occ[tid] = 0;
                }
              }
            }
          }
          tl_bar_markers(kMarkers);   // A: bitmap ready

          const uint32_t wbase = T * nw - hc;                     // absolute emit word of bitmap word 0
          const uint32_t l_one = 1u - 31u * wbase;                // local region of absolute region 1
          // events of one step: up to two emitted regions per mark: l (a second list, or the right
          // neighbour is occupied) and l - 1 (the left neighbour is occupied; also the virtual region 0
          // of aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an unoccupied
          // region 0 still emits when region 1 alone reaches the threshold)
          auto events = [&](uint32_t l, uint32_t old, uint32_t old2) {
            const uint32_t qw = tl_div31(l), b = l - qw * 31u, bit = 1u << b;
            const uint32_t self = old & (3u << b);
            const uint32_t left = (old & (bit >> 1)) | (old2 & 0x40000000u);
            if (__any_sync(kFull, (self | left) != 0)) {
              const bool lo = left != 0 || (l == l_one && (old & bit) != 0);
              if (lo) {
                const uint32_t w = wbase + qw - (b == 0 ? 1u : 0u);
                atomicOr(emap + w, b == 0 ? 0x40000000u : bit >> 1);
                atomicOr(sring + ((w >> 7) & rs_mask), 1u << ((w >> 2) & 31u));
              }
              if (self) {
                const uint32_t w = wbase + qw;
                atomicOr(emap + w, bit);
                atomicOr(sring + ((w >> 7) & rs_mask), 1u << ((w >> 2) & 31u));
              }
            }
          };
          auto region_of = [&](uint32_t i, uint32_t pv) {
            const uint2 cc = *reinterpret_cast<const uint2 *>(&wt[i].z);
            return (pv - (lane == 0 ? cc.y : cc.x)) >> r;
          };
          // Both atomics are unconditional: a lane without a mark ORs 0 into its own dummy word (32
          // words, 32 banks).  No divergence between the steps of a batch, so their shuffle -> atomic
          // chains overlap.
          auto arrive = [&](uint32_t l, uint32_t &old, uint32_t &old2) {
            const uint32_t lp = __shfl_up_sync(kFull, l, 1);   // lane 0 receives its own l: never a mark
            const bool mark = l != lp;
            const uint32_t qw = tl_div31(l), b = l - qw * 31u;
            const uint32_t wa = occ_s + 4u * qw;
            const bool low = mark && b == 0;
            old = atomicOr_shared(mark ? wa : dummy_s, mark ? 1u << b : 0u);
            old2 = atomicOr_shared(low ? wa - 4u : dummy_s, low ? 0x80000000u : 0u);
            if (!mark) old = 0;
            if (!low) old2 = 0;
          };
          // a full batch: regions, then all atomics back to back, then the (rare) events
          auto process = [&](uint32_t i, const uint32_t (&pv)[kTlUnroll]) {
            uint32_t l[kTlUnroll], old[kTlUnroll], old2[kTlUnroll];
#pragma unroll
            for (int u = 0; u < kTlUnroll; ++u) l[u] = region_of(i + u, pv[u]);
#pragma unroll
            for (int u = 0; u < kTlUnroll; ++u) arrive(l[u], old[u], old2[u]);
#pragma unroll
            for (int u = 0; u < kTlUnroll; ++u) events(l[u], old[u], old2[u]);
          };
          // the last, partial batch of a round: cnt < kTlUnroll steps, one at a time
          auto process_tail = [&](uint32_t i, const uint32_t (&pv)[kTlUnroll], uint32_t cnt) {
#pragma unroll
            for (int u = 0; u < kTlUnroll - 1; ++u) {
              if ((uint32_t)u < cnt) {
                uint32_t old, old2;
                const uint32_t l = region_of(i + u, pv[u]);
                arrive(l, old, old2);
                events(l, old, old2);
              }
            }
          };

          if (n_r) {
            while (true) {
              // batches of kTlUnroll steps, the loads of the next batch in flight while one is processed
              // (the table is padded with null steps, so a partial batch may be loaded as a whole)
              for (uint32_t i = 0;; i += 2 * kTlUnroll) {
                if (i + kTlUnroll < n_r) issue(i + kTlUnroll, pvb);
                if (n_r - i >= (uint32_t)kTlUnroll) process(i, pva); else process_tail(i, pva, n_r - i);
                if (i + kTlUnroll >= n_r) break;
                if (i + 2 * kTlUnroll < n_r) issue(i + 2 * kTlUnroll, pva);
                if (n_r - i >= 2u * kTlUnroll) process(i + kTlUnroll, pvb);
                else process_tail(i + kTlUnroll, pvb, n_r - i - kTlUnroll);
                if (i + 2 * kTlUnroll >= n_r) break;
              }
              s0 += 32;
              if (s0 >= s_hi) break;
              n_r = min(32u, s_hi - s0);
              build_round(T);
              issue(0, pva);
            }
          }
          // the next tile's first round and first batch before the barrier: their latency hides behind
          // the barrier and the clearing of the bitmap
          if (T + 1 < nT) {
            open_tile(T + 1);
            if (n_r) issue(0, pva);
          }
          __syncthreads();   // B: every mark of tile T is in (the scanner may take the final range)
        }
      } else {
        // =============================== scanner warp ===============================
        uint32_t at = 0;            // candidates of this attempt so far
        uint32_t n_f = 0;           // flagged groups in flist
        // flagged groups flist[0, n_f): fetch, clear, append their regions in order
        auto flush = [&]() {
          __syncwarp();
          for (uint32_t i0 = 0; i0 < n_f; i0 += 64) {
            uint4 v[2];
            uint32_t g[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t idx = i0 + 32 * k + lane;
              g[k] = idx < n_f ? flist[idx] : kFull;
              v[k] = make_uint4(0, 0, 0, 0);
              if (g[k] != kFull) v[k] = __ldcg(reinterpret_cast<const uint4 *>(emap) + g[k]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (i0 + 32 * k >= n_f) break;
              if (g[k] != kFull) __stcg(reinterpret_cast<uint4 *>(emap) + g[k], make_uint4(0, 0, 0, 0));
              const uint32_t cnt = __popc(v[k].x) + __popc(v[k].y) + __popc(v[k].z) + __popc(v[k].w);
              uint32_t incl = cnt;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
              }
              uint32_t o = at + incl - cnt;
              if (cnt) {
                const uint32_t wv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                  uint32_t bb = wv[wi];
                  const uint32_t reg0 = 31u * (4u * g[k] + wi);
                  while (bb) {
                    const uint32_t b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    if (o < out_cap) out[o] = (reg0 + b) << r;
                    ++o;
                  }
                }
              }
              at += __shfl_sync(kFull, incl, 31);
            }
          }
          n_f = 0;
          __syncwarp();
        };
        // summary bits of the groups [glo, ghi) -> flist (ascending), bits cleared.  Every lane takes
        // K consecutive summary words, so lane order is group order; K halves when a pass would
        // overflow flist (one word per lane always fits).
        auto scan_range = [&](uint32_t glo, uint32_t ghi) {
          if (glo >= ghi) return;
          uint32_t w = glo >> 5;                         // next summary word (absolute)
          const uint32_t w_end = (ghi + 31u) >> 5;
          uint32_t K = (w_end - w + 31u) / 32u;
          while (w < w_end) {
            K = min(K, (w_end - w + 31u) / 32u);
            const uint32_t w0 = w + lane * K;
            uint32_t cnt = 0;
            for (uint32_t k = 0; k < K; ++k) {
              const uint32_t ww = w0 + k;
              if (ww < w_end) {
                uint32_t sv = sring[ww & rs_mask];
                if ((ww << 5) < glo) sv &= kFull << (glo - (ww << 5));
                if ((ww << 5) + 32u > ghi) sv &= (1u << (ghi - (ww << 5))) - 1u;
                cnt += __popc(sv);
              }
            }
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const uint32_t t = __shfl_up_sync(kFull, incl, o);
              if (lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (total > (uint32_t)kTlFlist && K > 1) { K >>= 1; continue; }
            if (n_f + total > (uint32_t)kTlFlist) flush();
            if (cnt) {
              uint32_t o = n_f + incl - cnt;
              for (uint32_t k = 0; k < K; ++k) {
                const uint32_t ww = w0 + k;
                if (ww < w_end) {
                  uint32_t *sw = sring + (ww & rs_mask);
                  uint32_t sv = *sw;
                  if ((ww << 5) < glo) sv &= kFull << (glo - (ww << 5));
                  if ((ww << 5) + 32u > ghi) sv &= (1u << (ghi - (ww << 5))) - 1u;
                  if (sv) atomicAnd(sw, ~sv);
                  while (sv) {
                    const uint32_t b = __ffs(sv) - 1;
                    sv &= sv - 1;
                    flist[o++] = (ww << 5) + b;
                  }
                }
              }
            }
            n_f += total;
            w += 32u * K;
          }
        };
        uint32_t glo = 0;           // first emit group not scanned yet
        for (uint32_t T = 0; T < nT; ++T) {
          // while the markers work on tile T: the range that tile T-1 made final.  Nothing above
          // word (T' + 1) * nw - hc can change after tile T'.
          if (T) {
            const uint32_t ghi = (T * nw - hc) >> 2;
            scan_range(glo, ghi);
            flush();
            glo = max(glo, ghi);
          }
          __syncthreads();   // B of tile T
        }
        scan_range(glo, (nT * nw + 3u) >> 2);   // the last tile: everything that is left
        flush();
        if (lane == 0) sh.stage_n = at;
      }
      __syncthreads();

      n = sh.stage_n;
      if (direct || n <= out_cap) break;
      // more candidates than the staging area holds: the exact count is known now, write in place
      if (tid == 0) {
        sh.base = atomicAdd(p.cand_cursor, (unsigned long long)n);
        if (sh.base + n > p.cand_capacity) atomicExch(p.overflow, 1);
      }
      __syncthreads();
      if (sh.base + n > p.cand_capacity) break;     // reported through p.overflow
      out = p.cand_start + sh.base;
      out_cap = n;
      direct = true;
    }

    if (direct || n > out_cap) {
      if (tid == 0) {
        p.cand_off[q] = (uint32_t)sh.base;
        p.cand_cnt[q] = direct && sh.base + n <= p.cand_capacity ? n : 0u;
      }
      continue;
    }
    if (tid == 0) {
      sh.base = n ? atomicAdd(p.cand_cursor, (unsigned long long)n) : 0ull;
      const bool fits = sh.base + n <= p.cand_capacity;
      if (n && !fits) atomicExch(p.overflow, 1);
      p.cand_off[q] = (uint32_t)sh.base;
      p.cand_cnt[q] = fits ? n : 0u;
    }
    __syncthreads();
    const unsigned long long cbase = sh.base;
    if (n && cbase + n <= p.cand_capacity)
      for (uint32_t i = tid; i < n; i += kThreads) p.cand_start[cbase + i] = __ldcg(stage + i);
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

struct TileTuning { int nw_warps, minb; };

TileTuning tile_tuning() {
  TileTuning t = {10, 3};
  if (const char *env = getenv("GM_TILE_CFG")) {
    int a = 0, b = 0;
    if (sscanf(env, "%d,%d", &a, &b) == 2) t = {a, b};
  }
  return t;
}

size_t tile_smem_bytes(const TileGeometry &g, uint32_t list_len, int n_warps) {
  const size_t occ_words = (g.nw + g.hc + 3u) & ~3u;
  return (occ_words + g.ring + (size_t)(n_warps - 1) * 32 * 4 + kTlFlist +
          (size_t)list_len * (g.n_tiles + 1) + (size_t)g.n_tiles * (list_len + 1)) * sizeof(uint32_t);
}

template <int NW, int MINB>
cudaError_t tile_launch(const SearchParams &p, size_t smem, int sm_count, int max_grid, cudaStream_t stream) {
  auto kern = seed_search_tile_kernel<NW, MINB>;
  cudaError_t err = allow_max_dynamic_smem(kern);
  if (err != cudaSuccess) return err;
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  int grid = sm_count * per_sm;
  if (grid > max_grid) grid = max_grid;
  kern<<<grid, NW * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

// Geometry of the tiled search for one db chunk and option set; false = not eligible (the other
// kernels of seed_search.cu take over).
bool search_tile_geometry(uint32_t threshold, uint32_t list_len, uint32_t shift, uint32_t log_region,
                          uint32_t seq_len, uint32_t n_keys, size_t smem_per_sm, TileGeometry *g) {
  if (threshold != 2 || list_len > (uint32_t)kTlLists || log_region > 10) return false;
  const TileTuning t = tile_tuning();
  if (t.minb < 1 || t.nw_warps < 2) return false;
  const uint32_t max_off = (list_len - 1) * shift;
  const uint32_t hn = (max_off + (1u << log_region) - 1) >> log_region;
  const uint32_t hc = (hn + 30) / 31 + 1;
  if (hc > 32) return false;
  const uint32_t n_regions = (seq_len >> log_region) + 1;
  const uint32_t words_needed = (n_regions + 30) / 31 + 1;
  // dynamic shared memory one CTA may use: its share of the SM minus the 1 KB the system reserves
  // per CTA and the kernel's static part
  size_t budget = smem_per_sm / (size_t)t.minb;
  if (budget > (size_t)227 * 1024) budget = (size_t)227 * 1024;
  if (budget < 1024 + 1024 + 8192) return false;
  budget -= 1024 + 1024;
  uint32_t nw = (uint32_t)((budget / 4) & ~(size_t)127);
  if (nw > ((words_needed + 127u) & ~127u)) nw = (words_needed + 127u) & ~127u;
  if (const char *env = getenv("GM_TILE_NW")) {   // tests: small tiles, so that small inputs span many of them
    const uint32_t cap = ((uint32_t)atoi(env) + 127u) & ~127u;
    if (cap >= 256 && cap < nw) nw = cap;
  }
  TileGeometry best = {};
  for (; nw >= 256; nw -= 128) {
    const unsigned long long tile_pos = ((unsigned long long)31 * nw) << log_region;
    if (tile_pos >= (1ull << 31)) continue;
    if (tile_pos <= max_off || hc * 4 > nw) return false;
    const uint32_t n_tiles = (uint32_t)(((unsigned long long)seq_len + tile_pos - 1) / tile_pos);
    if (n_tiles > (uint32_t)kTlMaxTiles) return false;
    uint32_t ring = 4;
    while (ring * 128u < 2u * nw + hc + 256u) ring <<= 1;   // summary bits of two tiles + halo + slack
    best.nw = nw;
    best.hc = hc;
    best.n_tiles = n_tiles ? n_tiles : 1;
    best.tile_pos = (uint32_t)tile_pos;
    best.ring = ring;
    if (tile_smem_bytes(best, list_len, t.nw_warps) <= budget) break;
    best.nw = 0;
  }
  if (best.nw == 0) return false;
  if ((unsigned long long)n_keys * best.n_tiles > (96ull << 20)) return false;   // split table <= 384 MiB
  *g = best;
  return true;
}

int search_tile_grid(int sm_count) { return sm_count * 8; }   // upper bound for staging area / emit bitmaps

// words of one CTA's emit bitmap (absolute region space of the chunk, groups of 4 words)
size_t search_tile_emap_words(const TileGeometry &g) { return ((size_t)g.n_tiles * g.nw + 255u) & ~(size_t)127; }

cudaError_t search_split_build(const uint32_t *keys_count, uint32_t n_keys, const uint32_t *positions,
                               const TileGeometry &g, uint32_t *split, int sm_count, cudaStream_t stream) {
  split_build_kernel<<<sm_count * 8, 256, 0, stream>>>(keys_count, n_keys, positions, g.n_tiles,
                                                      g.tile_pos, split);
  return cudaGetLastError();
}

cudaError_t seed_search_tile_launch(SearchParams p, const TileGeometry &g, int sm_count,
                                    cudaStream_t stream) {
  p.tl_nw = g.nw;
  p.tl_hc = g.hc;
  p.tl_tiles = g.n_tiles;
  p.tl_ring = g.ring;
  const TileTuning t = tile_tuning();
  const size_t smem = tile_smem_bytes(g, p.list_len, t.nw_warps);
  const int max_grid = min(search_tile_grid(sm_count), sm_count * t.minb);
#define GM_TILE_CASE(NW, MINB) \
  if (t.nw_warps == NW && t.minb == MINB) return tile_launch<NW, MINB>(p, smem, sm_count, max_grid, stream);
  GM_TILE_CASE(16, 2)
  GM_TILE_CASE(12, 2)
  GM_TILE_CASE(20, 2)
  GM_TILE_CASE(10, 3)
  GM_TILE_CASE(9, 3)
  GM_TILE_CASE(11, 3)
  GM_TILE_CASE(7, 4)
  GM_TILE_CASE(9, 4)
  GM_TILE_CASE(8, 4)
  GM_TILE_CASE(6, 5)
  GM_TILE_CASE(6, 6)
  GM_TILE_CASE(8, 5)
  GM_TILE_CASE(12, 3)
  GM_TILE_CASE(8, 2)
  GM_TILE_CASE(8, 3)
  GM_TILE_CASE(6, 3)
  GM_TILE_CASE(6, 4)
  GM_TILE_CASE(4, 6)
  GM_TILE_CASE(5, 5)
  GM_TILE_CASE(24, 1)
  GM_TILE_CASE(32, 1)
#undef GM_TILE_CASE
  return cudaErrorInvalidValue;
}

}  // namespace gm
