// Tiled seed search (threshold 2, list_len <= 64): the default seed-lookup + region-count filter.
//
// Semantics as in seed_search.cu (reference SearchNextCpu, aligner.cpp:418-509): with cnt(d) the
// number of lists (query k-mer offsets j) that have a position p >= j*shift in region
// d = (p - j*shift) >> r, every occupied region - and the virtual region 0, aligner.cpp:451 - emits
// the candidate d << r iff cnt(d) + cnt(d+1) >= 2, candidates ascending per query.
//
// What is different here is the data layout the kernel reads and the fact that every index position
// is touched exactly ONCE, by about one warp instruction:
//
//   * the db chunk's position space is cut into tiles of Wd = 31 * nw * 2^r positions and the index
//     carries, next to keys_count/positions (index.h:105-114), a SPLIT table
//     split[key * n_tiles + T] = first entry of key's list that is >= T * Wd (built once per chunk on
//     the device).  The slice of list j that falls into tile T is then two table reads - the "top
//     levels" of the search tree over each k-mer's interval live in this table - and a query's 36
//     rows of it are staged in shared memory once per query;
//   * a CTA owns one query and walks the tiles in ascending order with ONE occupancy bitmap of the
//     tile in shared memory.  A bitmap word holds 31 regions plus, in bit 31, a copy of the first
//     region of the next word, so ANY two adjacent regions share a word.  A position that starts a
//     (list, region) run - a MARK - does one atomicOr on its word (a second one on the previous
//     word when it sits in bit 0) and looks at the value the atomic returns: its own bit already
//     set = a second list in the region; the bit above = right neighbour occupied; the bit below =
//     left neighbour occupied.  Atomics on one word are totally ordered, so of two marks that make a
//     region emit, the LATER one always sees the earlier one: every emission is detected exactly
//     where it happens, in one pass, without a second bitmap, without barriers between "arrive" and
//     "decide", and without fences (the store-buffer pattern of two different words never arises);
//   * lists of different tiles meet in the hc topmost words of a tile (a position p >= T * Wd of
//     list j lies up to j*shift positions below its tile in region space): those words are carried
//     into the next tile's bitmap instead of being cleared, as earlier arrivals;
//   * detected emissions are rare (~1.5 % of the marks).  They go to small range buckets (monotone
//     in the region, a ring of two tiles), duplicates are dropped on insertion where visible and for
//     good when one warp sorts each bucket (<= 8 entries, a register sorting network) and appends
//     the tile's buckets in order to the query's staging area.  One atomicAdd on the global cursor
//     per query; every query's candidates are contiguous and ascending as before.
//
// Streaming: a tile's slices are cut into steps of 31 positions (+1 predecessor in lane 0, so "first
// of its list in the region" is one shuffle and one compare) listed in a shared-memory table, dealt
// round robin to the warps, kTlUnroll loads in flight per warp.  Queries that exceed a fixed
// capacity (a bucket, the staging area) are queued for the sweep kernel of seed_search.cu.
#include "gm_common.cuh"

#include <stdlib.h>

namespace gm {

namespace {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kTlLists = 64;
constexpr int kTlMaxTiles = 64;
constexpr int kTlBucketCap = 8;          // entries per range bucket
constexpr int kTlMaxBuckets = 64;        // range buckets per tile
constexpr int kTlTabCap = 256;           // steps per table round
constexpr int kTlUnroll = 4;
constexpr uint32_t kTlMagic31 = 138547333u;   // ceil(2^32 / 31): x / 31 == umulhi(x, magic) for x < 2^27
constexpr uint32_t kTlNone = 0xFFFFFFFFu;

struct TileShared {
  uint32_t lbeg[kTlLists];               // first usable entry of list j (aligner.cpp:430-431)
  uint32_t query, stage_n, bad, n_steps, weight;
  uint32_t wsum[32];                     // dense mode: per-warp counts of the ordered scan
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

// old = atomicOr on a shared-memory word when `on`, else 0.
__device__ __forceinline__ uint32_t tl_atoms_or(uint32_t addr, uint32_t v, bool on) {
  uint32_t old = 0;
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q atom.shared.or.b32 %0, [%1], %2;\n\t}"
      : "+r"(old)
      : "r"(addr), "r"(v), "r"((uint32_t)on)
      : "memory");
  return old;
}

__device__ __forceinline__ void tl_cswap(uint32_t &a, uint32_t &b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

// Emitted global region g -> range bucket `e` (kTlBucketCap entries): a small concurrent set (first
// free slot wins, an equal entry ends the probe), so a region is stored once however many marks
// report it.  Returns false when the bucket is full.
__device__ __noinline__ bool tl_emit(uint32_t *e, uint32_t g) {
  for (int k = 0; k < kTlBucketCap; ++k) {
    const uint32_t old = atomicCAS(e + k, kTlNone, g);
    if (old == kTlNone || old == g) return true;
  }
  return false;
}

// Dense mode: emitted tile-local region x -> bit of the CTA's emit bitmap (global memory, L2).
__device__ __noinline__ void tl_emit_dense(uint32_t *emap, uint32_t x) {
  const uint32_t w = tl_div31(x);
  atomicOr(emap + w, 1u << (x - 31u * w));
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

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) seed_search_tile_kernel(const SearchParams p) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ TileShared sh;
  constexpr uint32_t kThreads = NW * 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const uint32_t warp = __shfl_sync(kFull, tid >> 5, 0);   // provably warp-uniform for the compiler
  const uint32_t r = p.log_region, nw = p.tl_nw, hc = p.tl_hc, nb = p.tl_nb, nT = p.tl_tiles;
  const uint32_t wpb_log = p.tl_wpb_log;
  const uint32_t tile_pos = (31u * nw) << r;             // Wd
  const uint32_t occ_words = (nw + hc + 3u) & ~3u;
  uint32_t *occ = dyn;                                   // [nw + hc] 31 regions + 1 overlap bit per word
  uint4 *tab = reinterpret_cast<uint4 *>(dyn + occ_words);          // [kTlTabCap] steps of this round
  uint32_t *bent = dyn + occ_words + 4 * kTlTabCap;      // [2 * nb][kTlBucketCap] ring of range buckets
  uint32_t *bounds = bent + 2 * nb * kTlBucketCap;       // [list_len][nT + 1] slices of this query
  uint32_t *stage = p.staging + (size_t)blockIdx.x * p.staging_cap;
  uint32_t *emap = p.tl_emap + (size_t)blockIdx.x * occ_words;   // dense mode: emitted regions of the tile
  uint32_t occ_s = tl_smem_addr(occ);
  asm volatile("mov.u32 %0, %0;" : "+r"(occ_s));   // keep the window address in a register
  const uint32_t *__restrict__ positions = p.positions;

  for (uint32_t i = tid; i < 2 * nb * kTlBucketCap; i += kThreads) bent[i] = kTlNone;
  if (tid == 0) { sh.visited = 0; sh.weight = 0; }
  __syncthreads();

  uint32_t *out = stage;            // where the ordered candidates of the current attempt go
  uint32_t out_cap = p.staging_cap;

  // ---- one warp: sort the first n_b buckets of tile T, append them in order to the staging area
  auto finalize = [&](uint32_t T, uint32_t n_b) {
    uint32_t total = *reinterpret_cast<volatile uint32_t *>(&sh.stage_n);
    const uint32_t ring = (T & 1u) * nb;
    for (uint32_t l0 = 0; l0 < n_b; l0 += 32) {
      const uint32_t lbk = l0 + lane;
      uint32_t v[kTlBucketCap];
#pragma unroll
      for (int i = 0; i < kTlBucketCap; ++i) v[i] = kTlNone;
      if (lbk < n_b) {
        uint32_t slot = ring + lbk;
        if (slot >= 2 * nb) slot -= 2 * nb;
        uint4 *e = reinterpret_cast<uint4 *>(bent + slot * kTlBucketCap);
        const uint4 a = e[0];
        if (a.x != kTlNone) {
          const uint4 b = e[1];
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          e[0] = make_uint4(kTlNone, kTlNone, kTlNone, kTlNone);
          if (b.x != kTlNone) e[1] = make_uint4(kTlNone, kTlNone, kTlNone, kTlNone);
        }
      }
      if (__any_sync(kFull, v[0] != kTlNone)) {
        if (v[1] != kTlNone) {   // 19-comparator sorting network for 8 keys
          tl_cswap(v[0], v[1]); tl_cswap(v[2], v[3]); tl_cswap(v[4], v[5]); tl_cswap(v[6], v[7]);
          tl_cswap(v[0], v[2]); tl_cswap(v[1], v[3]); tl_cswap(v[4], v[6]); tl_cswap(v[5], v[7]);
          tl_cswap(v[1], v[2]); tl_cswap(v[5], v[6]); tl_cswap(v[0], v[4]); tl_cswap(v[3], v[7]);
          tl_cswap(v[1], v[5]); tl_cswap(v[2], v[6]);
          tl_cswap(v[1], v[4]); tl_cswap(v[3], v[6]);
          tl_cswap(v[2], v[4]); tl_cswap(v[3], v[5]);
          tl_cswap(v[3], v[4]);
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < kTlBucketCap; ++i) cnt += v[i] != kTlNone;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += t;
        }
        uint32_t at = total + incl - cnt;
#pragma unroll
        for (int i = 0; i < kTlBucketCap; ++i)
          if ((uint32_t)i < cnt) {          // sorted: the entries come first, kTlNone last
            if (at < out_cap) out[at] = v[i] << r;
            ++at;
          }
        total += __shfl_sync(kFull, incl, 31);
      }
    }
    if (lane == 0) sh.stage_n = total;
  };

  // ---- warp 0: steps [s0, s0 + kTlTabCap) of tile T into the table; total steps -> sh.n_steps
  auto build_table = [&](uint32_t T, uint32_t s0) {
    const uint32_t cb = T * tile_pos - ((31u * hc) << r);   // wraps for T == 0: only differences matter
    uint32_t carry = 0;
#pragma unroll
    for (int jj = 0; jj < kTlLists / 32; ++jj) {
      const uint32_t j = lane + 32 * jj;
      uint32_t b = 0, e = 0;
      if (j < p.list_len) {
        b = bounds[j * (nT + 1) + T];
        e = bounds[j * (nT + 1) + T + 1];
      }
      const uint32_t st = tl_div31(e - b + 30u);
      uint32_t incl = st;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t pre = carry + incl - st;
      carry += __shfl_sync(kFull, incl, 31);
      if (st && pre < s0 + kTlTabCap && pre + st > s0) {
        const uint32_t k_lo = s0 > pre ? s0 - pre : 0u;
        const uint32_t k_hi = min(st, s0 + kTlTabCap - pre);
        const uint32_t c = cb + j * p.shift, lb = sh.lbeg[j];
        // lane i of step k looks at entry lb + (b - 1 - lb + 31 k) + i of the list: lane 0 is the
        // predecessor of the step's first position (or nothing: the difference wraps to ~0)
        for (uint32_t k = k_lo; k < k_hi; ++k)
          tab[pre + k - s0] = make_uint4(b - 1u - lb + 31u * k, e - lb, c, lb);
      }
    }
    if (lane == 0) sh.n_steps = carry;
  };

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    const uint32_t q = sh.query;
    if (q >= p.n_queries) break;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    // ---- phase 0: the query's rows of the split table; leading positions < j*shift dropped
    if (tid < p.list_len) {
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
      if (bj[nT] > b) {
        atomicAdd(&sh.visited, (unsigned long long)(bj[nT] - b));
        atomicAdd(&sh.weight, bj[nT] - b);
      }
    }
    __syncthreads();

    // Attempt 0 is the sparse path (emissions in range buckets).  A query that overflows a bucket
    // or the staging area is redone in DENSE mode: emissions become bits of a per-CTA bitmap in
    // global memory that every tile scans in order - no capacity anywhere; if even the staging
    // area is too small, a last attempt writes straight into the query's slice of the output.
    out = stage;
    out_cap = p.staging_cap;
    // W marks in R regions emit about 1.5 W^2 / R regions, i.e. 1.5 * 31 * 2^wpb_log * (W / R)^2 per
    // bucket: beyond ~1.5 per bucket an overflow is likely, so such a query starts in dense mode
    const float wr = (float)sh.weight / (float)p.n_regions;
    bool dense = p.tl_force_dense != 0 || wr * wr * (float)(31u << wpb_log) > 1.0f, direct = false;
    __syncthreads();
    if (tid == 0) sh.weight = 0;
    uint32_t n = 0;
    while (true) {
      if (tid == 0) { sh.stage_n = 0; sh.bad = 0; }
      __syncthreads();

      for (uint32_t T = 0; T < nT; ++T) {
        // ---- prologue: table of tile T (warp 0), buckets of tile T-1 (warp 1), bitmap (the rest)
        if (warp == 0) {
          build_table(T, 0);
        } else if (warp == 1) {
          if (T && !dense) finalize(T - 1, nb);
        } else {
          const uint32_t ct = tid - 64, cn = kThreads - 64;
          uint4 *o4 = reinterpret_cast<uint4 *>(occ);
          if (T == 0) {
#pragma unroll 4
            for (uint32_t i = ct; i < occ_words / 4; i += cn) o4[i] = make_uint4(0, 0, 0, 0);
          } else {
            const uint32_t h4 = (hc + 3u) & ~3u;                  // nw % 4 == 0
#pragma unroll 4
            for (uint32_t i = h4 / 4 + ct; i < nw / 4; i += cn) o4[i] = make_uint4(0, 0, 0, 0);
            if (ct < h4) {
              if (ct < hc) {          // carry: the top hc words become the bottom ones
                const uint32_t v = occ[nw + ct];
                occ[nw + ct] = 0;
                occ[ct] = v;
                if (dense) {
                  const uint32_t ev = __ldcg(emap + nw + ct);
                  emap[nw + ct] = 0;
                  emap[ct] = ev;
                }
              } else {
                occ[ct] = 0;
              }
            }
          }
        }
        const uint32_t gbase = T * 31u * nw - 31u * hc;          // global region of local region 0
        const uint32_t ring = (T & 1u) * nb;
        for (uint32_t s0 = 0;; s0 += kTlTabCap) {
          if (s0) {
            __syncthreads();
            if (warp == 0) build_table(T, s0);
          }
          __syncthreads();   // A
          const uint32_t S = sh.n_steps;
          const uint32_t n_round = min(S - s0, (uint32_t)kTlTabCap);
          if (s0 == 0 && warp == NW - 1 && T + 1 < nT) {
            // the next tile's slices into L2 while this one is processed (one request per 128 B line)
            for (uint32_t j = lane; j < p.list_len; j += 32) {
              const uint32_t pb = bounds[j * (nT + 1) + T + 1] & ~31u, pe = bounds[j * (nT + 1) + T + 2];
              for (uint32_t a = pb; a < pe; a += 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(positions + a) : "memory");
            }
          }
          // contiguous share of the round's steps for every warp, kTlUnroll loads in flight
          const uint32_t per = (n_round + NW - 1) / NW;
          const uint32_t i1 = min(warp * per + per, n_round);
          for (uint32_t i = warp * per; i < i1; i += kTlUnroll) {
            const uint32_t cnt = i1 - i;
            uint32_t pv[kTlUnroll], cj[kTlUnroll];
#pragma unroll
            for (int u = 0; u < kTlUnroll; ++u) {
              if (u == 0 || (uint32_t)u < cnt) {
                const uint4 ent = tab[i + u];
                // lanes past the end re-read the last entry (same region as their left neighbour: no
                // mark); lane 0 without a predecessor gets a region no position can have
                const uint32_t idx = ent.w + min(ent.x + lane, ent.y - 1u);
                cj[u] = (lane == 0 && ent.x == kTlNone) ? ent.z ^ 0x80000000u : ent.z;
                pv[u] = __ldg(positions + idx);
              }
            }
#pragma unroll
            for (int u = 0; u < kTlUnroll; ++u) {
              if (u == 0 || (uint32_t)u < cnt) {
                const uint32_t l = (pv[u] - cj[u]) >> r;
                const uint32_t lp = __shfl_up_sync(kFull, l, 1);   // lane 0 receives its own l: never a mark
                const bool mark = l != lp;
                const uint32_t qw = tl_div31(l), b = l - qw * 31u, bit = 1u << b;
                const uint32_t wa = occ_s + 4u * qw;
                const uint32_t old = tl_atoms_or(wa, bit, mark);
                const uint32_t old2 = tl_atoms_or(wa - 4u, 0x80000000u, mark && b == 0);
                const uint32_t self = old & (3u << b);
                const uint32_t left = (old & (bit >> 1)) | (old2 & 0x40000000u);
                if (__any_sync(kFull, (self | left) != 0)) {
                  // up to two emitted regions per mark: l (a second list, or the right neighbour is
                  // occupied) and l - 1 (the left neighbour is occupied; also the virtual region 0 of
                  // aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an
                  // unoccupied region 0 still emits when region 1 alone reaches the threshold)
                  const bool lo = left != 0 || (gbase + l == 1u && (old & bit) != 0);
                  if (dense) {
                    if (lo) tl_emit_dense(emap, l - 1u);
                    if (self) tl_emit_dense(emap, l);
                  } else {
                    if (lo) {
                      const uint32_t x = l - 1u;
                      uint32_t slot = ring + (tl_div31(x) >> wpb_log);
                      if (slot >= 2 * nb) slot -= 2 * nb;
                      if (!tl_emit(bent + slot * kTlBucketCap, gbase + x)) sh.bad = 1;
                    }
                    if (self) {
                      uint32_t slot = ring + (qw >> wpb_log);
                      if (slot >= 2 * nb) slot -= 2 * nb;
                      if (!tl_emit(bent + slot * kTlBucketCap, gbase + l)) sh.bad = 1;
                    }
                  }
                }
              }
            }
          }
          if (s0 + kTlTabCap >= S) break;
        }
        __syncthreads();   // B
        if (dense) {
          // ordered scan of the tile's emit bitmap: everything below the top hc words is final
          // (the top words travel to the next tile with the carry); the last tile scans them too
          const uint32_t n_words = T + 1 < nT ? nw : nw + hc;
          const uint32_t wpw = (((n_words + NW - 1) / NW) + 127u) & ~127u;   // words per warp
          const uint32_t wb = min(warp * wpw, n_words), we = min(wb + wpw, n_words);
          uint32_t c = 0;
#pragma unroll 8
          for (uint32_t w = wb + lane; w < we; w += 32) c += __popc(__ldcg(emap + w));
          c = __reduce_add_sync(kFull, c);
          if (lane == 0) sh.wsum[warp] = c;
          __syncthreads();
          uint32_t at = sh.stage_n, total = 0;
          for (uint32_t w = 0; w < NW; ++w) {
            const uint32_t t = sh.wsum[w];
            if (w < warp) at += t;
            total += t;
          }
          if (c) {
            for (uint32_t w0 = wb; w0 < we; w0 += 128) {   // 4 coalesced rows of 32 words in flight
              uint32_t bits[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t w = w0 + 32 * k + lane;
                bits[k] = w < we ? __ldcg(emap + w) : 0u;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (!__any_sync(kFull, bits[k] != 0)) continue;
                const uint32_t w = w0 + 32 * k + lane;
                const uint32_t cnt = __popc(bits[k]);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const uint32_t t = __shfl_up_sync(kFull, incl, o);
                  if (lane >= o) incl += t;
                }
                uint32_t o = at + incl - cnt;
                if (bits[k]) {
                  emap[w] = 0;
                  uint32_t bb = bits[k];
                  while (bb) {
                    const uint32_t b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    if (o < out_cap) out[o] = (gbase + 31u * w + b) << r;
                    ++o;
                  }
                }
                at += __shfl_sync(kFull, incl, 31);
              }
            }
          }
          __syncthreads();
          if (tid == 0) sh.stage_n += total;
        }
      }
      if (warp == 1 && !dense) finalize(nT - 1, nb + 1);
      __syncthreads();

      n = sh.stage_n;
      if (!dense) {
        if (sh.bad == 0 && n <= out_cap) break;
        dense = true;     // redo: the bucket ring goes back to empty first
        __syncthreads();
        for (uint32_t i = tid; i < 2 * nb * kTlBucketCap; i += kThreads) bent[i] = kTlNone;
        continue;
      }
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
      for (uint32_t i = tid; i < n; i += kThreads) p.cand_start[cbase + i] = stage[i];
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

struct TileTuning { int nw_warps, nb, wpb_log, minb; };

TileTuning tile_tuning() {
  TileTuning t = {8, 32, 8, 5};
  if (const char *env = getenv("GM_TILE_CFG")) {
    int a = 0, b = 0, c = 0, d = 0;
    if (sscanf(env, "%d,%d,%d,%d", &a, &b, &c, &d) == 4) t = {a, b, c, d};
  }
  if (t.nb < 1) t.nb = 1;
  if (t.nb > kTlMaxBuckets) t.nb = kTlMaxBuckets;
  if (t.wpb_log < 2) t.wpb_log = 2;
  if (t.wpb_log > 10) t.wpb_log = 10;
  return t;
}

size_t tile_smem_bytes(const TileGeometry &g, uint32_t list_len) {
  const size_t occ_words = (g.nw + g.hc + 3u) & ~3u;
  return (occ_words + 4 * kTlTabCap + 2 * g.nb * kTlBucketCap +
          (size_t)list_len * (g.n_tiles + 1)) * sizeof(uint32_t);
}

template <int NW, int MINB>
cudaError_t tile_launch(const SearchParams &p, size_t smem, int sm_count, int max_grid, cudaStream_t stream) {
  auto kern = seed_search_tile_kernel<NW, MINB>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
                          uint32_t seq_len, uint32_t n_keys, TileGeometry *g) {
  if (threshold != 2 || list_len > (uint32_t)kTlLists || log_region > 10) return false;
  const TileTuning t = tile_tuning();
  const uint32_t max_off = (list_len - 1) * shift;
  const uint32_t hn = (max_off + (1u << log_region) - 1) >> log_region;
  const uint32_t hc = (hn + 30) / 31 + 1;
  const uint32_t wpb = 1u << t.wpb_log;
  if (hc > 32 || hc > wpb) return false;
  const uint32_t n_regions = (seq_len >> log_region) + 1;
  const uint32_t words_needed = (n_regions + 30) / 31 + 1;
  uint32_t nb = (words_needed + wpb - 1) / wpb;
  if (nb > (uint32_t)t.nb) nb = t.nb;
  if (nb < 1) nb = 1;
  const uint32_t nw = nb * wpb;
  const unsigned long long tile_pos = ((unsigned long long)31 * nw) << log_region;
  if (tile_pos >= (1ull << 31) || tile_pos <= max_off) return false;
  const uint32_t n_tiles = (uint32_t)(((unsigned long long)seq_len + tile_pos - 1) / tile_pos);
  if (n_tiles > (uint32_t)kTlMaxTiles) return false;
  if ((unsigned long long)n_keys * (n_tiles ? n_tiles : 1) > (96ull << 20)) return false;   // split table <= 384 MiB
  g->nw = nw;
  g->hc = hc;
  g->nb = nb;
  g->wpb_log = t.wpb_log;
  g->n_tiles = n_tiles ? n_tiles : 1;
  g->tile_pos = (uint32_t)tile_pos;
  return true;
}

int search_tile_grid(int sm_count) { return sm_count * 6; }   // upper bound for the staging area

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
  p.tl_nb = g.nb;
  p.tl_wpb_log = g.wpb_log;
  p.tl_tiles = g.n_tiles;
  const size_t smem = tile_smem_bytes(g, p.list_len);
  const TileTuning t = tile_tuning();
  const int max_grid = search_tile_grid(sm_count);
#define GM_TILE_CASE(NW, MINB) \
  if (t.nw_warps == NW && t.minb == MINB) return tile_launch<NW, MINB>(p, smem, sm_count, max_grid, stream);
  GM_TILE_CASE(12, 3)
  GM_TILE_CASE(8, 4)
  GM_TILE_CASE(8, 5)
  GM_TILE_CASE(8, 3)
  GM_TILE_CASE(16, 2)
  GM_TILE_CASE(12, 2)
  GM_TILE_CASE(6, 6)
  GM_TILE_CASE(4, 8)
#undef GM_TILE_CASE
  return cudaErrorInvalidValue;
}

}  // namespace gm
