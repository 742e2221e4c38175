// Seed lookup + region-count filter + ordered candidate compaction, sm_100a.
//
// Semantics: reference SearchNextCpu, aligner.cpp:418-509 (GPU twins aligner_gpu.cu:86-367):
// per query, the k-mers at offsets j*shift (j < list_len) are looked up in the CSR index
// (index.h:86-114); every position p >= j*shift of list j marks region d = (p - j*shift) >> r,
// each list at most once per region; with cnt(d) = number of lists marking d, every occupied
// region d (and the virtual region 0, aligner.cpp:451) whose cnt(d) + cnt(d+1) >= threshold
// emits the candidate d << r.  Candidates come out ascending per query.
//
// The reference walks this as a sequential k-way merge (twice on its GPU path).  Here the
// histogram is built directly: a CTA owns one query at a time and sweeps the region space in
// tiles held in shared memory as `threshold` thermometer BIT-PLANES (plane i set <=> cnt > i,
// filled with atomicOr), so a tile covers ~0.5 M regions in <200 KB and a 128 MiB db chunk is
// ~16 tiles.  Each list is sorted, so its slice for a tile is a contiguous run found by a
// cursor; the runs are read with coalesced warp-wide loads (one warp per list, UNROLL loads in
// flight).  Three sparse passes per tile touch only what the positions touch:
//   pass 1 arrive   - first position of every (list, region) run sets the lowest clear plane;
//   pass 2 decide   - the same positions evaluate cnt(d)+cnt(d+1) >= t from the planes and set
//                     the emit bitmap (idempotent, no election needed);
//   pass 3 clear    - the same positions zero the words they touched and commit the cursors.
// Ordered compaction: one thread per 1024-region group counts the emit bits it owns, a block
// scan gives its output slot.  A query's candidates are staged per CTA, then appended to the
// global candidate buffer in ONE allocation, so every query's candidates are contiguous and in
// reference order; (cand_off, cand_cnt) locate them.
#include "gm_common.cuh"

namespace gm {

namespace {

constexpr int kSearchThreads = 1024;
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kUnroll = 4;            // warp-wide loads in flight per list visit
constexpr int kMaxListLen = 1024;     // (L - seed_len)/shift + 1 for L <= 1024
constexpr uint32_t kFull = 0xFFFFFFFFu;

struct SearchShared {
  uint32_t cur[kMaxListLen];
  uint32_t endp[kMaxListLen];
  uint32_t scan[kSearchWarps];
  uint32_t min_d, max_d;
  uint32_t query;
  uint32_t out_n;
  unsigned long long base;
  unsigned long long visited;
};

__device__ __forceinline__ uint32_t get_key(const uint8_t *s, uint32_t seed) {  // index.h:86-101
  uint32_t key = 0;
  for (uint32_t i = 0; seed != 0; ++i, seed >>= 1)
    if (seed & 1) key = (key << kCharBits) | s[i];
  return key;
}

// cnt(x) >= a  <=>  plane a-1 has bit x
template <int T>
__device__ __forceinline__ bool emits(const uint32_t *planes, uint32_t words, uint32_t x) {
  const uint32_t w0 = x >> 5, b0 = 1u << (x & 31), w1 = (x + 1) >> 5, b1 = 1u << ((x + 1) & 31);
  bool r = (planes[(T - 1) * words + w0] & b0) != 0;  // cnt(x) >= T on its own
#pragma unroll
  for (int a = 1; a < T; ++a)
    r |= (planes[(a - 1) * words + w0] & b0) && (planes[(T - a - 1) * words + w1] & b1);
  return r;
}

// the same for a threshold only known at run time (thresholds above 4: generic kernel, T = 0)
__device__ __forceinline__ bool emits_rt(const uint32_t *planes, uint32_t words, uint32_t x, int T) {
  const uint32_t w0 = x >> 5, b0 = 1u << (x & 31), w1 = (x + 1) >> 5, b1 = 1u << ((x + 1) & 31);
  bool r = (planes[(T - 1) * words + w0] & b0) != 0;
  for (int a = 1; a < T && !r; ++a)
    r = (planes[(a - 1) * words + w0] & b0) && (planes[(T - a - 1) * words + w1] & b1);
  return r;
}

template <int T>
__global__ void __launch_bounds__(kSearchThreads, 1) seed_search_kernel(const SearchParams p) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ SearchShared sh;
  const int TT = T ? T : (int)p.threshold;     // T == 0: threshold known at run time only (> 4)
  const uint32_t M = p.tile_regions;
  const uint32_t words = M / 32 + 1;          // +1: halo word for region base+M
  uint32_t *planes = dyn;                     // [T][words]
  uint32_t *emitb = dyn + TT * words;          // [words]
  uint32_t *summary = emitb + words;          // [M/1024/32 + 1] one bit per 32-word group
  const uint32_t groups = M / 1024;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *staging = p.staging + (size_t)blockIdx.x * p.staging_cap;

  for (uint32_t i = tid; i < (TT + 1) * words + groups / 32 + 1; i += kSearchThreads) dyn[i] = 0;
  if (tid == 0) sh.visited = 0;
  __syncthreads();

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    if (sh.query >= p.n_queries) break;
    const uint32_t q = p.query_list ? p.query_list[sh.query] : sh.query;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    uint32_t *out = staging;
    uint32_t out_cap = p.staging_cap;
    for (int attempt = 0; attempt < 2; ++attempt) {
      // ---- lookup: interval of every query k-mer, leading positions < j*shift skipped
      if (tid == 0) { sh.min_d = 0xFFFFFFFFu; sh.max_d = 0; sh.out_n = 0; }
      __syncthreads();
      unsigned long long my_visited = 0;
      for (uint32_t j = tid; j < p.list_len; j += kSearchThreads) {
        const uint32_t key = get_key(query + j * p.shift, p.seed);
        uint32_t b = p.keys_count[key];
        const uint32_t e = p.keys_count[key + 1];                       // index.h:105-114
        const uint32_t off = j * p.shift;
        while (b < e && p.positions[b] < off) ++b;                      // aligner.cpp:430-431
        sh.cur[j] = b;
        sh.endp[j] = e;
        if (b < e) {
          atomicMin(&sh.min_d, (p.positions[b] - off) >> p.log_region);
          atomicMax(&sh.max_d, (p.positions[e - 1] - off) >> p.log_region);
          my_visited += e - b;
        }
      }
      __syncthreads();
      if (attempt == 0 && my_visited) atomicAdd(&sh.visited, my_visited);
      const uint32_t min_d = sh.min_d, max_d = sh.max_d;

      if (min_d != 0xFFFFFFFFu) {
        for (uint32_t tile = min_d / M; tile <= max_d / M; ++tile) {
          const uint32_t base = tile * M;
          // ---------------- passes over the tile's slice of every list
          for (int pass = 1; pass <= 3; ++pass) {
            for (uint32_t j = warp; j < p.list_len; j += kSearchWarps) {
              uint32_t c = sh.cur[j];
              const uint32_t e = sh.endp[j], off = j * p.shift;
              uint32_t carry = 0xFFFFFFFFu;  // region of the previous position of this list
              while (c < e) {
                uint32_t pos[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                  const uint32_t idx = c + u * 32 + lane;
                  pos[u] = idx < e ? __ldg(p.positions + idx) : 0xFFFFFFFFu;
                }
                bool done = false;
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                  const bool valid = pos[u] != 0xFFFFFFFFu;
                  const uint32_t d = valid ? (pos[u] - off) >> p.log_region : 0xFFFFFFFFu;
                  uint32_t dprev = __shfl_up_sync(kFull, d, 1);
                  if (lane == 0) dprev = carry;
                  const bool in_tile = valid && d < base + M;
                  // run start inside the tile, or the first position of region base+M (halo)
                  const bool mark = valid && d <= base + M && d != dprev;
                  if (!done && mark) {
                    const uint32_t x = d - base, w = x >> 5, bit = 1u << (x & 31);
                    if (pass == 1) {
                      uint32_t old = atomicOr(&planes[w], bit);
#pragma unroll
                      for (int t = 1; t < TT; ++t)
                        if (old & bit) old = atomicOr(&planes[t * words + w], bit); else break;
                    } else if (pass == 2) {
                      if (in_tile && (T ? emits<T ? T : 1>(planes, words, x) : emits_rt(planes, words, x, TT))) {
                        atomicOr(&emitb[w], bit);
                        atomicOr(&summary[w >> 10], 1u << ((w >> 5) & 31));
                      }
                    } else {
#pragma unroll
                      for (int t = 0; t < TT; ++t) planes[t * words + w] = 0;
                    }
                  }
                  const uint32_t nin = __popc(__ballot_sync(kFull, in_tile));
                  carry = __shfl_sync(kFull, d, 31);
                  if (!done) {
                    c += nin;
                    if (nin < 32) done = true;
                  }
                }
                if (done) break;
              }
              if (pass == 3 && lane == 0) sh.cur[j] = c;
            }
            __syncthreads();
            if (pass == 1 && tile == 0 && tid == 0) {
              // aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an
              // unoccupied region 0 still emits when region 1 alone reaches the threshold.
              if (!(planes[0] & 1u) && (planes[(TT - 1) * words] & 2u)) {
                emitb[0] |= 1u;
                summary[0] |= 1u;
              }
            }
            if (pass == 1) __syncthreads();
          }
          // ---------------- ordered compaction of the emit bitmap
          uint32_t mine = 0;
          const bool owner = tid < groups && (summary[tid >> 5] >> (tid & 31) & 1u);
          if (owner)
            for (uint32_t w = 0; w < 32; ++w) mine += __popc(emitb[tid * 32 + w]);
          uint32_t incl = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
          }
          if (lane == 31) sh.scan[warp] = incl;
          __syncthreads();
          uint32_t before = sh.out_n;
          for (uint32_t w = 0; w < warp; ++w) before += sh.scan[w];
          uint32_t slot = before + incl - mine;
          if (owner) {
            for (uint32_t w = 0; w < 32; ++w) {
              uint32_t bits = emitb[tid * 32 + w];
              emitb[tid * 32 + w] = 0;
              while (bits) {
                const uint32_t b = __ffs(bits) - 1;
                bits &= bits - 1;
                if (slot < out_cap) out[slot] = (base + tid * 32 * 32 + w * 32 + b) << p.log_region;
                ++slot;
              }
            }
          }
          __syncthreads();
          if (tid == 0) {
            uint32_t total = sh.out_n;
            for (uint32_t w = 0; w < kSearchWarps; ++w) total += sh.scan[w];
            sh.out_n = total;
          }
          if (tid < groups / 32 + 1) summary[tid] = 0;
          if (tid == 0) planes[words - 1] = 0;   // halo word of plane 0 .. T-1
          for (int t = tid; t < TT; t += kSearchThreads) planes[t * words + words - 1] = 0;
          __syncthreads();
        }
      }
      // ---- hand the query's candidates over
      const uint32_t n = sh.out_n;
      if (attempt == 0) {
        if (tid == 0) {
          sh.base = n ? atomicAdd(p.cand_cursor, (unsigned long long)n) : 0ull;
          if (n && sh.base + n > p.cand_capacity) atomicExch(p.overflow, 1);
        }
        __syncthreads();
        const unsigned long long base = sh.base;
        const bool fits = base + n <= p.cand_capacity;
        if (tid == 0) {
          p.cand_off[q] = (uint32_t)base;
          p.cand_cnt[q] = fits ? n : 0u;
        }
        if (!fits || n == 0) break;
        if (n <= p.staging_cap) {
          for (uint32_t i = tid; i < n; i += kSearchThreads) p.cand_start[base + i] = staging[i];
          break;
        }
        // more candidates than the staging area holds: redo the query writing in place
        out = p.cand_start + base;
        out_cap = n;
        __syncthreads();
      }
    }
    __syncthreads();
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

// ---------------------------------------------------------------------------------------------
// Fast path (list_len <= 64): same bit-plane tiles and the same three passes, reorganised so
// that the work is balanced and the passes run out of registers.
//   * slice boundaries: for a batch of kTileBatch tiles, thread (list j, boundary t) finds by
//     binary search the first index of list j whose region is >= the tile base (the lists are
//     sorted, so every tile's slice of every list is a contiguous index range);
//   * balanced chunks: a tile's work is the concatenation of its slices cut into 32-position
//     chunks; chunk k goes to warp k % 32, whichever list it belongs to, so skewed k-mer
//     frequencies (one list with thousands of positions, others with a handful) do not leave
//     warps idle at the barriers;
//   * a warp keeps the mark index of its (up to kSlots) chunks in registers across the three
//     passes and loads the next tile's chunks right after computing the current marks, so the
//     HBM latency of the next tile hides behind the passes of the current one;
//   * per tile 3 block barriers; ordered compaction of the sparse emit bitmap by one warp.
constexpr int kTileBatch = 16;
constexpr int kFastThreads = 512;        // two CTAs per SM: one CTA's barrier/latency phases overlap
constexpr int kFastWarps = kFastThreads / 32;   // with the other CTA's issue slots
constexpr int kSlots = 8;                // register-resident chunks per warp per tile
constexpr int kFastLists = 64;
constexpr int kChunk = 31;               // new positions per chunk (lane 0 carries the predecessor)
constexpr int kChunkTable = kSlots * kFastWarps;   // chunks per tile with a list-id table entry
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct FastShared {
  uint32_t lbeg[kFastLists], lend[kFastLists];
  uint32_t bnd[kTileBatch + 1][kFastLists];       // first index with region >= base of tile t
  uint32_t fr[kTileBatch + 1][kFastLists];        // region of that position (kNone: none left)
  uint32_t pre[kTileBatch][kFastLists + 1];       // exclusive prefix of chunks per list
  uint8_t chunk_list[kTileBatch][kChunkTable];    // chunk k -> list (k < kChunkTable)
  uint32_t scan[32];
  uint32_t min_d, max_d;
  uint32_t query, out_n;
  uint32_t r1_double;   // one-pass variant: region 1 was hit twice (virtual region 0 rule)
  unsigned long long base;
  unsigned long long visited;
};

// list of chunk k of tile t
__device__ __forceinline__ uint32_t chunk_to_list(const SearchParams &p, const FastShared &sh, int t,
                                                  uint32_t k, uint32_t lane) {
  if (k < kChunkTable) return sh.chunk_list[t][k];
  const uint32_t *pre = sh.pre[t];   // rare: very dense tile, search the prefix table
  const bool hit0 = lane < p.list_len && pre[lane] <= k && k < pre[lane + 1];
  const bool hit1 = lane + 32 < p.list_len && pre[lane + 32] <= k && k < pre[lane + 33];
  const uint32_t b0 = __ballot_sync(kFull, hit0), b1 = __ballot_sync(kFull, hit1);
  return b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
}

// Lane l > 0 of chunk k holds position (slice start + 31*(k - pre[j]) + l - 1); lane 0 holds the
// position before the chunk (kNone for the first chunk of a slice), so that "first position of a
// (list, region) run" is a plain compare with the left neighbour lane.
__device__ __forceinline__ uint32_t load_chunk(const SearchParams &p, const FastShared &sh, int t,
                                               uint32_t k, uint32_t lane) {
  const uint32_t j = chunk_to_list(p, sh, t, k, lane);
  const uint32_t c = k - sh.pre[t][j];
  const uint32_t idx = sh.bnd[t][j] + kChunk * c + lane - 1;
  const bool ok = (lane > 0 || c > 0) && idx < sh.bnd[t + 1][j];
  return ok ? __ldg(p.positions + idx) : kNone;
}

// mark index (region - base) of a loaded chunk element, kNone unless it starts a run
__device__ __forceinline__ uint32_t chunk_mark(const SearchParams &p, const FastShared &sh, int t,
                                               uint32_t k, uint32_t lane, uint32_t pos, uint32_t base) {
  const uint32_t off = chunk_to_list(p, sh, t, k, lane) * p.shift;
  const uint32_t d = pos != kNone ? (pos - off) >> p.log_region : kNone;
  const uint32_t dprev = __shfl_up_sync(kFull, d, 1);
  return (lane > 0 && pos != kNone && d != dprev) ? d - base : kNone;
}

template <int T>
__device__ __forceinline__ void arrive(uint32_t *planes, uint32_t words, uint32_t x) {
  const uint32_t w = x >> 5, bit = 1u << (x & 31);
  uint32_t old = atomicOr(&planes[w], bit);
#pragma unroll
  for (int tt = 1; tt < T; ++tt)
    if (old & bit) old = atomicOr(&planes[tt * words + w], bit); else break;
}

// threshold 2 in ONE pass: cnt(d) + cnt(d+1) >= 2 with cnt(d) >= 1 holds iff two marks hit d, or
// one hits d and one hits d+1.  Whichever of the two marks arrives LATER sees the other's bit in
// the value returned by its own atomicOr (or, across a word boundary, in an ordered atomic read
// issued after it), so the decision is taken at arrival time: no second count plane, no decide
// pass.  Setting the emit bit is idempotent.  A halo mark (region base+M) only vouches for its
// left neighbour.  Returns true on a double arrival at x.
__device__ __forceinline__ bool arrive_decide(uint32_t *p0, uint32_t *emitb, uint32_t *summary,
                                              uint32_t x, bool halo) {
  const uint32_t w = x >> 5, b = x & 31, bit = 1u << b;
  const uint32_t old = atomicOr(&p0[w], bit);
  // two marks in adjacent regions that straddle a word boundary each set their own word and then
  // read the other's: without a fence between the two accesses both could miss each other
  if (b == 0 || b == 31) __threadfence_block();
  bool self = false, left = false;
  if (!halo) {
    self = (old & bit) != 0;
    self |= b < 31 ? ((old >> (b + 1)) & 1u) != 0 : (atomicOr(&p0[w + 1], 0u) & 1u) != 0;
  }
  if (b > 0) left = ((old >> (b - 1)) & 1u) != 0;
  else if (w > 0) left = (atomicOr(&p0[w - 1], 0u) >> 31) != 0;
  if (self) {
    atomicOr(&emitb[w], bit);
    atomicOr(&summary[w >> 10], 1u << ((w >> 5) & 31));
  }
  if (left) {
    const uint32_t y = x - 1, wy = y >> 5;
    atomicOr(&emitb[wy], 1u << (y & 31));
    atomicOr(&summary[wy >> 10], 1u << ((wy >> 5) & 31));
  }
  return !halo && (old & bit) != 0;
}

template <int T, bool ONEPASS>
__global__ void __launch_bounds__(kFastThreads, 2) seed_search_fast_kernel(const SearchParams p) {
  static_assert(!ONEPASS || T == 2, "the one-pass decision is the threshold-2 identity");
  constexpr int NP = ONEPASS ? 1 : T;   // count planes
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ FastShared sh;
  const uint32_t M = p.tile_regions;
  const uint32_t words = M / 32 + 1;
  uint32_t *planes = dyn;
  uint32_t *emitb = dyn + NP * words;
  uint32_t *summary = emitb + words;
  const uint32_t groups = M / 1024;
  const uint32_t n_sw = (groups + 31) / 32;       // <= 32 summary words, one per warp
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *staging = p.staging + (size_t)blockIdx.x * p.staging_cap;
  const uint32_t r = p.log_region;

  for (uint32_t i = tid; i < (NP + 1) * words + groups / 32 + 1; i += kFastThreads) dyn[i] = 0;
  if (tid == 0) sh.visited = 0;
  __syncthreads();

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    if (sh.query >= p.n_queries) break;
    const uint32_t q = p.query_list ? p.query_list[sh.query] : sh.query;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    uint32_t *out = staging;
    uint32_t out_cap = p.staging_cap;
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (tid == 0) { sh.min_d = kNone; sh.max_d = 0; sh.out_n = 0; sh.r1_double = 0; }
      __syncthreads();
      if (tid < p.list_len) {
        const uint32_t j = tid, off = j * p.shift;
        const uint32_t key = get_key(query + off, p.seed);
        uint32_t b = p.keys_count[key];
        const uint32_t e = p.keys_count[key + 1];                       // index.h:105-114
        while (b < e && p.positions[b] < off) ++b;                      // aligner.cpp:430-431
        sh.lbeg[j] = b;
        sh.lend[j] = e;
        if (b < e) {
          atomicMin(&sh.min_d, (p.positions[b] - off) >> r);
          atomicMax(&sh.max_d, (p.positions[e - 1] - off) >> r);
          if (attempt == 0) atomicAdd(&sh.visited, (unsigned long long)(e - b));
        }
      }
      __syncthreads();
      const uint32_t min_d = sh.min_d, max_d = sh.max_d;

      if (min_d != kNone) {
        const uint32_t tile_first = min_d / M, tile_last = max_d / M;
        for (uint32_t batch0 = tile_first; batch0 <= tile_last; batch0 += kTileBatch) {
          const uint32_t n_tiles = min((uint32_t)kTileBatch, tile_last - batch0 + 1);
          // ---- slice boundaries of this batch: binary search per (list, boundary)
          for (uint32_t i = tid; i < (n_tiles + 1) * p.list_len; i += kFastThreads) {
            const uint32_t t = i / p.list_len, j = i - t * p.list_len;
            const uint32_t off = j * p.shift;
            const unsigned long long target = (unsigned long long)(batch0 + t) * M;  // region
            uint32_t lo = sh.lbeg[j], hi = sh.lend[j];
            while (lo < hi) {  // first index whose region >= target
              const uint32_t mid = (lo + hi) >> 1;
              if ((unsigned long long)((__ldg(p.positions + mid) - off) >> r) < target) lo = mid + 1;
              else hi = mid;
            }
            sh.bnd[t][j] = lo;
            sh.fr[t][j] = lo < sh.lend[j] ? (__ldg(p.positions + lo) - off) >> r : kNone;
          }
          __syncthreads();
          if (tid < n_tiles) {
            uint32_t acc = 0;
            for (uint32_t j = 0; j < p.list_len; ++j) {
              sh.pre[tid][j] = acc;
              acc += (sh.bnd[tid + 1][j] - sh.bnd[tid][j] + kChunk - 1) / kChunk;
            }
            sh.pre[tid][p.list_len] = acc;
          }
          __syncthreads();
          for (uint32_t i = tid; i < n_tiles * p.list_len; i += kFastThreads) {
            const uint32_t t = i / p.list_len, j = i - t * p.list_len;
            const uint32_t k1 = min(sh.pre[t][j + 1], (uint32_t)kChunkTable);
            for (uint32_t k = sh.pre[t][j]; k < k1; ++k) sh.chunk_list[t][k] = (uint8_t)j;
          }
          __syncthreads();

          // ---- first non-empty tile of the batch: load its chunks
          uint32_t t = 0;
          while (t < n_tiles && sh.pre[t][p.list_len] == 0) ++t;
          uint32_t nxt[kSlots];
          if (t < n_tiles) {
            const uint32_t total = sh.pre[t][p.list_len];
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              const uint32_t k = warp + kFastWarps * s;
              if (k >= total) break;
              nxt[s] = load_chunk(p, sh, t, k, lane);
            }
          }
          while (t < n_tiles) {
            const uint32_t base = (batch0 + t) * M;
            const uint32_t total = sh.pre[t][p.list_len];
            // ---- marks of this tile (run starts), from the loaded chunks
            uint32_t mark[kSlots];
#pragma unroll
            for (int s = 0; s < kSlots; ++s) mark[s] = kNone;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              const uint32_t k = warp + kFastWarps * s;
              if (k >= total) break;
              mark[s] = chunk_mark(p, sh, t, k, lane, nxt[s], base);
            }
            // ---- next non-empty tile: issue its loads now (latency hides behind the passes)
            uint32_t tn = t + 1;
            while (tn < n_tiles && sh.pre[tn][p.list_len] == 0) ++tn;
            if (tn < n_tiles) {
              const uint32_t total_n = sh.pre[tn][p.list_len];
#pragma unroll
              for (int s = 0; s < kSlots; ++s) {
                const uint32_t k = warp + kFastWarps * s;
                if (k >= total_n) break;
                nxt[s] = load_chunk(p, sh, tn, k, lane);
              }
            }
            // ---- pass 1: arrive
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
              if (mark[s] != kNone) {
                if (ONEPASS) {
                  if (arrive_decide(planes, emitb, summary, mark[s], false) && base + mark[s] == 1)
                    sh.r1_double = 1;
                } else {
                  arrive<T>(planes, words, mark[s]);
                }
              }
            // chunks beyond the register slots (very dense tiles): streamed, re-read per pass
            for (uint32_t k = warp + kFastWarps * kSlots; k < total; k += kFastWarps) {
              const uint32_t x = chunk_mark(p, sh, t, k, lane, load_chunk(p, sh, t, k, lane), base);
              if (x != kNone) {
                if (ONEPASS) {
                  if (arrive_decide(planes, emitb, summary, x, false) && base + x == 1) sh.r1_double = 1;
                } else {
                  arrive<T>(planes, words, x);
                }
              }
            }
            // halo: the first position of region base+M of every list counts for region base+M-1
            if (tid < p.list_len && sh.fr[t + 1][tid] == base + M) {
              if (ONEPASS) arrive_decide(planes, emitb, summary, M, true);
              else arrive<T>(planes, words, M);
            }
            __syncthreads();  // A
            if (tid == 0 && batch0 + t == 0) {
              // aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an
              // unoccupied region 0 still emits when region 1 alone reaches the threshold.
              const bool r1 = ONEPASS ? sh.r1_double != 0 : (planes[(NP - 1) * words] & 2u) != 0;
              if (!(planes[0] & 1u) && r1) {
                atomicOr(&emitb[0], 1u);
                atomicOr(&summary[0], 1u);
              }
            }
            // ---- pass 2: decide (multi-plane variant only)
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              const uint32_t x = mark[s];
              if (!ONEPASS && x != kNone && emits<T>(planes, words, x)) {
                const uint32_t w = x >> 5;
                atomicOr(&emitb[w], 1u << (x & 31));
                atomicOr(&summary[w >> 10], 1u << ((w >> 5) & 31));
              }
            }
            for (uint32_t k = warp + kFastWarps * kSlots; !ONEPASS && k < total; k += kFastWarps) {
              const uint32_t x = chunk_mark(p, sh, t, k, lane, load_chunk(p, sh, t, k, lane), base);
              if (x != kNone && emits<T>(planes, words, x)) {
                const uint32_t w = x >> 5;
                atomicOr(&emitb[w], 1u << (x & 31));
                atomicOr(&summary[w >> 10], 1u << ((w >> 5) & 31));
              }
            }
            if (!ONEPASS) __syncthreads();  // B (the one-pass variant has nothing between A and the clear)
            // ---- pass 3: clear; compaction step 1: every warp counts the emit bits of the
            // groups of "its" summary word
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              const uint32_t x = mark[s];
              if (x != kNone) {
#pragma unroll
                for (int tt = 0; tt < NP; ++tt) planes[tt * words + (x >> 5)] = 0;
              }
            }
            for (uint32_t k = warp + kFastWarps * kSlots; k < total; k += kFastWarps) {
              const uint32_t x = chunk_mark(p, sh, t, k, lane, load_chunk(p, sh, t, k, lane), base);
              if (x != kNone) {
#pragma unroll
                for (int tt = 0; tt < NP; ++tt) planes[tt * words + (x >> 5)] = 0;
              }
            }
            if (tid < NP) planes[tid * words + (M >> 5)] = 0;   // halo word
            const uint32_t my_groups = warp < n_sw ? summary[warp] : 0u;
            {
              uint32_t cnt = 0;
              for (uint32_t gb = my_groups; gb; gb &= gb - 1)
                cnt += __popc(emitb[(warp * 32 + __ffs(gb) - 1) * 32 + lane]);
              cnt = __reduce_add_sync(kFull, cnt);
              if (lane == 0) sh.scan[warp] = cnt;
            }
            __syncthreads();  // B2
            // ---- compaction step 2: ordered write
            {
              const uint32_t mine = lane < kFastWarps ? sh.scan[lane] : 0u;
              uint32_t before = __reduce_add_sync(kFull, lane < warp ? mine : 0u) + sh.out_n;
              const uint32_t all = __reduce_add_sync(kFull, mine);
              for (uint32_t gb = my_groups; gb; gb &= gb - 1) {
                const uint32_t g = warp * 32 + __ffs(gb) - 1;
                uint32_t bits = emitb[g * 32 + lane];
                emitb[g * 32 + lane] = 0;
                const uint32_t cnt = __popc(bits);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const uint32_t v = __shfl_up_sync(kFull, incl, o);
                  if (lane >= o) incl += v;
                }
                uint32_t slot = before + incl - cnt;
                while (bits) {
                  const uint32_t b = __ffs(bits) - 1;
                  bits &= bits - 1;
                  if (slot < out_cap) out[slot] = (base + g * 1024 + lane * 32 + b) << r;
                  ++slot;
                }
                before += __shfl_sync(kFull, incl, 31);
              }
              if (warp < n_sw && lane == 0) summary[warp] = 0;
              __syncthreads();  // C
              if (tid == 0) sh.out_n += all;
            }
            t = tn;
          }
          __syncthreads();
        }
      }
      __syncthreads();
      // ---- hand the query's candidates over
      const uint32_t n = sh.out_n;
      if (attempt == 0) {
        if (tid == 0) {
          sh.base = n ? atomicAdd(p.cand_cursor, (unsigned long long)n) : 0ull;
          if (n && sh.base + n > p.cand_capacity) atomicExch(p.overflow, 1);
        }
        __syncthreads();
        const unsigned long long cbase = sh.base;
        const bool fits = cbase + n <= p.cand_capacity;
        if (tid == 0) {
          p.cand_off[q] = (uint32_t)cbase;
          p.cand_cnt[q] = fits ? n : 0u;
        }
        if (!fits || n == 0) break;
        if (n <= p.staging_cap) {
          for (uint32_t i = tid; i < n; i += kFastThreads) p.cand_start[cbase + i] = staging[i];
          break;
        }
        out = p.cand_start + cbase;   // more candidates than the staging area: redo in place
        out_cap = n;
        __syncthreads();
      }
    }
    __syncthreads();
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

// ---------------------------------------------------------------------------------------------
// Bucket path (threshold 2, list_len <= 64): no shared-memory sweep of the region space, no slice
// boundaries, no per-tile block barriers.  Per query, with the region space cut into <= 512
// tiles of 2^tile_bits regions:
//   phase A  the CTA streams ALL positions of the query's lists once, list by list, in 32-position
//            chunks dealt round robin to the warps (global chunk number mod warps, so the skewed
//            list lengths stay balanced while every per-list value is hoisted out of the chunk
//            loop); a position that starts a (list, region) run becomes a MARK and is appended as
//            a 16-bit in-tile region to its tile's bucket (shared-memory atomicAdd for the slot;
//            the buckets live in an L2-resident per-CTA scratch, 2 B per mark);
//   phase B  every WARP owns whole tiles (dynamic fetch): it reads the tile's bucket and sets the
//            marks in its PRIVATE occupancy bitmap; the first mark to arrive at a region owns it,
//            later ones (other lists) set the region's bit in a second bitmap.  An owner emits
//            its region d iff it was hit twice or region d+1 is occupied (threshold 2:
//            cnt(d) + cnt(d+1) >= 2 with cnt(d) >= 1) - exactly one emitter per region, so the few
//            emitted regions of a tile are ordered by counting ranks, no bitmap scan.  Only
//            __syncwarp inside a tile;
//   phase C  an exclusive scan over the per-tile counts orders the tiles; the candidates are
//            copied to the query's slice of the global candidate buffer (one atomicAdd).
// Region 0 of tile t+1 vouches for the last region of tile t through a per-tile halo bit.
// Queries that do not fit the fixed capacities (a bucket, a tile's emit list, the staging area -
// low-complexity queries against repetitive databases) are queued for the sweep kernel above,
// which has no such limits; the results are identical either way.
constexpr int kBkMaxTiles = 1024;
constexpr int kBkLists = 64;
constexpr int kBkUnroll = 4;            // chunks per batch in phase A (two batches in flight)
constexpr int kBkEmitCap = 126;         // emitted regions per tile (u16 list, 64 words with its counter)
constexpr uint32_t kBkEmpty = 0x7FFFFFFFu;   // empty register slot: bit 31 (owner) clear, no valid region
constexpr uint32_t kBkMinTileBits = 10;

struct BucketShared {
  uint32_t lbeg[kBkLists], lend[kBkLists];
  uint32_t pre[kBkLists + 1];             // exclusive prefix of chunks per list
  uint32_t cnt[kBkMaxTiles];              // marks per tile; phase C: output offset of the tile
  uint32_t tile_off[kBkMaxTiles];         // staging offset of the tile's candidates
  uint16_t tile_cnt[kBkMaxTiles];         // <= kBkEmitCap + 1
  uint32_t halo[kBkMaxTiles / 32 + 1];    // bit t: region 0 of tile t is occupied
  uint32_t query, stage_n, bad;
  unsigned long long base;
  unsigned long long visited;
};

// NW warps per CTA, SLOTS register-resident marks per lane per tile, MINB CTAs per SM.
template <int NW, int SLOTS, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) seed_search_bucket_kernel(const SearchParams p) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ BucketShared sh;
  constexpr uint32_t kThreads = NW * 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t r = p.log_region, TB = p.tile_bits, xmask = (1u << TB) - 1u;
  const uint32_t W = 1u << (TB - 5);                     // words per private bitmap
  const uint32_t n_tiles = (p.n_regions + xmask) >> TB;
  const uint32_t S = p.bucket_cap;
  uint32_t *occ = dyn + (size_t)warp * (2 * W + 1 + 64);  // [W + 1]: +1 halo word
  uint32_t *multi = occ + W + 1;                          // [W] regions hit by >= 2 lists
  uint32_t *elist_n = multi + W;                          // emit list: counter word + u16 regions
  uint16_t *elist = reinterpret_cast<uint16_t *>(elist_n + 1);
  uint16_t *bucket = p.buckets + (size_t)blockIdx.x * n_tiles * S;
  uint32_t *stage = p.staging + (size_t)blockIdx.x * p.staging_cap;

  for (uint32_t i = tid; i < NW * (2 * W + 1 + 64); i += kThreads) dyn[i] = 0;
  if (tid == 0) sh.visited = 0;
  __syncthreads();

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    const uint32_t q = sh.query;
    if (q >= p.n_queries) break;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    // ---- phase 0: intervals (index.h:105-114), chunk prefix, counters
    if (tid < p.list_len) {
      const uint32_t j = tid, off = j * p.shift;
      const uint32_t key = get_key(query + off, p.seed);
      uint32_t b = p.keys_count[key];
      const uint32_t e = p.keys_count[key + 1];
      while (b < e && p.positions[b] < off) ++b;                      // aligner.cpp:430-431
      sh.lbeg[j] = b;
      sh.lend[j] = e;
      if (e > b) atomicAdd(&sh.visited, (unsigned long long)(e - b));
    }
    for (uint32_t t = tid; t < n_tiles; t += kThreads) { sh.cnt[t] = 0; sh.tile_cnt[t] = 0; }
    if (tid < kBkMaxTiles / 32 + 1) sh.halo[tid] = 0;
    if (tid == 0) { sh.stage_n = 0; sh.bad = 0; }
    __syncthreads();
    if (warp == 0) {
      uint32_t carry = 0;
      for (uint32_t j0 = 0; j0 < p.list_len; j0 += 32) {
        const uint32_t j = j0 + lane;
        const uint32_t c = j < p.list_len ? (sh.lend[j] - sh.lbeg[j] + 31) / 32 : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
        if (j < p.list_len) sh.pre[j] = carry + incl - c;
        carry += __shfl_sync(kFull, incl, 31);
      }
      if (lane == 0) sh.pre[p.list_len] = carry;
    }
    __syncthreads();

    // ---- phase A: marks -> tile buckets.  The warp walks its batches (up to kBkUnroll chunks of
    // one list) with the loads of the next batch in flight while it scatters the current one.
    {
      uint32_t j = 0, c = 0, lb = 0, le = 0, nj = 0;
      auto seek = [&]() {   // first list >= j holding a chunk for this warp
        while (j < p.list_len) {
          lb = sh.lbeg[j];
          le = sh.lend[j];
          nj = (le - lb + 31) / 32;
          c = (warp + NW - sh.pre[j] % NW) % NW;   // global chunk number (pre[j] + c) % NW == warp
          if (c < nj) return;
          ++j;
        }
      };
      auto load = [&](uint32_t (&pos)[kBkUnroll], uint32_t (&prev)[kBkUnroll]) {
#pragma unroll
        for (int u = 0; u < kBkUnroll; ++u) {
          const uint32_t idx = lb + 32u * (c + u * NW) + lane;
          pos[u] = kNone;
          prev[u] = kNone;
          if (j < p.list_len && idx < le) {
            pos[u] = __ldg(p.positions + idx);
            if (idx > lb) prev[u] = __ldg(p.positions + idx - 1);
          }
        }
      };
      seek();
      uint32_t pos[kBkUnroll], prev[kBkUnroll];
      load(pos, prev);
      while (j < p.list_len) {
        const uint32_t off = j * p.shift;
        c += NW * kBkUnroll;
        if (c >= nj) { ++j; seek(); }
        uint32_t npos[kBkUnroll], nprev[kBkUnroll];
        load(npos, nprev);
#pragma unroll
        for (int u = 0; u < kBkUnroll; ++u) {
          if (pos[u] != kNone) {
            const uint32_t d = (pos[u] - off) >> r;
            if (prev[u] == kNone || ((prev[u] - off) >> r) != d) {   // first of its list in region d
              const uint32_t t = d >> TB, x = d & xmask;
              const uint32_t slot = atomicAdd(&sh.cnt[t], 1u);
              if (slot < S) bucket[t * S + slot] = (uint16_t)x;
              else sh.bad = 1;
              if (x == 0) atomicOr(&sh.halo[t >> 5], 1u << (t & 31));
            }
          }
          pos[u] = npos[u];
          prev[u] = nprev[u];
        }
      }
    }
    __syncthreads();
    bool bad = sh.bad != 0;

    if (!bad) {
      // ---- phase B: one warp per tile (t = warp, warp + NW, ...), next tile's bucket prefetched
      uint32_t t = warp;
      uint32_t n = t < n_tiles ? sh.cnt[t] : 0u;          // n == 0 also guards the address below
      uint32_t xs[SLOTS];
      {
        const uint16_t *b0 = bucket + (t < n_tiles ? t : 0u) * S + lane;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) xs[s] = (s * 32 + lane < n) ? b0[s * 32] : 0xFFFFu;
      }
      while (t < n_tiles) {
        const uint32_t tn = t + NW;
        const uint32_t nn = tn < n_tiles ? sh.cnt[tn] : 0u;
        uint32_t nxs[SLOTS];
        {
          const uint16_t *bn = bucket + (tn < n_tiles ? tn : 0u) * S + lane;
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) nxs[s] = (s * 32 + lane < nn) ? bn[s * 32] : 0xFFFFu;
        }
        if (n != 0) {
          uint16_t *bk = bucket + t * S;
          if (lane == 0)
            occ[W] = (t + 1 < n_tiles) ? (sh.halo[(t + 1) >> 5] >> ((t + 1) & 31)) & 1u : 0u;
          // arrive: the first mark at a region owns it (bit 31 of m), later ones flag it in `multi`
          uint32_t m[SLOTS];
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            m[s] = kBkEmpty;
            if (xs[s] != 0xFFFFu) {
              const uint32_t x = xs[s], bit = 1u << (x & 31);
              const uint32_t old = atomicOr(&occ[x >> 5], bit);
              m[s] = x;
              if (old & bit) atomicOr(&multi[x >> 5], bit);
              else m[s] = x | 0x80000000u;
            }
          }
          for (uint32_t i = SLOTS * 32 + lane; i < n; i += 32) {   // beyond the register slots
            const uint32_t x = bk[i], bit = 1u << (x & 31);
            const uint32_t old = atomicOr(&occ[x >> 5], bit);
            if (old & bit) atomicOr(&multi[x >> 5], bit);
            else bk[i] = (uint16_t)(x | 0x8000u);                    // x < 2^15: bit 15 = owner
          }
          __syncwarp();
          // decide: owners only; emitted regions go to the warp's emit list (unordered)
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            if ((int)m[s] >= 0) continue;                            // not an owner / empty slot
            const uint32_t x = m[s] & 0x7FFFFFFFu, y = x + 1;
            if (((multi[x >> 5] >> (x & 31)) | (occ[y >> 5] >> (y & 31))) & 1u) {
              const uint32_t i = atomicAdd(elist_n, 1u);
              if (i < kBkEmitCap) elist[i] = (uint16_t)x;
            }
          }
          for (uint32_t i = SLOTS * 32 + lane; i < n; i += 32) {
            const uint32_t v = bk[i];
            if (!(v & 0x8000u)) continue;
            const uint32_t x = v & 0x7FFFu, y = x + 1;
            if (((multi[x >> 5] >> (x & 31)) | (occ[y >> 5] >> (y & 31))) & 1u) {
              const uint32_t k = atomicAdd(elist_n, 1u);
              if (k < kBkEmitCap) elist[k] = (uint16_t)x;
            }
          }
          // aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an
          // unoccupied region 0 still emits when region 1 alone reaches the threshold.
          const uint32_t virt = (t == 0 && !(occ[0] & 1u) && (multi[0] & 2u)) ? 1u : 0u;
          __syncwarp();
          // clear what this tile set
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            if (m[s] == kBkEmpty) continue;
            const uint32_t w = (m[s] & 0x7FFFFFFFu) >> 5;
            if ((int)m[s] < 0) occ[w] = 0; else multi[w] = 0;
          }
          for (uint32_t i = SLOTS * 32 + lane; i < n; i += 32) {
            const uint32_t v = bk[i], w = (v & 0x7FFFu) >> 5;
            if (v & 0x8000u) occ[w] = 0; else multi[w] = 0;
          }
          const uint32_t k = *elist_n;
          __syncwarp();
          if (lane == 0) { occ[W] = 0; *elist_n = 0; }
          if (k > kBkEmitCap) {
            if (lane == 0) sh.bad = 1;
          } else if (k + virt) {
            // exactly one entry per emitted region: its rank is the number of smaller entries
            uint32_t off = 0;
            if (lane == 0) {
              off = atomicAdd(&sh.stage_n, k + virt);
              sh.tile_off[t] = off;
              sh.tile_cnt[t] = (uint16_t)(k + virt);
            }
            off = __shfl_sync(kFull, off, 0);
            if (virt && lane == 0 && off < p.staging_cap) stage[off] = 0;   // region 0 of tile 0
            for (uint32_t i = lane; i < k; i += 32) {
              const uint32_t x = elist[i];
              uint32_t rank = virt;
              for (uint32_t i2 = 0; i2 < k; ++i2) rank += elist[i2] < x;
              if (off + rank < p.staging_cap) stage[off + rank] = ((t << TB) + x) << r;
            }
          }
          __syncwarp();
        }
        t = tn;
        n = nn;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) xs[s] = nxs[s];
      }
      __syncthreads();
      bad = sh.bad != 0 || sh.stage_n > p.staging_cap;
    }

    if (bad) {   // exceeds a fixed capacity: the sweep kernel redoes this query (planes are clean)
      if (tid == 0) {
        p.fallback_list[atomicAdd(p.fallback_n, 1u)] = q;
        p.cand_off[q] = 0;
        p.cand_cnt[q] = 0;
      }
      continue;
    }

    // ---- phase C: order the tiles, hand the candidates over
    const uint32_t n = sh.stage_n;
    if (warp == 0) {
      uint32_t carry = 0;
      for (uint32_t t0 = 0; t0 < n_tiles; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint32_t c = t < n_tiles ? sh.tile_cnt[t] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
        if (t < n_tiles) sh.cnt[t] = carry + incl - c;
        carry += __shfl_sync(kFull, incl, 31);
      }
      if (lane == 0) {
        sh.base = n ? atomicAdd(p.cand_cursor, (unsigned long long)n) : 0ull;
        const bool fits = sh.base + n <= p.cand_capacity;
        if (n && !fits) atomicExch(p.overflow, 1);
        p.cand_off[q] = (uint32_t)sh.base;
        p.cand_cnt[q] = fits ? n : 0u;
      }
    }
    __syncthreads();
    const unsigned long long cbase = sh.base;
    if (n && cbase + n <= p.cand_capacity) {
      for (uint32_t t = warp; t < n_tiles; t += NW) {
        const uint32_t c = sh.tile_cnt[t], src = sh.tile_off[t], dst = sh.cnt[t];
        for (uint32_t i = lane; i < c; i += 32) p.cand_start[cbase + dst + i] = stage[src + i];
      }
    }
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

// ---------------------------------------------------------------------------------------------
// Hash path (threshold 2, list_len <= 64): no tiles at all.  Per query, one CTA of 32 warps:
//   pass 1  stream all positions (32-position chunks dealt to the warps through a chunk->list
//           table); every MARK (first position of its list in a region d) sets bit d mod 2^20 of a
//           FOLDED occupancy bitmap (128 KB); a mark that finds its bit set also sets the coarse
//           collision bitmap (one bit per 8 folded bits, 16 KB);
//   pass 2  stream the same positions again (L2 hits): a mark is a SUSPECT iff its collision bit
//           or one of its two neighbour bits in the folded bitmap is set.  Every mark that takes
//           part in an emission is a suspect: two marks of one region collide, a mark of region d
//           and one of d+1 are each other's neighbours; aliases of the fold only add false
//           suspects (~14 % of the marks).  Suspects are appended with their exact region;
//   pass 3  the bitmaps are cleared and their space becomes an open-addressing hash set of the
//           suspects' exact regions: the first inserter of a region owns it, later ones flag it;
//   pass 4  every owner emits its region d iff it is flagged (two lists) or d+1 is in the set
//           (threshold 2: cnt(d) + cnt(d+1) >= 2, cnt(d) >= 1), plus the reference's virtual
//           region 0 (aligner.cpp:451,483-494).  Emitted regions go to 32 range buckets
//           (monotone in d), one warp ranks each bucket by counting: ascending output.
// Everything is exact; the fold only decides how many suspects there are.  Queries that exceed
// a capacity (chunk table, suspect list, a range bucket) are queued for the sweep kernel.
constexpr int kHsThreads = 1024;
constexpr int kHsWarps = kHsThreads / 32;
constexpr uint32_t kHsOccBits = 1u << 20;
constexpr uint32_t kHsOccWords = kHsOccBits / 32;       // 32768 words = 128 KB; later the hash set
constexpr uint32_t kHsCollWords = kHsOccWords / 8;      // 4096 words = 16 KB; later the range buckets
constexpr uint32_t kHsSuspCap = 16384;                  // 64 KB
constexpr uint32_t kHsBuckets = 32;
constexpr uint32_t kHsBucketCap = kHsCollWords / kHsBuckets;   // 128 emitted regions per bucket
constexpr int kHsMaxChunks = 4096;
constexpr int kHsUnroll = 4;
constexpr size_t kHsSmemBytes = (size_t)(kHsOccWords + kHsCollWords + kHsSuspCap) * 4;

struct HashShared {
  uint32_t lbeg[kBkLists], lend[kBkLists];
  uint32_t pre[kBkLists + 1];
  uint8_t chunk_list[kHsMaxChunks];
  uint32_t bcnt[kHsBuckets], boff[kHsBuckets + 1];
  uint32_t query, n_susp, bad;
  unsigned long long base;
  unsigned long long visited;
};

__device__ __forceinline__ uint32_t hs_slot(uint32_t d) { return (d * 2654435761u) >> 17; }   // 15 bits

// region d present in the hash set?  (*flagged: inserted more than once)
__device__ __forceinline__ bool hs_find(const uint32_t *set, uint32_t d, bool *flagged) {
  uint32_t s = hs_slot(d);
  while (true) {
    const uint32_t v = set[s];
    if (v == 0) return false;
    if ((v & 0x7FFFFFFFu) == d + 1) { *flagged = (v >> 31) != 0; return true; }
    s = (s + 1) & (kHsOccWords - 1);
  }
}

// One streaming pass over the query's positions.  Branch-free chunk addressing: chunk k of the
// warp's round (clamped to the last chunk when the round runs past the end) -> list j, first index;
// lane 0 also fetches the predecessor of the chunk's first position, the other lanes get theirs
// from the left neighbour, so "first position of my list in region d" is one compare.
template <int PASS>
__device__ __forceinline__ void hs_stream(const SearchParams &p, HashShared &sh, uint32_t *occ,
                                          uint32_t *coll, uint32_t *susp, uint32_t C, uint32_t warp,
                                          uint32_t lane) {
  const uint32_t r = p.log_region, hmask = kHsOccBits - 1, lt = (1u << lane) - 1u;
  for (uint32_t k0 = warp; k0 < C; k0 += kHsWarps * kHsUnroll) {
    uint32_t pos[kHsUnroll], prev0[kHsUnroll], offs[kHsUnroll];
#pragma unroll
    for (int u = 0; u < kHsUnroll; ++u) {
      const uint32_t k = k0 + u * kHsWarps;
      const uint32_t kc = k < C ? k : C - 1;
      const uint32_t j = sh.chunk_list[kc];
      const uint32_t b = sh.lbeg[j];
      const uint32_t idx = b + 32u * (kc - sh.pre[j]) + lane;
      const bool ok = k < C && idx < sh.lend[j];
      offs[u] = j * p.shift;
      pos[u] = ok ? __ldg(p.positions + idx) : kNone;
      prev0[u] = (ok && lane == 0 && idx > b) ? __ldg(p.positions + idx - 1) : kNone;
    }
#pragma unroll
    for (int u = 0; u < kHsUnroll; ++u) {
      const uint32_t d = (pos[u] - offs[u]) >> r;
      uint32_t pv = __shfl_up_sync(kFull, pos[u], 1);
      if (lane == 0) pv = prev0[u];
      // a valid lane > 0 always has a valid left neighbour of the same list; pv == kNone only for
      // the very first position of the list
      const bool mark = pos[u] != kNone && (pv == kNone || ((pv - offs[u]) >> r) != d);
      const uint32_t h = d & hmask;
      if (PASS == 1) {
        if (mark) {
          const uint32_t bit = 1u << (h & 31);
          const uint32_t old = atomicOr(&occ[h >> 5], bit);
          if (old & bit) atomicOr(&coll[h >> 8], 1u << ((h >> 3) & 31));
        }
      } else {
        const uint32_t hl = (h - 1) & hmask, hr = (h + 1) & hmask;
        const uint32_t t = ((occ[hl >> 5] >> (hl & 31)) | (occ[hr >> 5] >> (hr & 31)) |
                            (coll[h >> 8] >> ((h >> 3) & 31))) & 1u;
        const bool s = mark && t;
        const uint32_t bal = __ballot_sync(kFull, s);
        if (bal) {
          uint32_t at = 0;
          if (lane == 0) at = atomicAdd(&sh.n_susp, (uint32_t)__popc(bal));
          at = __shfl_sync(kFull, at, 0);
          const uint32_t i = at + __popc(bal & lt);
          if (s && i < kHsSuspCap) susp[i] = d;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kHsThreads, 1) seed_search_hash_kernel(const SearchParams p) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ HashShared sh;
  uint32_t *occ = dyn;                       // [kHsOccWords]   passes 1-2: folded bitmap; 3-4: hash set
  uint32_t *coll = dyn + kHsOccWords;        // [kHsCollWords]  passes 1-2: collisions; 4: range buckets
  uint32_t *susp = coll + kHsCollWords;      // [kHsSuspCap]
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t r = p.log_region;
  // range bucket of a region: monotone in d, 32 buckets over [0, n_regions)
  uint32_t bshift = 0;
  while ((p.n_regions >> bshift) > kHsBuckets) ++bshift;
  if ((p.n_regions >> bshift) == kHsBuckets) ++bshift;    // d < n_regions  =>  d >> bshift < 32

  for (uint32_t i = tid; i < kHsOccWords + kHsCollWords; i += kHsThreads) dyn[i] = 0;
  if (tid == 0) sh.visited = 0;
  __syncthreads();

  while (true) {
    if (tid == 0) sh.query = atomicAdd(p.query_counter, 1u);
    __syncthreads();
    const uint32_t q = sh.query;
    if (q >= p.n_queries) break;
    const uint8_t *query = p.queries + (size_t)q * p.query_len;

    // ---- phase 0: intervals (index.h:105-114), chunk prefix and chunk -> list table
    if (tid < p.list_len) {
      const uint32_t j = tid, off = j * p.shift;
      const uint32_t key = get_key(query + off, p.seed);
      uint32_t b = p.keys_count[key];
      const uint32_t e = p.keys_count[key + 1];
      while (b < e && p.positions[b] < off) ++b;                      // aligner.cpp:430-431
      sh.lbeg[j] = b;
      sh.lend[j] = e;
      if (e > b) atomicAdd(&sh.visited, (unsigned long long)(e - b));
    }
    if (tid < kHsBuckets) sh.bcnt[tid] = 0;
    if (tid == 0) { sh.n_susp = 0; sh.bad = 0; }
    __syncthreads();
    if (warp == 0) {
      uint32_t carry = 0;
      for (uint32_t j0 = 0; j0 < p.list_len; j0 += 32) {
        const uint32_t j = j0 + lane;
        const uint32_t c = j < p.list_len ? (sh.lend[j] - sh.lbeg[j] + 31) / 32 : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
        if (j < p.list_len) sh.pre[j] = carry + incl - c;
        carry += __shfl_sync(kFull, incl, 31);
      }
      if (lane == 0) sh.pre[p.list_len] = carry;
    }
    __syncthreads();
    const uint32_t C = sh.pre[p.list_len];
    bool bad = C > kHsMaxChunks;
    if (!bad) {
      for (uint32_t k = tid; k < C; k += kHsThreads) {
        uint32_t lo = 0, hi = p.list_len;      // pre[lo] <= k < pre[hi]
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (sh.pre[mid] <= k) lo = mid; else hi = mid;
        }
        sh.chunk_list[k] = (uint8_t)lo;
      }
      __syncthreads();

      // ---- passes 1 and 2 over the same chunks
      hs_stream<1>(p, sh, occ, coll, susp, C, warp, lane);
      __syncthreads();
      hs_stream<2>(p, sh, occ, coll, susp, C, warp, lane);
      __syncthreads();
      bad = sh.n_susp > kHsSuspCap;
    }

    // ---- the bitmaps are done: clear them (also on the bad path), their space is reused
    {
      uint4 *z = reinterpret_cast<uint4 *>(dyn);
      for (uint32_t i = tid; i < (kHsOccWords + kHsCollWords) / 4; i += kHsThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    const uint32_t n_susp = sh.n_susp;
    if (!bad) {
      // ---- pass 3: hash set of the suspects' regions; bit 31 of susp[i] = owner of its region
      for (uint32_t i = tid; i < n_susp; i += kHsThreads) {
        const uint32_t d = susp[i];
        uint32_t s = hs_slot(d);
        while (true) {
          const uint32_t cur = atomicCAS(&occ[s], 0u, d + 1);
          if (cur == 0) { susp[i] = d | 0x80000000u; break; }
          if ((cur & 0x7FFFFFFFu) == d + 1) { atomicOr(&occ[s], 0x80000000u); break; }
          s = (s + 1) & (kHsOccWords - 1);
        }
      }
      __syncthreads();
      // ---- pass 4: owners decide; emitted regions into 32 range buckets
      for (uint32_t i = tid; i < n_susp + 1; i += kHsThreads) {
        uint32_t d;
        bool emit;
        if (i < n_susp) {
          const uint32_t v = susp[i];
          if (!(v >> 31)) continue;
          d = v & 0x7FFFFFFFu;
          bool flagged = false, f2 = false;
          hs_find(occ, d, &flagged);
          emit = flagged || hs_find(occ, d + 1, &f2);
        } else {
          // aligner.cpp:451,483-494: `distance` starts at region 0 with count 0, so an unoccupied
          // region 0 still emits when region 1 alone reaches the threshold.
          bool f0 = false, f1 = false;
          d = 0;
          emit = !hs_find(occ, 0, &f0) && hs_find(occ, 1, &f1) && f1;
        }
        if (emit) {
          const uint32_t bk = d >> bshift;
          const uint32_t at = atomicAdd(&sh.bcnt[bk], 1u);
          if (at < kHsBucketCap) coll[bk * kHsBucketCap + at] = d;
          else sh.bad = 1;
        }
      }
      __syncthreads();
      bad = sh.bad != 0;
    }
    // ---- clear the hash set for the next query
    {
      uint4 *z = reinterpret_cast<uint4 *>(occ);
      for (uint32_t i = tid; i < kHsOccWords / 4; i += kHsThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (bad) {
      if (tid < kHsBuckets) {   // bucket storage back to zero
        const uint32_t c = min(sh.bcnt[tid], kHsBucketCap);
        for (uint32_t i = 0; i < c; ++i) coll[tid * kHsBucketCap + i] = 0;
      }
      if (tid == 0) {
        p.fallback_list[atomicAdd(p.fallback_n, 1u)] = q;
        p.cand_off[q] = 0;
        p.cand_cnt[q] = 0;
      }
      __syncthreads();
      continue;
    }
    // ---- output: bucket offsets, global slice, ranks inside each bucket (one warp per bucket)
    if (warp == 0) {
      const uint32_t c = sh.bcnt[lane];
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      sh.boff[lane] = incl - c;
      if (lane == 31) {
        const uint32_t n = incl;
        sh.boff[32] = n;
        sh.base = n ? atomicAdd(p.cand_cursor, (unsigned long long)n) : 0ull;
        const bool fits = sh.base + n <= p.cand_capacity;
        if (n && !fits) atomicExch(p.overflow, 1);
        p.cand_off[q] = (uint32_t)sh.base;
        p.cand_cnt[q] = fits ? n : 0u;
      }
    }
    __syncthreads();
    {
      const uint32_t n = sh.boff[32];
      const unsigned long long cbase = sh.base;
      const bool fits = cbase + n <= p.cand_capacity;
      const uint32_t c = sh.bcnt[warp];
      uint32_t *bk = coll + warp * kHsBucketCap;
      for (uint32_t i = lane; i < c; i += 32) {
        const uint32_t d = bk[i];
        uint32_t rank = 0;
        for (uint32_t i2 = 0; i2 < c; ++i2) rank += bk[i2] < d;      // regions are distinct
        if (fits) p.cand_start[cbase + sh.boff[warp] + rank] = d << r;
      }
      __syncwarp();
      for (uint32_t i = lane; i < c; i += 32) bk[i] = 0;
    }
    __syncthreads();
  }
  if (tid == 0 && sh.visited) atomicAdd(p.positions_visited, sh.visited);
}

// Generic dynamic shared memory size for a tile of M regions.
size_t search_smem_bytes(int planes, uint32_t M) {   // count planes + the emit bitmap + summary
  const uint32_t words = M / 32 + 1;
  return ((size_t)(planes + 1) * words + M / 1024 / 32 + 1) * sizeof(uint32_t);
}

int search_planes(int T, bool fast) { return (fast && T == 2) ? 1 : T; }

}  // namespace

bool search_uses_fast(uint32_t list_len, uint32_t threshold, bool allow_fast) {
  return allow_fast && list_len <= kFastLists && threshold <= 4;
}

// Largest tile (multiple of 1024 regions, at most 1024 groups) that fits `smem_limit`.
uint32_t search_tile_regions(int T_in, size_t smem_limit, uint32_t n_regions, bool fast) {
  const int T = search_planes(T_in, fast);
  uint32_t m = 1024u * 1024u;
  if (fast) {  // two CTAs per SM, at most kFastWarps summary words (32 groups of 1024 regions each)
    smem_limit = smem_limit / 2 - 2048;
    m = kFastWarps * 32u * 1024u;
  }
  while (m > 1024 && search_smem_bytes(T, m) + sizeof(SearchShared) + 256 > smem_limit) m -= 1024;
  while (m > 1024 && search_smem_bytes(T, m) + sizeof(FastShared) + 256 > smem_limit) m -= 1024;
  const uint32_t need = ((n_regions + 1023) / 1024) * 1024;
  return m < need ? m : (need ? need : 1024);
}

cudaError_t seed_search_launch(const SearchParams &p, int grid, cudaStream_t stream, bool allow_fast) {
  const int T = (int)p.threshold;
  const bool fast = search_uses_fast(p.list_len, p.threshold, allow_fast);
  const size_t smem = search_smem_bytes(search_planes(T, fast), p.tile_regions);
  cudaError_t err = cudaSuccess;
#define GM_LAUNCH_SEARCH(TT)                                                                   \
  case TT:                                                                                     \
    if (fast) {                                                                                \
      err = allow_max_dynamic_smem(seed_search_fast_kernel<TT, TT == 2>);                      \
      if (err != cudaSuccess) return err;                                                      \
      seed_search_fast_kernel<TT, TT == 2><<<grid * 2, kFastThreads, smem, stream>>>(p);       \
    } else {                                                                                   \
      err = allow_max_dynamic_smem(seed_search_kernel<TT>);                                    \
      if (err != cudaSuccess) return err;                                                      \
      seed_search_kernel<TT><<<grid, kSearchThreads, smem, stream>>>(p);                       \
    }                                                                                          \
    break;
  switch (T) {
    GM_LAUNCH_SEARCH(1)
    GM_LAUNCH_SEARCH(2)
    GM_LAUNCH_SEARCH(3)
    GM_LAUNCH_SEARCH(4)
    default:   // thresholds above 4: the generic kernel with run-time plane count
      if (fast) return cudaErrorInvalidValue;
      err = allow_max_dynamic_smem(seed_search_kernel<0>);
      if (err != cudaSuccess) return err;
      seed_search_kernel<0><<<grid, kSearchThreads, smem, stream>>>(p);
      break;
  }
#undef GM_LAUNCH_SEARCH
  return cudaGetLastError();
}

// ---- bucket path configuration
struct BucketConfig { int nw, slots, minb; uint32_t tile_bits, bucket_cap; };

// Tiles of about 2 * 32 * slots marks for the expected mark density (~ list_len * interval length
// per query); GM_BUCKET_CFG="nw,slots,max_tile_bits" overrides the default for tuning runs.
static BucketConfig bucket_config(uint32_t n_regions) {
  BucketConfig c = {12, 4, 3, 14, 384};
  if (const char *env = getenv("GM_BUCKET_CFG")) {
    int nw = 0, slots = 0, tb = 0;
    if (sscanf(env, "%d,%d,%d", &nw, &slots, &tb) == 3) {
      c.nw = nw; c.slots = slots; c.tile_bits = (uint32_t)tb;
      c.bucket_cap = 384;
    }
  }
  uint32_t tb = kBkMinTileBits;
  const uint32_t want_tiles = c.tile_bits >= 15 ? 256 : 512;
  while (tb < c.tile_bits && ((n_regions + (1u << tb) - 1) >> tb) > want_tiles) ++tb;
  c.tile_bits = tb;
  return c;
}

bool search_bucket_ok(uint32_t threshold, uint32_t list_len, uint32_t n_regions, uint32_t *tile_bits,
                      uint32_t *bucket_cap) {
  if (threshold != 2 || list_len > kBkLists) return false;
  const BucketConfig c = bucket_config(n_regions);
  if (((n_regions + (1u << c.tile_bits) - 1) >> c.tile_bits) > kBkMaxTiles) return false;
  *tile_bits = c.tile_bits;
  *bucket_cap = c.bucket_cap;
  return true;
}

int search_bucket_grid(int sm_count) { return sm_count * 6; }   // upper bound for the scratch size

template <int NW, int SLOTS, int MINB>
static cudaError_t bucket_launch(const SearchParams &p, int sm_count, cudaStream_t stream) {
  const size_t smem = (size_t)NW * (2 * (1u << (p.tile_bits - 5)) + 1 + 64) * sizeof(uint32_t);
  auto kern = seed_search_bucket_kernel<NW, SLOTS, MINB>;
  cudaError_t err = allow_max_dynamic_smem(kern);
  if (err != cudaSuccess) return err;
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  if (per_sm > 6) per_sm = 6;
  kern<<<sm_count * per_sm, NW * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t seed_search_bucket_launch(const SearchParams &p, int sm_count, cudaStream_t stream) {
  const BucketConfig c = bucket_config(p.n_regions);
  if (c.nw == 12 && c.slots == 8) return bucket_launch<12, 8, 2>(p, sm_count, stream);
  if (c.nw == 12 && c.slots == 4) return bucket_launch<12, 4, 3>(p, sm_count, stream);
  if (c.nw == 8 && c.slots == 8) return bucket_launch<8, 8, 3>(p, sm_count, stream);
  if (c.nw == 8 && c.slots == 4) return bucket_launch<8, 4, 4>(p, sm_count, stream);
  if (c.nw == 4 && c.slots == 4) return bucket_launch<4, 4, 8>(p, sm_count, stream);
  if (c.nw == 16 && c.slots == 4) return bucket_launch<16, 4, 2>(p, sm_count, stream);
  return cudaErrorInvalidValue;
}

bool search_hash_ok(uint32_t threshold, uint32_t list_len, uint32_t n_regions) {
  return threshold == 2 && list_len <= kBkLists && n_regions <= (1u << 27);
}

cudaError_t seed_search_hash_launch(const SearchParams &p, int sm_count, cudaStream_t stream) {
  cudaError_t err = allow_max_dynamic_smem(seed_search_hash_kernel);
  if (err != cudaSuccess) return err;
  seed_search_hash_kernel<<<sm_count, kHsThreads, kHsSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

int search_max_list_len() { return kMaxListLen; }
int search_max_threshold() { return 1024; }   // generic kernel: one bit plane per count

}  // namespace gm
