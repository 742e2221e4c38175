// Host side of the C ABI (include/ghostm_b200.h): device buffers, stage orchestration, the
// extended gm_* entry points and the ten legacy symbols of the reference's aligner_gpu.h.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>


#include "gm_common.cuh"

namespace gm {

// ---- kernels defined in the other translation units -----------------------------------
int sw_rows_per_strip(uint32_t query_len, uint32_t *n_strips);
int sw_pair_rows(uint32_t query_len, uint32_t *n_strips);
cudaError_t sw_extend_pair_launch(const SwParams &p, int rows, int sm_count, cudaStream_t stream);
cudaError_t sw_extend_launch(const SwParams &p, int rows, int sm_count, cudaStream_t stream);
cudaError_t sw_extend_s32_launch(const SwParams &p, int sm_count, cudaStream_t stream);
uint32_t search_tile_regions(int T, size_t smem_limit, uint32_t n_regions, bool fast);
bool search_uses_fast(uint32_t list_len, uint32_t threshold, bool allow_fast);
cudaError_t seed_search_launch(const SearchParams &p, int grid, cudaStream_t stream, bool allow_fast);
bool search_bucket_ok(uint32_t threshold, uint32_t list_len, uint32_t n_regions, uint32_t *tile_bits,
                      uint32_t *bucket_cap);
int search_bucket_grid(int sm_count);
bool search_hash_ok(uint32_t threshold, uint32_t list_len, uint32_t n_regions);
cudaError_t seed_search_hash_launch(const SearchParams &p, int sm_count, cudaStream_t stream);
cudaError_t seed_search_bucket_launch(const SearchParams &p, int sm_count, cudaStream_t stream);
bool search_tile_geometry(uint32_t threshold, uint32_t list_len, uint32_t shift, uint32_t log_region,
                          uint32_t seq_len, uint32_t n_keys, size_t smem_per_sm, TileGeometry *g);
int search_tile_grid(int sm_count);
size_t search_tile_emap_words(const TileGeometry &g);
size_t index_sort_scratch_words(uint32_t n);
cudaError_t index_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp,
                             uint32_t *vals_final, uint32_t n, uint32_t bits, uint32_t *scratch,
                             uint32_t **sorted_keys, uint32_t *launches, cudaStream_t stream);
cudaError_t search_split_build(const uint32_t *keys_count, uint32_t n_keys, const uint32_t *positions,
                               const TileGeometry &g, uint32_t *split, int sm_count, cudaStream_t stream);
cudaError_t seed_search_tile_launch(SearchParams p, const TileGeometry &g, int sm_count,
                                    cudaStream_t stream);
int search_max_list_len();
int search_max_threshold();
size_t merge_smem_bytes(uint32_t elems_per_warp);
cudaError_t merge_launch(const MergeParams &p, int sm_count, cudaStream_t stream);
int traceback_grid(int sm_count);
int traceback_threads();
cudaError_t traceback_launch(const TracebackParams &p, int sm_count, cudaStream_t stream, bool fast);
bool traceback_fast_ok(uint32_t query_len, int open_gap, int extend_gap);
bool traceback_warp_ok(uint32_t query_len, int open_gap, int extend_gap);
cudaError_t collect_pending_launch(const gm_hit *hits, const uint32_t *counts, uint32_t n_queries,
                                   uint32_t cap, const ChunkRef *chunks, uint32_t *jobs,
                                   uint32_t *n_jobs, uint32_t *n_left, int sm_count,
                                   cudaStream_t stream);

namespace {

thread_local std::string g_error;

int fail(gm_status st, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return (int)st;
}

#define GM_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(GM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

// ---- small utility kernels --------------------------------------------------------------

// out[i] = exclusive prefix over i of f(in[first + i]); out[n] = total.  One CTA.
// mode 0: f(x) = x      mode 1: f(x) = ceil(x / 64)   mode 2: f(x) = ceil(x / 32)   (SW tasks per query)
__global__ void __launch_bounds__(1024) scan_kernel(const uint32_t *in, uint32_t first, uint32_t n,
                                                    uint32_t *out, int mode) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + tid;
    uint32_t v = 0;
    if (i < n) {
      v = in[first + i];
      if (mode == 1) v = (v + kSwCandPerTask - 1) / kSwCandPerTask;
      if (mode == 2) v = (v + kSwCandPerTask / 2 - 1) / (kSwCandPerTask / 2);
    }
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = carry;
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    if (i < n) out[i] = before + incl - v;
    __syncthreads();
    if (tid == 1023) carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry;
}

// Copy per-query candidate slices into reference order (queries ascending).
__global__ void gather_kernel(const uint32_t *cand_off, const uint32_t *cand_cnt,
                              const uint32_t *prefix, uint32_t first, uint32_t n_q,
                              const uint32_t *src0, const uint32_t *src1, uint32_t *dst0,
                              uint32_t *dst1, uint32_t *dst_query) {
  for (uint32_t qi = blockIdx.x; qi < n_q; qi += gridDim.x) {
    const uint32_t q = first + qi, off = cand_off[q], cnt = cand_cnt[q], o = prefix[qi];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      if (dst0) dst0[o + i] = src0[off + i];
      if (dst1) dst1[o + i] = src1[off + i];
      if (dst_query) dst_query[o + i] = q;
    }
  }
}

// gm_candidates_pack: per-query candidate slices -> per-part [score | end] blocks in
// reference order.  prefix[] is the exclusive scan of cand_cnt over all queries.
__global__ void pack_kernel(const uint32_t *cand_off, const uint32_t *cand_cnt,
                            const uint32_t *prefix, uint32_t n_q, const uint32_t *bounds,
                            uint32_t n_parts, const uint32_t *score, const uint32_t *end,
                            uint32_t *out) {
  for (uint32_t q = blockIdx.x; q < n_q; q += gridDim.x) {
    uint32_t lo = 0, hi = n_parts;  // bounds[lo] <= q < bounds[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (bounds[mid] <= q) lo = mid; else hi = mid;
    }
    const size_t part_base = prefix[bounds[lo]], m = prefix[bounds[lo + 1]] - part_base;
    uint32_t *dst = out + 2 * part_base + (prefix[q] - part_base);
    const uint32_t off = cand_off[q], cnt = cand_cnt[q];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      dst[i] = score[off + i];
      dst[m + i] = end[off + i];
    }
  }
}

// After the seed search of an asynchronously enqueued chunk: the host never sees the counts, so the
// conditions it would act on are recorded on the device.  One candidate chunk covers all queries
// iff the total does not exceed the budget -l (aligner.cpp:511-516).
__global__ void async_check_kernel(const unsigned long long *cursor, const unsigned long long *visited,
                                   const int *overflow, unsigned long long max_list_length,
                                   unsigned long long *ctr, uint32_t *flag) {
  if (*cursor > max_list_length || *overflow) flag[0] = 1;
  ctr[1] += *visited;
  ctr[2] += *cursor;
}

__global__ void async_cells_kernel(const unsigned long long *cells, const uint32_t *merge_error,
                                   unsigned long long *ctr, uint32_t *flag) {
  ctr[0] += *cells;
  if (*merge_error) flag[1] = 1;
}

// db_creator.cpp:167-241 on the device: key of every indexable position (or 0xFFFFFFFF).
__global__ void index_keys_kernel(const uint8_t *seq, uint32_t seq_len, const uint32_t *seq_starts,
                                  uint32_t n_seqs, uint32_t seed, uint32_t seed_len, uint32_t *keys,
                                  uint32_t *pos) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < seq_len;
       j += gridDim.x * blockDim.x) {
    uint32_t key = 0;
    bool ok = j + seed_len <= seq_len;
    uint32_t s = seed;
    for (uint32_t i = 0; ok && s != 0; ++i, s >>= 1) {
      const uint8_t c = seq[j + i];
      if (c == kSeqEnd || c == kBaseX) ok = false;        // :198 loop bound, :201-212
      if (s & 1) key = (key << kCharBits) | c;
    }
    if (ok) {
      // :197 only sequences strictly longer than the seed span are indexed
      uint32_t lo = 0, hi = n_seqs;  // largest id with seq_starts[id] <= j
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (seq_starts[mid] <= j) lo = mid; else hi = mid;
      }
      const uint32_t end = (lo + 1 < n_seqs) ? seq_starts[lo + 1] : seq_len;
      if (end - seq_starts[lo] - 1 <= seed_len) ok = false;
    }
    keys[j] = ok ? key : 0xFFFFFFFFu;
    pos[j] = j;
  }
}

__global__ void index_bounds_kernel(const uint32_t *sorted_keys, uint32_t n, uint32_t n_keys,
                                    uint32_t *keys_count) {
  // keys_count[k] = first index i with sorted_keys[i] >= k  (exclusive prefix sums)
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k <= n_keys;
       k += gridDim.x * blockDim.x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sorted_keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    keys_count[k] = lo;
  }
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct DbChunk {
  DevBuf<uint8_t> seq;
  DevBuf<uint32_t> keys_count, positions, seq_starts;
  uint32_t seq_len = 0, keys_count_len = 0, positions_len = 0, n_seqs = 0;
  bool valid = false;
  bool has_index = false;   // false: sequence-only chunk (gm_db_upload_seq), Merge/TraceBack side
  // tiled seed search: per-key tile boundaries inside positions[], built on first use
  DevBuf<uint32_t> split;
  TileGeometry split_geom = {};
  bool split_valid = false;
};

}  // namespace
}  // namespace gm

using namespace gm;

struct gm_context {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  size_t smem_per_sm = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {};
  // asynchronous layer (gm_align_chunk_async / gm_wait)
  std::vector<uint32_t> chunks_since_upload;   // every chunk aligned against the resident queries, in order
  std::vector<cudaEvent_t> async_ev;           // 4 per enqueued chunk: start, search, score, merge
  uint32_t async_open = 0;                     // chunks enqueued and not waited for
  DevBuf<unsigned long long> async_ctr;        // [0] cells [1] index positions [2] candidates (accumulating)
  DevBuf<uint32_t> async_flag;                 // [0] candidate budget / capacity exceeded [1] merge scratch
  uint32_t *h_async = nullptr;                 // pinned read-back of flags and counters
  bool no_sync_upload = false, no_sync_download = false;
  std::vector<uint32_t> upload_keep, upload_keep2;
  bool traced_async = false;                   // a TraceBack was enqueued without reading its counters
  gm_hit *dl_hits = nullptr;                   // host buffers of an enqueued gm_results_download_async
  uint32_t *dl_counts = nullptr;
  bool dl_open = false;

  bool has_opt = false;
  gm_options opt = {};
  uint32_t seed_len = 0, list_len = 0;
  bool use_s32 = false;
  bool wide_scores = false;
  DevBuf<int32_t> matrix;

  DbChunk chunks[GM_MAX_DB_CHUNKS];
  int cur_chunk = -1;   // chunk the resident candidates belong to

  // queries
  DevBuf<uint8_t> queries;
  uint32_t n_queries = 0, query_len = 0;
  DevBuf<uint32_t> run_first, run_last;
  uint32_t n_runs = 0;

  // candidates of (resident queries x cur_chunk)
  uint64_t cand_capacity = 1ull << 26;
  DevBuf<uint32_t> cand_off, cand_cnt, cand_start, cand_score, cand_end;
  DevBuf<uint32_t> staging;
  uint32_t staging_cap = 1u << 15;
  DevBuf<uint16_t> buckets;           // bucket seed search: [grid][tiles][bucket_cap] marks
  DevBuf<uint32_t> fallback;          // queries the bucket kernel hands to the sweep kernel
  DevBuf<uint32_t> tile_emap;         // tile seed search: dense-mode emit bitmaps [grid][words]
  DevBuf<uint32_t> prefix;            // scan output (n_queries + 1)
  DevBuf<uint32_t> bounds;            // gm_candidates_pack part boundaries
  DevBuf<uint32_t> gather0, gather1, gather2;
  DevBuf<uint32_t> strip_scratch;
  DevBuf<unsigned long long> counters;  // [0] cand cursor [1] positions visited [2] cells [3] big cursor
  DevBuf<uint32_t> small;               // [0] query counter [1] task counter [2] overflow [3] n_jobs
                                        // [4] merge error [5] run counter [6] search fallbacks
  uint64_t cand_total = 0;
  std::vector<uint32_t> h_counts;

  // hit lists
  DevBuf<gm_hit> hits[2];
  DevBuf<uint32_t> hit_cnt[2];
  int cur_hits = 0;
  uint32_t cap = 1;
  DevBuf<uint32_t> jobs;
  DevBuf<unsigned long long> big_scratch;
  DevBuf<int> tb_work;
  DevBuf<ChunkRef> chunk_tab;
  bool chunk_tab_dirty = true;
  bool search_fast = true;   // balanced register-resident search kernel when the options allow it
  bool search_bucket = true; // bucket kernel (threshold 2) in front of it
  bool search_hash = false;  // hash kernel (threshold 2) instead of the bucket kernel: variant 3 only
  bool search_tile = true;   // tile kernel (threshold 2) in front of all of them: variant 4, the default
  int tile_test = 0;         // variant 6: the tile kernel's in-place path forced by a tiny staging area (tests)
  bool traceback_fast = true;
  bool deferred = true;      // TraceBack only for the survivors (gm_traceback_pending)
  bool pending = false;      // some resident hit list may hold untraced hits
  uint32_t pending_left = 0; // untraced hits the last gm_traceback_pending could not reach
  uint32_t serial = 0;
  bool imported = false;           // resident candidates came from gm_candidates_import
  uint64_t search_fallbacks = 0;   // queries redone by the sweep kernel (bucket capacities exceeded)
};

namespace {

int check_ctx(gm_context *ctx) {
  if (!ctx) return fail(GM_ERR_ARGUMENT, "null context");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail(GM_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
  return 0;
}

uint32_t seed_length_of(uint32_t seed) {  // index.h:137-147
  uint32_t n = 0;
  for (; seed; seed >>= 1) ++n;
  return n;
}

int ensure_query_state(gm_context *c) {
  if (!c->has_opt) return fail(GM_ERR_ARGUMENT, "gm_set_options must be called first");
  if (c->n_queries == 0) return fail(GM_ERR_ARGUMENT, "no queries resident (gm_query_upload)");
  return 0;
}

int derive_query_options(gm_context *c) {
  // called when both options and queries are known
  if (c->query_len < c->seed_len) return fail(GM_ERR_ARGUMENT, "query shorter than the seed");
  if (c->opt.shift == 0) return fail(GM_ERR_ARGUMENT, "shift must be > 0");
  c->list_len = (c->query_len - c->seed_len) / c->opt.shift + 1;  // aligner.cpp:399
  if (c->list_len > (uint32_t)search_max_list_len())
    return fail(GM_ERR_UNSUPPORTED, "list length %u exceeds %d", c->list_len, search_max_list_len());
  // packed s16 range check for the DPX kernel: every intermediate stays inside
  // [-16384 + open, L * max_score]; fall back to the 32-bit kernel otherwise.
  int hi = 0, lo = 0;
  for (int i = 0; i < 1024; ++i) {
    hi = std::max(hi, c->opt.score_matrix[i]);
    lo = std::min(lo, c->opt.score_matrix[i]);
  }
  const long max_score = (long)c->query_len * hi;
  c->wide_scores = max_score >= 65536;   // Merge packs (score << 16 | index) records otherwise
  c->use_s32 = max_score - c->opt.open_gap > 16000 || lo - c->opt.open_gap < -16000 ||
               c->opt.open_gap < -4000 || c->opt.extend_gap < -4000 || c->opt.open_gap > 0 ||
               c->opt.extend_gap > 0;
  if (c->use_s32 && c->query_len > 1024)
    return fail(GM_ERR_UNSUPPORTED, "score range needs the 32-bit kernel, which is limited to L <= 1024");
  return 0;
}

}  // namespace

// =========================================================================================
// extended API
// =========================================================================================

extern "C" const char *gm_version(void) { return "ghostm_b200 0.1 (sm_100a)"; }
extern "C" const char *gm_last_error(void) { return g_error.c_str(); }

extern "C" int gm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int gm_create(int device, gm_context **out) {
  if (!out) return fail(GM_ERR_ARGUMENT, "null out pointer");
  *out = nullptr;
  GM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(GM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  gm_context *c = new gm_context();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->smem_per_sm = prop.sharedMemPerMultiprocessor;
  GM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto &e : c->ev) GM_CUDA(cudaEventCreate(&e));
  GM_CUDA(c->counters.ensure(8));
  GM_CUDA(c->small.ensure(8));
  GM_CUDA(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
  GM_CUDA(cudaMemsetAsync(c->small.p, 0, 8 * sizeof(uint32_t), c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  *out = c;
  return 0;
}

extern "C" int gm_device_memory(gm_context *c, uint64_t *free_bytes, uint64_t *total_bytes) {
  if (int r = check_ctx(c)) return r;
  size_t f = 0, t = 0;
  GM_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}

extern "C" void gm_destroy(gm_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &ch : c->chunks) {
    ch.seq.release(); ch.keys_count.release(); ch.positions.release(); ch.seq_starts.release();
    ch.split.release();
  }
  c->matrix.release(); c->queries.release(); c->run_first.release(); c->run_last.release();
  c->cand_off.release(); c->cand_cnt.release(); c->cand_start.release(); c->cand_score.release();
  c->cand_end.release(); c->staging.release(); c->prefix.release(); c->gather0.release();
  c->gather1.release(); c->gather2.release(); c->bounds.release(); c->buckets.release(); c->fallback.release(); c->tile_emap.release(); c->strip_scratch.release(); c->counters.release();
  c->small.release(); c->hits[0].release(); c->hits[1].release(); c->hit_cnt[0].release();
  c->hit_cnt[1].release(); c->jobs.release(); c->big_scratch.release(); c->tb_work.release(); c->chunk_tab.release();
  for (auto &e : c->ev) cudaEventDestroy(e);
  for (auto &e : c->async_ev) cudaEventDestroy(e);
  c->async_ctr.release(); c->async_flag.release();
  if (c->h_async) cudaFreeHost(c->h_async);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int gm_set_options(gm_context *c, const gm_options *opt) {
  if (int r = check_ctx(c)) return r;
  if (!opt) return fail(GM_ERR_ARGUMENT, "null options");
  if (opt->seed == 0) return fail(GM_ERR_ARGUMENT, "seed mask is 0");
  if (opt->log_region > 20) return fail(GM_ERR_ARGUMENT, "log_region %u out of range", opt->log_region);
  c->opt = *opt;
  c->seed_len = seed_length_of(opt->seed);
  c->has_opt = true;
  c->cap = std::max<uint32_t>(opt->best, 1);
  GM_CUDA(c->matrix.ensure(1024));
  GM_CUDA(cudaMemcpyAsync(c->matrix.p, opt->score_matrix, 1024 * sizeof(int32_t),
                          cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  if (c->n_queries) {
    for (int b = 0; b < 2; ++b)
      if (c->hits[b].n < (size_t)c->n_queries * c->cap) {
        GM_CUDA(c->hits[0].ensure((size_t)c->n_queries * c->cap));
        GM_CUDA(c->hits[1].ensure((size_t)c->n_queries * c->cap));
        GM_CUDA(c->jobs.ensure((size_t)c->n_queries * c->cap));
        GM_CUDA(cudaMemset(c->hit_cnt[0].p, 0, (size_t)c->n_queries * 4));
        GM_CUDA(cudaMemset(c->hit_cnt[1].p, 0, (size_t)c->n_queries * 4));
        break;
      }
    return derive_query_options(c);
  }
  return 0;
}

extern "C" int gm_set_candidate_capacity(gm_context *c, uint64_t n) {
  if (int r = check_ctx(c)) return r;
  if (n == 0 || n >= (1ull << 32)) return fail(GM_ERR_ARGUMENT, "capacity must be in [1, 2^32)");
  c->cand_capacity = n;
  return 0;
}

// Deferred TraceBack reads the residues of the hit's db chunk: hits still pending when a resident
// chunk is replaced or released are traced first (they could never be finished afterwards).
int trace_before_slot_change(gm_context *c, uint32_t id);

extern "C" int gm_db_upload(gm_context *c, uint32_t id, const uint8_t *seq, uint32_t seq_len,
                            const uint32_t *keys_count, uint32_t keys_count_len,
                            const uint32_t *positions, uint32_t positions_len,
                            const uint32_t *seq_starts, uint32_t n_seqs) {
  if (int r = check_ctx(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS) return fail(GM_ERR_ARGUMENT, "chunk id %u out of range", id);
  if (int r = trace_before_slot_change(c, id)) return r;
  if (!seq || !keys_count || (!positions && positions_len) || !seq_starts || n_seqs == 0)
    return fail(GM_ERR_ARGUMENT, "null db arrays");
  DbChunk &ch = c->chunks[id];
  ch.valid = false;
  GM_CUDA(ch.seq.ensure(seq_len));
  GM_CUDA(ch.keys_count.ensure(keys_count_len));
  GM_CUDA(ch.positions.ensure(positions_len));
  GM_CUDA(ch.seq_starts.ensure(n_seqs));
  GM_CUDA(cudaMemcpyAsync(ch.seq.p, seq, seq_len, cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(ch.keys_count.p, keys_count, (size_t)keys_count_len * 4,
                          cudaMemcpyHostToDevice, c->stream));
  if (positions_len)
    GM_CUDA(cudaMemcpyAsync(ch.positions.p, positions, (size_t)positions_len * 4,
                            cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(ch.seq_starts.p, seq_starts, (size_t)n_seqs * 4, cudaMemcpyHostToDevice,
                          c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  ch.seq_len = seq_len;
  ch.keys_count_len = keys_count_len;
  ch.positions_len = positions_len;
  ch.n_seqs = n_seqs;
  ch.valid = true;
  ch.has_index = true;
  ch.split_valid = false;
  c->chunk_tab_dirty = true;
  if (c->cur_chunk == (int)id) c->cur_chunk = -1;
  return 0;
}

extern "C" int gm_db_upload_seq(gm_context *c, uint32_t id, const uint8_t *seq, uint32_t seq_len,
                                const uint32_t *seq_starts, uint32_t n_seqs) {
  if (int r = check_ctx(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS) return fail(GM_ERR_ARGUMENT, "chunk id %u out of range", id);
  if (int r = trace_before_slot_change(c, id)) return r;
  if (!seq || !seq_starts || n_seqs == 0) return fail(GM_ERR_ARGUMENT, "null db arrays");
  DbChunk &ch = c->chunks[id];
  ch.valid = false;
  ch.keys_count.release();
  ch.positions.release();
  ch.split.release();
  ch.split_valid = false;
  GM_CUDA(ch.seq.ensure(seq_len));
  GM_CUDA(ch.seq_starts.ensure(n_seqs));
  GM_CUDA(cudaMemcpyAsync(ch.seq.p, seq, seq_len, cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(ch.seq_starts.p, seq_starts, (size_t)n_seqs * 4, cudaMemcpyHostToDevice,
                          c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  ch.seq_len = seq_len;
  ch.keys_count_len = 0;
  ch.positions_len = 0;
  ch.n_seqs = n_seqs;
  ch.valid = true;
  ch.has_index = false;
  c->chunk_tab_dirty = true;
  if (c->cur_chunk == (int)id) c->cur_chunk = -1;
  return 0;
}

extern "C" int gm_db_release(gm_context *c, uint32_t id) {
  if (int r = check_ctx(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS) return fail(GM_ERR_ARGUMENT, "chunk id %u out of range", id);
  if (int r = trace_before_slot_change(c, id)) return r;
  DbChunk &ch = c->chunks[id];
  GM_CUDA(cudaStreamSynchronize(c->stream));
  ch.seq.release(); ch.keys_count.release(); ch.positions.release(); ch.seq_starts.release();
  ch.split.release();
  ch.split_valid = false;
  ch.valid = false;
  c->chunk_tab_dirty = true;
  if (c->cur_chunk == (int)id) c->cur_chunk = -1;
  return 0;
}

extern "C" int gm_query_upload(gm_context *c, const uint8_t *seqs, uint32_t n, uint32_t L,
                               const uint8_t *name_break) {
  if (int r = check_ctx(c)) return r;
  if (!seqs || n == 0 || L == 0) return fail(GM_ERR_ARGUMENT, "empty query chunk");
  if (!c->has_opt) return fail(GM_ERR_ARGUMENT, "gm_set_options must be called first");
  c->n_queries = n;
  c->query_len = L;
  if (int r = derive_query_options(c)) { c->n_queries = 0; return r; }
  GM_CUDA(c->queries.ensure((size_t)n * L));
  GM_CUDA(cudaMemcpyAsync(c->queries.p, seqs, (size_t)n * L, cudaMemcpyHostToDevice, c->stream));
  // same-name runs (aligner.cpp:697-700)
  std::vector<uint32_t> first, last;
  for (uint32_t i = 0; i < n; ++i) {
    if (i == 0 || !name_break || name_break[i]) {
      if (i) last.push_back(i - 1);
      first.push_back(i);
    }
  }
  last.push_back(n - 1);
  c->n_runs = (uint32_t)first.size();
  GM_CUDA(c->run_first.ensure(c->n_runs));
  GM_CUDA(c->run_last.ensure(c->n_runs));
  GM_CUDA(cudaMemcpyAsync(c->run_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(c->run_last.p, last.data(), last.size() * 4, cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(c->cand_off.ensure(n));
  GM_CUDA(c->cand_cnt.ensure(n));
  GM_CUDA(c->prefix.ensure((size_t)n + 1));
  for (int b = 0; b < 2; ++b) {
    GM_CUDA(c->hits[b].ensure((size_t)n * c->cap));
    GM_CUDA(c->hit_cnt[b].ensure(n));
    GM_CUDA(cudaMemsetAsync(c->hit_cnt[b].p, 0, (size_t)n * 4, c->stream));
  }
  GM_CUDA(c->jobs.ensure((size_t)n * c->cap));
  GM_CUDA(cudaMemsetAsync(c->cand_cnt.p, 0, (size_t)n * 4, c->stream));
  if (!c->no_sync_upload) GM_CUDA(cudaStreamSynchronize(c->stream));
  else c->upload_keep.swap(first), c->upload_keep2.swap(last);   // the copies read them until the stream gets there
  c->chunks_since_upload.clear();
  c->cur_hits = 0;
  c->cur_chunk = -1;
  c->cand_total = 0;
  c->pending = false;
  c->imported = false;
  return 0;
}

namespace {
// Seed search of all resident queries against chunk id.  async: nothing comes back to the host and
// nothing waits - the candidate total and the overflow flag are checked on the device
// (async_check_kernel); returns 1 when the option set needs a search path with a host decision.
int search_impl(gm_context *c, uint32_t id, uint32_t *counts, uint64_t *total, gm_stats *stats, bool async) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS || !c->chunks[id].valid)
    return fail(GM_ERR_ARGUMENT, "db chunk %u is not resident", id);
  DbChunk &ch = c->chunks[id];
  if (!ch.has_index)
    return fail(GM_ERR_ARGUMENT, "db chunk %u is sequence-only (gm_db_upload_seq): no index to search", id);
  GM_CUDA(c->cand_start.ensure(c->cand_capacity));
  const int grid = c->sm_count;
  GM_CUDA(c->staging.ensure((size_t)grid * 2 * c->staging_cap));  // the fast kernel runs 2 CTAs per SM
  GM_CUDA(cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(unsigned long long), c->stream));
  GM_CUDA(cudaMemsetAsync(c->small.p, 0, 8 * sizeof(uint32_t), c->stream));
  c->cur_chunk = -1;
  if (!async) c->h_counts.assign(c->n_queries, 0);

  if (async && (c->opt.threshold == 0 || c->opt.threshold > 2 * c->list_len ||
                c->opt.threshold > (uint32_t)search_max_threshold()))
    return 1;
  if (c->opt.threshold == 0 || c->opt.threshold > 2 * c->list_len) {
    // aligner.cpp:416: threshold - 1 wraps to UINT_MAX, `count > threshold` is never true; and
    // cnt(d) + cnt(d+1) <= 2 * list_len (every list counts once per region, aligner.cpp:466-476)
    GM_CUDA(cudaMemsetAsync(c->cand_cnt.p, 0, (size_t)c->n_queries * 4, c->stream));
    GM_CUDA(cudaMemsetAsync(c->cand_off.p, 0, (size_t)c->n_queries * 4, c->stream));
    GM_CUDA(cudaStreamSynchronize(c->stream));
  } else if (c->opt.threshold > (uint32_t)search_max_threshold()) {
    return fail(GM_ERR_UNSUPPORTED, "threshold %u > %d with %u lists is not implemented by the "
                "seed-search kernel", c->opt.threshold, search_max_threshold(), c->list_len);
  } else {
    SearchParams p = {};
    p.queries = c->queries.p;
    p.query_len = c->query_len;
    p.n_queries = c->n_queries;
    p.keys_count = ch.keys_count.p;
    p.positions = ch.positions.p;
    p.seed = c->opt.seed;
    p.seed_len = c->seed_len;
    p.shift = c->opt.shift;
    p.log_region = c->opt.log_region;
    p.threshold = c->opt.threshold;
    p.list_len = c->list_len;
    p.n_regions = (ch.seq_len >> c->opt.log_region) + 1;
    const bool fast = search_uses_fast(p.list_len, p.threshold, c->search_fast);
    p.tile_regions = search_tile_regions((int)p.threshold, c->smem_optin, p.n_regions, fast);
    p.cand_off = c->cand_off.p;
    p.cand_cnt = c->cand_cnt.p;
    p.cand_start = c->cand_start.p;
    p.cand_capacity = c->cand_capacity;
    p.cand_cursor = c->counters.p + 0;
    p.staging = c->staging.p;
    p.staging_cap = c->staging_cap;
    p.query_counter = c->small.p + 0;
    p.positions_visited = c->counters.p + 1;
    p.overflow = reinterpret_cast<int *>(c->small.p + 2);
    uint32_t tile_bits = 0, bucket_cap = 0, launches = 1;
    TileGeometry tg = {};
    const bool tile = c->search_tile && c->search_fast && ch.keys_count_len > 1 &&
                      search_tile_geometry(p.threshold, p.list_len, p.shift, p.log_region, ch.seq_len,
                                           ch.keys_count_len - 1, c->smem_per_sm, &tg);
    if (tile) {
      if (!ch.split_valid || memcmp(&ch.split_geom, &tg, sizeof(tg)) != 0) {
        ch.split_valid = false;
        GM_CUDA(ch.split.ensure((size_t)(ch.keys_count_len - 1) * tg.n_tiles + 1));
        GM_CUDA(search_split_build(ch.keys_count.p, ch.keys_count_len - 1, ch.positions.p, tg,
                                   ch.split.p, c->sm_count, c->stream));
        ch.split_geom = tg;
        ch.split_valid = true;
      }
      GM_CUDA(c->staging.ensure((size_t)search_tile_grid(c->sm_count) * c->staging_cap));
      p.staging = c->staging.p;
      // emit bitmaps (absolute region words of the chunk), one per CTA; the kernel leaves them all zero
      const size_t emap_stride = search_tile_emap_words(tg);
      const size_t emap_words = (size_t)search_tile_grid(c->sm_count) * emap_stride;
      if (c->tile_emap.n < emap_words) {
        GM_CUDA(c->tile_emap.ensure(emap_words));
        GM_CUDA(cudaMemsetAsync(c->tile_emap.p, 0, c->tile_emap.n * 4, c->stream));
      }
      p.split = ch.split.p;
      p.tl_emap = c->tile_emap.p;
      p.tl_emap_stride = emap_stride;
      if (c->tile_test >= 2) p.staging_cap = 16;
    }
    const bool hash = !tile && c->search_hash && c->search_fast &&
                      search_hash_ok(p.threshold, p.list_len, p.n_regions);
    const bool bucket = !tile && !hash && c->search_bucket && c->search_fast &&
                        search_bucket_ok(p.threshold, p.list_len, p.n_regions, &tile_bits, &bucket_cap);
    if (async && (hash || bucket)) return 1;   // their over-capacity hand-over is decided on the host
    if (hash) {
      GM_CUDA(c->fallback.ensure(c->n_queries));
      p.fallback_list = c->fallback.p;
      p.fallback_n = c->small.p + 6;
    }
    if (bucket) {
      const uint32_t n_tiles = (p.n_regions + (1u << tile_bits) - 1) >> tile_bits;
      const int bgrid = search_bucket_grid(c->sm_count);
      GM_CUDA(c->buckets.ensure((size_t)bgrid * n_tiles * bucket_cap));
      GM_CUDA(c->staging.ensure((size_t)bgrid * c->staging_cap));
      p.staging = c->staging.p;   // ensure() may have moved it
      GM_CUDA(c->fallback.ensure(c->n_queries));
      p.tile_bits = tile_bits;
      p.bucket_cap = bucket_cap;
      p.buckets = c->buckets.p;
      p.fallback_list = c->fallback.p;
      p.fallback_n = c->small.p + 6;
    }
    GM_CUDA(cudaEventRecord(c->ev[0], c->stream));
    if (tile) {   // no capacity anywhere: staged, else counted and written in place
      GM_CUDA(seed_search_tile_launch(p, tg, c->sm_count, c->stream));
    } else if (bucket || hash) {
      if (hash) GM_CUDA(seed_search_hash_launch(p, c->sm_count, c->stream));
      else GM_CUDA(seed_search_bucket_launch(p, c->sm_count, c->stream));
      uint32_t n_fb = 0;
      GM_CUDA(cudaMemcpyAsync(&n_fb, c->small.p + 6, 4, cudaMemcpyDeviceToHost, c->stream));
      GM_CUDA(cudaStreamSynchronize(c->stream));
      if (getenv("GM_DEBUG_SEARCH")) fprintf(stderr, "[gm_search] %u of %u queries fall back to the sweep kernel\n", n_fb, c->n_queries);
      if (n_fb) {   // queries beyond the bucket kernel's capacities: the sweep kernel, same results
        GM_CUDA(cudaMemsetAsync(c->small.p + 0, 0, 4, c->stream));
        p.query_list = c->fallback.p;
        p.n_queries = n_fb;
        GM_CUDA(seed_search_launch(p, grid, c->stream, true));
        ++launches;
      }
      c->search_fallbacks += n_fb;
    } else {
      GM_CUDA(seed_search_launch(p, grid, c->stream, c->search_fast));
    }
    if (async) {
      async_check_kernel<<<1, 1, 0, c->stream>>>(c->counters.p + 0, c->counters.p + 1,
                                                 reinterpret_cast<const int *>(c->small.p + 2),
                                                 std::min<unsigned long long>(c->opt.max_list_length, c->cand_capacity),
                                                 c->async_ctr.p, c->async_flag.p);
      GM_CUDA(cudaGetLastError());
      c->cur_chunk = (int)id;
      c->imported = false;
      return 0;
    }
    GM_CUDA(cudaEventRecord(c->ev[1], c->stream));
    GM_CUDA(cudaMemcpyAsync(c->h_counts.data(), c->cand_cnt.p, (size_t)c->n_queries * 4,
                            cudaMemcpyDeviceToHost, c->stream));
    GM_CUDA(cudaStreamSynchronize(c->stream));
    int overflow = 0;
    GM_CUDA(cudaMemcpy(&overflow, c->small.p + 2, sizeof(int), cudaMemcpyDeviceToHost));
    if (overflow)
      return fail(GM_ERR_CAPACITY, "candidate buffer (%llu entries) exhausted; raise it with "
                  "gm_set_candidate_capacity", (unsigned long long)c->cand_capacity);
    if (stats) {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
      stats->ms_search += ms;
      stats->kernel_launches += launches;
      unsigned long long v[2];
      GM_CUDA(cudaMemcpy(v, c->counters.p, sizeof(v), cudaMemcpyDeviceToHost));
      stats->seed_positions += v[1];
    }
  }
  uint64_t sum = 0;
  for (uint32_t v : c->h_counts) sum += v;
  c->cand_total = sum;
  c->cur_chunk = (int)id;
  c->imported = false;
  if (counts) memcpy(counts, c->h_counts.data(), (size_t)c->n_queries * 4);
  if (total) *total = sum;
  return 0;
}
}  // namespace

extern "C" int gm_search(gm_context *c, uint32_t id, uint32_t *counts, uint64_t *total,
                         gm_stats *stats) {
  return search_impl(c, id, counts, total, stats, false);
}

extern "C" uint32_t gm_chunk_rule(const uint32_t *counts, uint32_t n, uint32_t first_query,
                                  uint32_t max_list_length, uint64_t *n_candidates, int *last) {
  // aligner.cpp:383-389 and :511-519.  first_query == 0 is the first call of a db chunk; any
  // later chunk starts at the query that overflowed the previous one, whose candidates were
  // carried over without a budget check.
  uint64_t count = 0;
  uint32_t i = first_query;
  int is_last = 0;
  uint32_t end = first_query;
  if (first_query != 0) {
    if (first_query + 1 >= n) {  // overflow on the very last query: its candidates are dropped
      if (n_candidates) *n_candidates = 0;
      if (last) *last = 1;
      return first_query;
    }
    count = counts[first_query];
    i = first_query + 1;
  }
  for (;; ++i) {
    if (i >= n) { end = n; is_last = 1; break; }
    count += counts[i];
    if (count > max_list_length) { end = i; count -= counts[i]; break; }
  }
  if (n_candidates) *n_candidates = count;
  if (last) *last = is_last || count == 0;  // an empty list ends the driver loop (aligner.cpp:136)
  return end;
}

namespace {

int scan_counts(gm_context *c, uint32_t first, uint32_t end, int mode, uint32_t *total) {
  scan_kernel<<<1, 1024, 0, c->stream>>>(c->cand_cnt.p, first, end - first, c->prefix.p, mode);
  GM_CUDA(cudaGetLastError());
  if (total)
    GM_CUDA(cudaMemcpyAsync(total, c->prefix.p + (end - first), 4, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

int check_range(gm_context *c, uint32_t first, uint32_t end, bool need_starts = false) {
  if (int r = ensure_query_state(c)) return r;
  if (c->cur_chunk < 0) return fail(GM_ERR_ARGUMENT, "no searched db chunk (gm_search)");
  if (need_starts && c->imported)
    return fail(GM_ERR_ARGUMENT, "imported candidates carry no region starts (gm_candidates_import)");
  if (first > end || end > c->n_queries) return fail(GM_ERR_ARGUMENT, "bad query range [%u,%u)", first, end);
  return 0;
}

}  // namespace

extern "C" int gm_candidates_download(gm_context *c, uint32_t first, uint32_t end,
                                      uint32_t *query_ids, uint32_t *starts) {
  if (int r = check_ctx(c)) return r;
  if (int r = check_range(c, first, end, true)) return r;
  if (first == end) return 0;
  uint32_t total = 0;
  if (int r = scan_counts(c, first, end, 0, &total)) return r;
  GM_CUDA(cudaStreamSynchronize(c->stream));
  if (total == 0) return 0;
  GM_CUDA(c->gather0.ensure(total));
  GM_CUDA(c->gather2.ensure(total));
  gather_kernel<<<c->sm_count * 4, 128, 0, c->stream>>>(c->cand_off.p, c->cand_cnt.p, c->prefix.p,
                                                        first, end - first, c->cand_start.p, nullptr,
                                                        c->gather0.p, nullptr, c->gather2.p);
  GM_CUDA(cudaGetLastError());
  if (starts) GM_CUDA(cudaMemcpyAsync(starts, c->gather0.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  if (query_ids) GM_CUDA(cudaMemcpyAsync(query_ids, c->gather2.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gm_candidates_pack(gm_context *c, uint32_t n_parts, const uint32_t *bounds,
                                  uint32_t *counts_dev, uint32_t *data_dev,
                                  uint64_t data_capacity_words, uint64_t *part_totals) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (c->cur_chunk < 0) return fail(GM_ERR_ARGUMENT, "no searched db chunk (gm_search)");
  if (n_parts == 0 || !bounds || bounds[0] != 0 || bounds[n_parts] != c->n_queries)
    return fail(GM_ERR_ARGUMENT, "bounds must run from 0 to n_queries");
  uint64_t total = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    if (bounds[p] > bounds[p + 1]) return fail(GM_ERR_ARGUMENT, "bounds must ascend");
    uint64_t m = 0;
    for (uint32_t q = bounds[p]; q < bounds[p + 1]; ++q) m += c->h_counts[q];
    if (part_totals) part_totals[p] = m;
    total += m;
  }
  if (counts_dev)
    GM_CUDA(cudaMemcpyAsync(counts_dev, c->cand_cnt.p, (size_t)c->n_queries * 4,
                            cudaMemcpyDeviceToDevice, c->stream));
  if (data_dev && total) {
    if (2 * total > data_capacity_words)
      return fail(GM_ERR_CAPACITY, "pack buffer holds %llu words, %llu needed",
                  (unsigned long long)data_capacity_words, (unsigned long long)(2 * total));
    if (!c->cand_score.p || !c->cand_end.p) return fail(GM_ERR_ARGUMENT, "candidates are not scored (gm_score)");
    if (int r = scan_counts(c, 0, c->n_queries, 0, nullptr)) return r;
    GM_CUDA(c->bounds.ensure(n_parts + 1));
    GM_CUDA(cudaMemcpyAsync(c->bounds.p, bounds, (size_t)(n_parts + 1) * 4, cudaMemcpyHostToDevice,
                            c->stream));
    pack_kernel<<<c->sm_count * 8, 128, 0, c->stream>>>(c->cand_off.p, c->cand_cnt.p, c->prefix.p,
                                                        c->n_queries, c->bounds.p, n_parts,
                                                        c->cand_score.p, c->cand_end.p, data_dev);
    GM_CUDA(cudaGetLastError());
  }
  GM_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gm_candidates_import(gm_context *c, uint32_t id, const uint32_t *counts_dev,
                                    const uint32_t *data_dev, uint64_t total) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS || !c->chunks[id].valid)
    return fail(GM_ERR_ARGUMENT, "db chunk %u is not resident", id);
  if (!counts_dev || (!data_dev && total)) return fail(GM_ERR_ARGUMENT, "null candidate buffers");
  if (total > c->cand_capacity)
    return fail(GM_ERR_CAPACITY, "candidate buffer (%llu entries) too small for %llu imported; raise "
                "it with gm_set_candidate_capacity", (unsigned long long)c->cand_capacity,
                (unsigned long long)total);
  c->cur_chunk = -1;
  GM_CUDA(c->cand_start.ensure(c->cand_capacity));
  GM_CUDA(c->cand_score.ensure(c->cand_capacity));
  GM_CUDA(c->cand_end.ensure(c->cand_capacity));
  c->h_counts.assign(c->n_queries, 0);
  GM_CUDA(cudaMemcpyAsync(c->cand_cnt.p, counts_dev, (size_t)c->n_queries * 4,
                          cudaMemcpyDeviceToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(c->h_counts.data(), counts_dev, (size_t)c->n_queries * 4,
                          cudaMemcpyDeviceToHost, c->stream));
  if (int r = scan_counts(c, 0, c->n_queries, 0, nullptr)) return r;
  GM_CUDA(cudaMemcpyAsync(c->cand_off.p, c->prefix.p, (size_t)c->n_queries * 4,
                          cudaMemcpyDeviceToDevice, c->stream));
  if (total) {
    GM_CUDA(cudaMemcpyAsync(c->cand_score.p, data_dev, total * 4, cudaMemcpyDeviceToDevice, c->stream));
    GM_CUDA(cudaMemcpyAsync(c->cand_end.p, data_dev + total, total * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  GM_CUDA(cudaStreamSynchronize(c->stream));
  uint64_t sum = 0;
  for (uint32_t v : c->h_counts) sum += v;
  if (sum != total)
    return fail(GM_ERR_ARGUMENT, "imported counts add up to %llu, not %llu", (unsigned long long)sum,
                (unsigned long long)total);
  c->cand_total = sum;
  c->cur_chunk = (int)id;
  c->imported = true;   // no region starts behind the exchange: hits get db_start from TraceBack
  return 0;
}

extern "C" int gm_candidates_transfer(gm_context *src, gm_context *dst, uint32_t id, uint32_t first,
                                      uint32_t end) {
  if (!src || !dst || src == dst) return fail(GM_ERR_ARGUMENT, "two distinct contexts are required");
  if (int r = check_ctx(dst)) return r;
  if (int r = ensure_query_state(dst)) return r;
  if (id >= GM_MAX_DB_CHUNKS || !dst->chunks[id].valid)
    return fail(GM_ERR_ARGUMENT, "db chunk %u is not resident in the receiving context", id);
  if (int r = check_ctx(src)) return r;
  if (int r = check_range(src, first, end, true)) return r;
  if (src->cur_chunk != (int)id) return fail(GM_ERR_ARGUMENT, "the sending context holds chunk %d, not %u", src->cur_chunk, id);
  if (end - first != dst->n_queries)
    return fail(GM_ERR_ARGUMENT, "slice [%u,%u) does not match the %u queries of the receiving context",
                first, end, dst->n_queries);
  if (!src->cand_score.p || !src->cand_end.p) return fail(GM_ERR_ARGUMENT, "candidates are not scored (gm_score)");
  uint64_t total = 0;
  for (uint32_t q = first; q < end; ++q) total += src->h_counts[q];
  if (total > dst->cand_capacity)
    return fail(GM_ERR_CAPACITY, "candidate buffer (%llu entries) too small for %llu transferred; raise "
                "it with gm_set_candidate_capacity", (unsigned long long)dst->cand_capacity,
                (unsigned long long)total);
  // sending side: the slice in reference order (query, region ascending), on its own stream
  if (total) {
    if (int r = scan_counts(src, first, end, 0, nullptr)) return r;
    GM_CUDA(src->gather0.ensure(total));
    GM_CUDA(src->gather1.ensure(total));
    gather_kernel<<<src->sm_count * 4, 128, 0, src->stream>>>(
        src->cand_off.p, src->cand_cnt.p, src->prefix.p, first, end - first, src->cand_score.p,
        src->cand_end.p, src->gather0.p, src->gather1.p, nullptr);
    GM_CUDA(cudaGetLastError());
  }
  // receiving side buffers (allocated under the receiver's device)
  if (int r = check_ctx(dst)) return r;
  dst->cur_chunk = -1;
  GM_CUDA(dst->cand_start.ensure(dst->cand_capacity));
  GM_CUDA(dst->cand_score.ensure(dst->cand_capacity));
  GM_CUDA(dst->cand_end.ensure(dst->cand_capacity));
  if (int r = check_ctx(src)) return r;
  // device to device over NVLink (peer copy; staged by the driver when there is no peer access)
  GM_CUDA(cudaMemcpyPeerAsync(dst->cand_cnt.p, dst->device, src->cand_cnt.p + first, src->device,
                              (size_t)(end - first) * 4, src->stream));
  if (total) {
    GM_CUDA(cudaMemcpyPeerAsync(dst->cand_score.p, dst->device, src->gather0.p, src->device, total * 4, src->stream));
    GM_CUDA(cudaMemcpyPeerAsync(dst->cand_end.p, dst->device, src->gather1.p, src->device, total * 4, src->stream));
  }
  GM_CUDA(cudaStreamSynchronize(src->stream));
  if (int r = check_ctx(dst)) return r;
  dst->h_counts.assign(src->h_counts.begin() + first, src->h_counts.begin() + end);
  if (int r = scan_counts(dst, 0, dst->n_queries, 0, nullptr)) return r;
  GM_CUDA(cudaMemcpyAsync(dst->cand_off.p, dst->prefix.p, (size_t)dst->n_queries * 4,
                          cudaMemcpyDeviceToDevice, dst->stream));
  GM_CUDA(cudaStreamSynchronize(dst->stream));
  dst->cand_total = total;
  dst->cur_chunk = (int)id;
  dst->imported = true;
  return 0;
}

namespace {
int score_impl(gm_context *c, uint32_t first, uint32_t end, uint32_t *scores, uint32_t *ends,
               gm_stats *stats, bool async) {
  if (int r = check_ctx(c)) return r;
  if (int r = check_range(c, first, end, true)) return r;
  if (first == end) return 0;
  DbChunk &ch = c->chunks[c->cur_chunk];
  GM_CUDA(c->cand_score.ensure(c->cand_capacity));
  GM_CUDA(c->cand_end.ensure(c->cand_capacity));
  uint32_t n_strips = 1, pair_strips = 1;
  const int rows = sw_rows_per_strip(c->query_len, &n_strips);
  const int pair_rows = c->use_s32 ? 0 : sw_pair_rows(c->query_len, &pair_strips);
  if (pair_rows) n_strips = pair_strips;
  SwParams p = {};
  p.db = ch.seq.p;
  p.db_len = ch.seq_len;
  p.queries = c->queries.p;
  p.query_len = c->query_len;
  p.first_query = first;
  p.n_q = end - first;
  p.task_prefix = c->prefix.p;
  p.cand_off = c->cand_off.p;
  p.cand_cnt = c->cand_cnt.p;
  p.cand_start = c->cand_start.p;
  p.cand_score = c->cand_score.p;
  p.cand_end = c->cand_end.p;
  p.matrix = c->matrix.p;
  p.open_gap = c->opt.open_gap;
  p.extend_gap = c->opt.extend_gap;
  p.extend = c->opt.extend;
  p.base_len = c->query_len + 2 * c->opt.extend + 2 * (1u << c->opt.log_region);  // aligner.cpp:549
  p.n_strips = n_strips;
  p.task_counter = c->small.p + 1;
  p.cells = c->counters.p + 2;
  GM_CUDA(cudaMemsetAsync(c->small.p + 1, 0, 4, c->stream));
  GM_CUDA(cudaMemsetAsync(c->counters.p + 2, 0, 8, c->stream));
  GM_CUDA(cudaEventRecord(c->ev[0], c->stream));
  uint32_t launches = 0;
  if (!c->use_s32) {
    if (n_strips > 1) {
      const size_t need = (size_t)c->sm_count * 2 * kSwWarps * p.base_len * 96;  // up to 2 CTAs per SM
      GM_CUDA(c->strip_scratch.ensure(need));
    }
    p.strip_scratch = c->strip_scratch.p;
    if (pair_rows) {   // two lanes per candidate pair, 32 candidates per task
      if (int r = scan_counts(c, first, end, 2, nullptr)) return r;
      GM_CUDA(sw_extend_pair_launch(p, pair_rows, c->sm_count, c->stream));
    } else {
      if (int r = scan_counts(c, first, end, 1, nullptr)) return r;
      GM_CUDA(sw_extend_launch(p, rows, c->sm_count, c->stream));
    }
    launches = 2;
  } else {
    GM_CUDA(sw_extend_s32_launch(p, c->sm_count, c->stream));
    launches = 1;
  }
  if (async) {   // the SW cells of this launch join the running sum; nothing is read back
    async_cells_kernel<<<1, 1, 0, c->stream>>>(c->counters.p + 2, c->async_flag.p + 2, c->async_ctr.p,
                                               c->async_flag.p);
    GM_CUDA(cudaGetLastError());
    return 0;
  }
  GM_CUDA(cudaEventRecord(c->ev[1], c->stream));
  if (scores || ends) {
    uint32_t total = 0;
    if (int r = scan_counts(c, first, end, 0, &total)) return r;
    GM_CUDA(cudaStreamSynchronize(c->stream));
    if (total) {
      GM_CUDA(c->gather0.ensure(total));
      GM_CUDA(c->gather1.ensure(total));
      gather_kernel<<<c->sm_count * 4, 128, 0, c->stream>>>(
          c->cand_off.p, c->cand_cnt.p, c->prefix.p, first, end - first, c->cand_score.p,
          c->cand_end.p, c->gather0.p, c->gather1.p, nullptr);
      GM_CUDA(cudaGetLastError());
      if (scores) GM_CUDA(cudaMemcpyAsync(scores, c->gather0.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
      if (ends) GM_CUDA(cudaMemcpyAsync(ends, c->gather1.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
      launches += 2;
    }
  }
  GM_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    stats->ms_score += ms;
    stats->kernel_launches += launches;
    unsigned long long cells = 0;
    GM_CUDA(cudaMemcpy(&cells, c->counters.p + 2, 8, cudaMemcpyDeviceToHost));
    stats->cells += cells;
    uint64_t n = 0;
    for (uint32_t q = first; q < end; ++q) n += c->h_counts[q];
    stats->candidates += n;
  }
  return 0;
}
}  // namespace

extern "C" int gm_score(gm_context *c, uint32_t first, uint32_t end, uint32_t *scores,
                        uint32_t *ends, gm_stats *stats) {
  return score_impl(c, first, end, scores, ends, stats, false);
}

namespace {

int sync_chunk_tab(gm_context *c) {
  if (!c->chunk_tab_dirty) return 0;
  std::vector<ChunkRef> tab(GM_MAX_DB_CHUNKS);
  for (int i = 0; i < GM_MAX_DB_CHUNKS; ++i) {
    tab[i].seq = c->chunks[i].valid ? c->chunks[i].seq.p : nullptr;
    tab[i].seq_starts = c->chunks[i].valid ? c->chunks[i].seq_starts.p : nullptr;
  }
  GM_CUDA(c->chunk_tab.ensure(GM_MAX_DB_CHUNKS));
  GM_CUDA(cudaMemcpyAsync(c->chunk_tab.p, tab.data(), tab.size() * sizeof(ChunkRef),
                          cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  c->chunk_tab_dirty = false;
  return 0;
}

// TraceBack of the jobs queued in c->jobs / small[3] on the hit lists `hits`.
int run_traceback(gm_context *c, gm_hit *hits) {
  if (int r = sync_chunk_tab(c)) return r;
  TracebackParams t = {};
  t.queries = c->queries.p;
  t.query_len = c->query_len;
  t.chunks = c->chunk_tab.p;
  t.hits = hits;
  t.jobs = c->jobs.p;
  t.n_jobs = c->small.p + 3;
  t.matrix = c->matrix.p;
  t.open_gap = c->opt.open_gap;
  t.extend_gap = c->opt.extend_gap;
  t.base_len = c->query_len + 2 * c->opt.extend * 2 * (1u << c->opt.log_region);  // aligner.cpp:775
  // the score range check of the packed SW kernel also covers the 16-bit H of this one
  // (register kernel for L <= 80, warp-cooperative kernel up to L = 1024, else / variant 0 the
  // generic one-thread-per-hit kernel with its global column scratch)
  const bool fast = c->traceback_fast && !c->use_s32 &&
                    (traceback_fast_ok(c->query_len, c->opt.open_gap, c->opt.extend_gap) ||
                     (traceback_warp_ok(c->query_len, c->opt.open_gap, c->opt.extend_gap) &&
                      t.base_len < (1u << 20)));
  if (!fast) {
    const size_t tb_threads = (size_t)traceback_grid(c->sm_count) * traceback_threads();
    GM_CUDA(c->tb_work.ensure(tb_threads * 4 * (c->query_len + 1)));
  }
  t.work = c->tb_work.p;
  GM_CUDA(traceback_launch(t, c->sm_count, c->stream, fast));
  return 0;
}

}  // namespace

extern "C" int gm_set_search_variant(gm_context *c, int fast) {
  if (int r = check_ctx(c)) return r;
  c->search_fast = fast != 0;
  c->search_bucket = fast >= 2;
  c->search_hash = fast == 3;
  c->search_tile = fast >= 4;
  c->tile_test = fast >= 5 ? fast - 4 : 0;   // 6: a 16-entry staging area (5 is the same as 4)
  c->traceback_fast = fast != 0;
  return 0;
}

extern "C" int gm_set_deferred_traceback(gm_context *c, int on) {
  if (int r = check_ctx(c)) return r;
  if (c->pending) return fail(GM_ERR_ARGUMENT, "pending tracebacks: call gm_traceback_pending first");
  c->deferred = on != 0;
  return 0;
}

extern "C" int gm_traceback_pending(gm_context *c, uint64_t *n_done, gm_stats *stats) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (int r = sync_chunk_tab(c)) return r;
  gm_hit *hits = c->hits[c->cur_hits].p;
  GM_CUDA(cudaMemsetAsync(c->small.p + 3, 0, 4, c->stream));
  GM_CUDA(cudaMemsetAsync(c->small.p + 7, 0, 4, c->stream));
  GM_CUDA(cudaEventRecord(c->ev[0], c->stream));
  GM_CUDA(collect_pending_launch(hits, c->hit_cnt[c->cur_hits].p, c->n_queries, c->cap,
                                 c->chunk_tab.p, c->jobs.p, c->small.p + 3, c->small.p + 7,
                                 c->sm_count, c->stream));
  if (int r = run_traceback(c, hits)) return r;
  GM_CUDA(cudaEventRecord(c->ev[1], c->stream));
  if (c->no_sync_download) {   // gm_results_download_async: the counts are looked at in gm_wait
    GM_CUDA(cudaMemcpyAsync(c->h_async + 12, c->small.p + 3, 4, cudaMemcpyDeviceToHost, c->stream));
    GM_CUDA(cudaMemcpyAsync(c->h_async + 13, c->small.p + 7, 4, cudaMemcpyDeviceToHost, c->stream));
    c->pending = false;
    c->traced_async = true;
    return 0;
  }
  uint32_t n = 0, left = 0;
  GM_CUDA(cudaMemcpyAsync(&n, c->small.p + 3, 4, cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaMemcpyAsync(&left, c->small.p + 7, 4, cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  c->pending = left != 0;     // hits whose db chunk is not resident here stay pending
  c->pending_left = left;
  if (n_done) *n_done = n;
  if (stats) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    stats->ms_traceback += ms;
    stats->tracebacks += n;
    stats->kernel_launches += 2;
  }
  return 0;
}

namespace {
int merge_impl(gm_context *c, uint32_t first, uint32_t end, gm_stats *stats, bool async) {
  if (int r = check_ctx(c)) return r;
  if (int r = check_range(c, first, end)) return r;
  DbChunk &ch = c->chunks[c->cur_chunk];
  const int src = c->cur_hits, dst = src ^ 1;
  uint64_t n_new = 0;
  if (async) n_new = std::min<uint64_t>(c->cand_capacity, c->opt.max_list_length);   // the host never saw the counts
  else for (uint32_t q = first; q < end; ++q) n_new += c->h_counts[q];
  const uint64_t big_cap = 2 * (n_new + (uint64_t)c->n_queries * c->cap) + 2;  // records + stopper positions
  if (big_cap > c->big_scratch.n)     // grows in steps of 1.5x: a reallocation stalls every stream of the device
    GM_CUDA(c->big_scratch.ensure(big_cap + big_cap / 2));
  MergeParams p = {};
  p.queries = c->queries.p;
  p.query_len = c->query_len;
  p.n_queries = c->n_queries;
  p.run_first = c->run_first.p;
  p.run_last = c->run_last.p;
  p.n_runs = c->n_runs;
  p.first_query = first;
  p.end_query = end;
  p.cand_off = c->cand_off.p;
  p.cand_cnt = c->cand_cnt.p;
  p.cand_start = c->imported ? nullptr : c->cand_start.p;
  p.cand_score = c->cand_score.p;
  p.cand_end = c->cand_end.p;
  p.db = ch.seq.p;
  p.db_len = ch.seq_len;
  p.seq_starts = ch.seq_starts.p;
  p.n_seqs = ch.n_seqs;
  p.db_chunk = (uint32_t)c->cur_chunk;
  p.old_hits = c->hits[src].p;
  p.old_cnt = c->hit_cnt[src].p;
  p.new_hits = c->hits[dst].p;
  p.new_cnt = c->hit_cnt[dst].p;
  p.cap = c->cap;
  p.best = c->opt.best;
  p.jobs = c->jobs.p;
  p.n_jobs = c->small.p + 3;
  p.big_scratch = c->big_scratch.p;
  p.big_capacity = c->big_scratch.n;
  p.big_cursor = c->counters.p + 3;
  p.error = reinterpret_cast<int *>(c->small.p + 4);
  p.run_counter = c->small.p + 5;
  p.smem_elems = 1536;
  p.wide_scores = c->wide_scores ? 1 : 0;
  p.serial = ++c->serial;
  if (p.serial == kNoId) p.serial = c->serial = 1;
  p.deferred = c->deferred ? 1 : 0;
  GM_CUDA(cudaMemsetAsync(c->small.p + 3, 0, 3 * 4, c->stream));
  GM_CUDA(cudaMemsetAsync(c->counters.p + 3, 0, 8, c->stream));
  GM_CUDA(cudaEventRecord(c->ev[0], c->stream));
  GM_CUDA(merge_launch(p, c->sm_count, c->stream));
  GM_CUDA(cudaEventRecord(c->ev[1], c->stream));

  if (!c->deferred) {
    if (int r = run_traceback(c, c->hits[dst].p)) return r;
  } else {
    c->pending = true;
  }
  if (async) {   // the scratch-exhausted flag becomes sticky on the device; gm_wait looks at it
    async_cells_kernel<<<1, 1, 0, c->stream>>>(c->async_ctr.p + 3, c->small.p + 4, c->async_ctr.p,
                                               c->async_flag.p);
    GM_CUDA(cudaGetLastError());
    c->cur_hits = dst;
    return 0;
  }
  GM_CUDA(cudaEventRecord(c->ev[2], c->stream));
  uint32_t small[8];
  GM_CUDA(cudaMemcpyAsync(small, c->small.p, sizeof(small), cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  if (small[4]) return fail(GM_ERR_CAPACITY, "merge scratch exhausted");
  c->cur_hits = dst;
  if (stats) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    stats->ms_merge += ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
    stats->ms_traceback += ms;
    if (!c->deferred) stats->tracebacks += small[3];
    stats->kernel_launches += c->deferred ? 1 : 2;
    stats->candidate_chunks += 1;
  }
  return 0;
}
}  // namespace

extern "C" int gm_merge(gm_context *c, uint32_t first, uint32_t end, gm_stats *stats) {
  return merge_impl(c, first, end, stats, false);
}

extern "C" int gm_align_prepare(gm_context *c, uint32_t id, gm_stats *stats) {
  uint64_t total = 0;
  if (int r = gm_search(c, id, nullptr, &total, stats)) return r;
  uint32_t first = 0;
  while (true) {
    uint64_t n = 0;
    int last = 0;
    const uint32_t end = gm_chunk_rule(c->h_counts.data(), c->n_queries, first,
                                       c->opt.max_list_length, &n, &last);
    if (n == 0) break;  // aligner.cpp:136-139
    if (int r = gm_score(c, first, end, nullptr, nullptr, stats)) return r;
    if (last) break;
    first = end;
  }
  return 0;
}

extern "C" int gm_align_merge(gm_context *c, gm_stats *stats) {
  if (int r = check_ctx(c)) return r;
  if (c->cur_chunk < 0) return fail(GM_ERR_ARGUMENT, "gm_align_prepare must run first");
  uint32_t first = 0;
  while (true) {
    uint64_t n = 0;
    int last = 0;
    const uint32_t end = gm_chunk_rule(c->h_counts.data(), c->n_queries, first,
                                       c->opt.max_list_length, &n, &last);
    if (n == 0) break;
    if (int r = gm_merge(c, first, end, stats)) return r;
    if (last) break;
    first = end;
  }
  return 0;
}

extern "C" int gm_align_chunk(gm_context *c, uint32_t id, gm_stats *stats) {
  if (c && c->async_open)
    if (int r = gm_wait(c, stats)) return r;
  if (int r = gm_align_prepare(c, id, stats)) return r;
  if (int r = gm_align_merge(c, stats)) return r;
  c->chunks_since_upload.push_back(id);
  return 0;
}

// ---- asynchronous layer -----------------------------------------------------------------------
// gm_align_chunk_async enqueues seed search, SW extension and Merge of one db chunk on the
// context's stream and returns; nothing is copied back and nothing waits.  What the host would
// decide from the per-query counts - whether the candidate budget -l cuts the chunk into several
// Merge calls (aligner.cpp:511-516), whether a buffer overflowed - is recorded on the device and
// looked at once, in gm_wait: if any enqueued chunk needed such a decision, gm_wait redoes the
// whole batch through the synchronous calls (same results, just slower), so the async path never
// changes a hit list.  Option sets whose search or Merge needs the host (bucket / hash kernels,
// best > 16 where an extra Merge call on an unchanged list is not a no-op, immediate TraceBack)
// fall back to the synchronous call inside gm_align_chunk_async itself.
namespace {

int async_state(gm_context *c) {
  if (!c->async_ctr.p) {
    GM_CUDA(c->async_ctr.ensure(4));
    GM_CUDA(c->async_flag.ensure(4));
    GM_CUDA(cudaMemsetAsync(c->async_ctr.p, 0, 4 * sizeof(unsigned long long), c->stream));
    GM_CUDA(cudaMemsetAsync(c->async_flag.p, 0, 4 * sizeof(uint32_t), c->stream));
    GM_CUDA(cudaMallocHost(&c->h_async, 16 * sizeof(uint32_t)));
  }
  return 0;
}

cudaEvent_t async_event(gm_context *c, size_t i) {
  while (c->async_ev.size() <= i) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->async_ev.push_back(e);
  }
  return c->async_ev[i];
}

}  // namespace

extern "C" int gm_align_chunk_async(gm_context *c, uint32_t id) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (int r = async_state(c)) return r;
  const bool eligible = c->deferred && c->opt.best <= 16;
  if (eligible) {
    const size_t e0 = (size_t)c->async_open * 4;
    GM_CUDA(cudaEventRecord(async_event(c, e0), c->stream));
    const int rs = search_impl(c, id, nullptr, nullptr, nullptr, true);
    if (rs < 0) return rs;
    if (rs == 0) {
      GM_CUDA(cudaEventRecord(async_event(c, e0 + 1), c->stream));
      if (int r = score_impl(c, 0, c->n_queries, nullptr, nullptr, nullptr, true)) return r;
      GM_CUDA(cudaEventRecord(async_event(c, e0 + 2), c->stream));
      if (int r = merge_impl(c, 0, c->n_queries, nullptr, true)) return r;
      GM_CUDA(cudaEventRecord(async_event(c, e0 + 3), c->stream));
      c->chunks_since_upload.push_back(id);
      ++c->async_open;
      return 0;
    }
  }
  // not eligible: the synchronous call (drains the stream first, order is kept)
  if (int r = gm_align_chunk(c, id, nullptr)) return r;
  return 0;
}

extern "C" int gm_wait(gm_context *c, gm_stats *stats) {
  if (int r = check_ctx(c)) return r;
  if (c->async_open == 0) {
    GM_CUDA(cudaStreamSynchronize(c->stream));
    const bool traced = c->traced_async;
    c->traced_async = c->dl_open = false;
    if (traced && c->h_async[13])
      return fail(GM_ERR_ARGUMENT, "%u hits cannot be traced back: their db chunk is not resident in this "
                  "context", c->h_async[13]);
    return 0;
  }
  GM_CUDA(cudaMemcpyAsync(c->h_async, c->async_flag.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaMemcpyAsync(c->h_async + 4, c->async_ctr.p, 4 * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaMemsetAsync(c->async_ctr.p, 0, 4 * sizeof(unsigned long long), c->stream));
  GM_CUDA(cudaMemsetAsync(c->async_flag.p, 0, 4 * sizeof(uint32_t), c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  const uint32_t n_async = c->async_open;
  c->async_open = 0;
  const bool redo = c->h_async[0] != 0 || c->h_async[1] != 0;
  if (redo) {
    // some chunk needed a host decision: the whole batch again, synchronously, from empty lists
    const std::vector<uint32_t> chunks = c->chunks_since_upload;
    if (int r = gm_results_clear(c)) return r;
    c->pending = false;
    c->chunks_since_upload.clear();
    for (uint32_t id : chunks)
      if (int r = gm_align_chunk(c, id, stats)) return r;
    const bool dl = c->dl_open;
    c->traced_async = c->dl_open = false;
    if (dl) return gm_results_download(c, c->dl_hits, c->dl_counts);   // the enqueued copy carried the wrong lists
    return 0;
  }
  {
    const bool traced = c->traced_async;
    c->traced_async = c->dl_open = false;
    if (traced && c->h_async[13])
      return fail(GM_ERR_ARGUMENT, "%u hits cannot be traced back: their db chunk is not resident in this "
                  "context", c->h_async[13]);
    if (traced && stats) {
      stats->tracebacks += c->h_async[12];
      stats->kernel_launches += 2;
    }
  }
  if (stats) {
    unsigned long long ctr[4];
    memcpy(ctr, c->h_async + 4, sizeof(ctr));
    stats->cells += ctr[0];
    stats->seed_positions += ctr[1];
    stats->candidates += ctr[2];
    stats->candidate_chunks += n_async;
    stats->kernel_launches += n_async * 7;   // search, check, scan, SW, cells, merge, flag
    for (uint32_t k = 0; k < n_async; ++k) {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->async_ev[k * 4 + 0], c->async_ev[k * 4 + 1]);
      stats->ms_search += ms;
      cudaEventElapsedTime(&ms, c->async_ev[k * 4 + 1], c->async_ev[k * 4 + 2]);
      stats->ms_score += ms;
      cudaEventElapsedTime(&ms, c->async_ev[k * 4 + 2], c->async_ev[k * 4 + 3]);
      stats->ms_merge += ms;
    }
  }
  return 0;
}

// gm_query_upload / gm_results_download without the trailing synchronisation: the copies are
// enqueued (host buffers should be pinned to overlap), completion is gm_wait.  The download first
// runs the deferred TraceBack of the survivors.
extern "C" int gm_query_upload_async(gm_context *c, const uint8_t *seqs, uint32_t n, uint32_t L,
                                     const uint8_t *name_break) {
  if (int r = check_ctx(c)) return r;
  if (c->async_open)
    if (int r = gm_wait(c, nullptr)) return r;   // a new batch replaces the lists the old one still merges into
  c->no_sync_upload = true;
  const int r = gm_query_upload(c, seqs, n, L, name_break);
  c->no_sync_upload = false;
  return r;
}

extern "C" int gm_results_download_async(gm_context *c, gm_hit *hits, uint32_t *counts) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (int r = async_state(c)) return r;
  c->no_sync_download = true;
  const int r = gm_results_download(c, hits, counts);
  c->no_sync_download = false;
  c->dl_hits = hits;
  c->dl_counts = counts;
  c->dl_open = r == 0;
  return r;
}

int trace_before_slot_change(gm_context *c, uint32_t id) {
  if (!c->pending || c->n_queries == 0 || !c->chunks[id].valid) return 0;
  return gm_traceback_pending(c, nullptr, nullptr);
}

extern "C" int gm_results_download(gm_context *c, gm_hit *hits, uint32_t *counts) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (c->pending) {
    if (int r = gm_traceback_pending(c, nullptr, nullptr)) return r;
    if (c->pending)   // never hand out half-finished records (absolute db_end, no alignment columns)
      return fail(GM_ERR_ARGUMENT, "%u hits cannot be traced back: their db chunk is not resident in "
                  "this context (gm_db_upload / gm_db_upload_seq it, or trace before releasing it)",
                  c->pending_left);
  }
  if (hits)
    GM_CUDA(cudaMemcpyAsync(hits, c->hits[c->cur_hits].p, (size_t)c->n_queries * c->cap * sizeof(gm_hit),
                            cudaMemcpyDeviceToHost, c->stream));
  if (counts)
    GM_CUDA(cudaMemcpyAsync(counts, c->hit_cnt[c->cur_hits].p, (size_t)c->n_queries * 4,
                            cudaMemcpyDeviceToHost, c->stream));
  if (!c->no_sync_download) GM_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gm_results_upload(gm_context *c, const gm_hit *hits, const uint32_t *counts) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (!hits || !counts) return fail(GM_ERR_ARGUMENT, "null hit lists");
  GM_CUDA(cudaMemcpyAsync(c->hits[c->cur_hits].p, hits, (size_t)c->n_queries * c->cap * sizeof(gm_hit),
                          cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(c->hit_cnt[c->cur_hits].p, counts, (size_t)c->n_queries * 4,
                          cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  c->pending = true;  // uploaded lists may carry untraced hits of chunks resident here
  return 0;
}

extern "C" int gm_results_clear(gm_context *c) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (c->async_open) {     // enqueued chunks still merge into the lists: finish them (their result is dropped)
    GM_CUDA(cudaStreamSynchronize(c->stream));
    GM_CUDA(cudaMemsetAsync(c->async_ctr.p, 0, 4 * sizeof(unsigned long long), c->stream));
    GM_CUDA(cudaMemsetAsync(c->async_flag.p, 0, 4 * sizeof(uint32_t), c->stream));
    c->async_open = 0;
    c->traced_async = c->dl_open = false;
  }
  GM_CUDA(cudaMemsetAsync(c->hit_cnt[c->cur_hits].p, 0, (size_t)c->n_queries * 4, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  c->chunks_since_upload.clear();     // the lists are empty again: nothing to redo
  c->pending = false;
  return 0;
}

extern "C" int gm_results_device(gm_context *c, void **hits, void **counts) {
  if (int r = check_ctx(c)) return r;
  if (int r = ensure_query_state(c)) return r;
  if (hits) *hits = c->hits[c->cur_hits].p;
  if (counts) *counts = c->hit_cnt[c->cur_hits].p;
  c->pending = true;  // the caller may overwrite the lists (peer-to-peer receive)
  return 0;
}

extern "C" void *gm_stream(gm_context *c) { return c ? (void *)c->stream : nullptr; }

namespace gm_peak {
__global__ void __launch_bounds__(1024) dpx_peak_kernel(uint32_t *out, uint32_t a, uint32_t b, int iters) {
  uint32_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 2654435761u + i * b;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __viaddmax_s16x2(acc[i], a, b);
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
}  // namespace gm_peak

extern "C" int gm_measure_dpx_peak(gm_context *c, double *lane_instr_per_s) {
  if (int r = check_ctx(c)) return r;
  if (!lane_instr_per_s) return fail(GM_ERR_ARGUMENT, "null output");
  DevBuf<uint32_t> out;
  GM_CUDA(out.ensure((size_t)c->sm_count * 1024));
  const int iters = 1 << 15;
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    GM_CUDA(cudaEventRecord(c->ev[3], c->stream));
    gm_peak::dpx_peak_kernel<<<c->sm_count, 1024, 0, c->stream>>>(out.p, 3u, 0x00050007u, iters);
    GM_CUDA(cudaGetLastError());
    GM_CUDA(cudaEventRecord(c->ev[4], c->stream));
    GM_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0;
    GM_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]));
    const double rate = (double)iters * 8 * 1024 * c->sm_count / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  out.release();
  *lane_instr_per_s = best;
  return 0;
}

extern "C" int gm_db_build_index(gm_context *c, uint32_t id, const uint8_t *seq, uint32_t seq_len,
                                 const uint32_t *seq_starts, uint32_t n_seqs, uint32_t seed) {
  if (int r = check_ctx(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS) return fail(GM_ERR_ARGUMENT, "chunk id %u out of range", id);
  if (!seq || !seq_starts || n_seqs == 0 || seed == 0) return fail(GM_ERR_ARGUMENT, "bad index input");
  DbChunk &ch = c->chunks[id];
  ch.valid = false;
  uint32_t weight = 0;
  for (uint32_t s = seed; s; s >>= 1) weight += s & 1;
  if (weight > 6) return fail(GM_ERR_UNSUPPORTED, "seed weight %u too large", weight);
  const uint32_t n_keys = 1u << (kCharBits * weight);
  GM_CUDA(ch.seq.ensure(seq_len));
  GM_CUDA(ch.seq_starts.ensure(n_seqs));
  GM_CUDA(ch.keys_count.ensure((size_t)n_keys + 1));
  GM_CUDA(ch.positions.ensure(seq_len));
  GM_CUDA(cudaMemcpyAsync(ch.seq.p, seq, seq_len, cudaMemcpyHostToDevice, c->stream));
  GM_CUDA(cudaMemcpyAsync(ch.seq_starts.p, seq_starts, (size_t)n_seqs * 4, cudaMemcpyHostToDevice, c->stream));
  // db_creator.cpp:167-241: keys of all offsets, then the stable counting sort of index_build.cu
  // (hand-written, 8-bit passes over one bit more than the key width: the 0xFFFFFFFF keys of
  // non-indexable offsets end up last), then the CSR boundaries
  DevBuf<uint32_t> keys, keys_tmp, pos, pos_tmp, scratch;
  GM_CUDA(keys.ensure(seq_len));
  GM_CUDA(keys_tmp.ensure(seq_len));
  GM_CUDA(pos.ensure(seq_len));
  GM_CUDA(pos_tmp.ensure(seq_len));
  GM_CUDA(scratch.ensure(index_sort_scratch_words(seq_len)));
  index_keys_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(ch.seq.p, seq_len, ch.seq_starts.p, n_seqs,
                                                            seed, seed_length_of(seed), keys.p, pos.p);
  GM_CUDA(cudaGetLastError());
  uint32_t *keys_sorted = nullptr;
  GM_CUDA(index_sort_pairs(keys.p, pos.p, keys_tmp.p, pos_tmp.p, ch.positions.p, seq_len,
                           kCharBits * weight + 1, scratch.p, &keys_sorted, nullptr, c->stream));
  index_bounds_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(keys_sorted, seq_len, n_keys,
                                                              ch.keys_count.p);
  GM_CUDA(cudaGetLastError());
  uint32_t n_pos = 0;
  GM_CUDA(cudaMemcpyAsync(&n_pos, ch.keys_count.p + n_keys, 4, cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  keys.release(); keys_tmp.release(); pos.release(); pos_tmp.release(); scratch.release();
  ch.seq_len = seq_len;
  ch.n_seqs = n_seqs;
  ch.keys_count_len = n_keys + 1;
  ch.positions_len = n_pos;
  ch.valid = true;
  ch.has_index = true;
  c->chunk_tab_dirty = true;
  if (c->cur_chunk == (int)id) c->cur_chunk = -1;
  return 0;
}

extern "C" int gm_db_download_index(gm_context *c, uint32_t id, uint32_t *keys_count,
                                    uint32_t *positions, uint32_t *positions_len) {
  if (int r = check_ctx(c)) return r;
  if (id >= GM_MAX_DB_CHUNKS || !c->chunks[id].valid)
    return fail(GM_ERR_ARGUMENT, "db chunk %u is not resident", id);
  DbChunk &ch = c->chunks[id];
  if (keys_count)
    GM_CUDA(cudaMemcpyAsync(keys_count, ch.keys_count.p, (size_t)ch.keys_count_len * 4,
                            cudaMemcpyDeviceToHost, c->stream));
  if (positions)
    GM_CUDA(cudaMemcpyAsync(positions, ch.positions.p, (size_t)ch.positions_len * 4,
                            cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  if (positions_len) *positions_len = ch.positions_len;
  return 0;
}

// =========================================================================================
// legacy drop-in: reference aligner_gpu.h:32-117 on one process-wide context
// =========================================================================================

namespace {

gm_context *g_legacy = nullptr;
gm_options g_legacy_opt;
bool g_legacy_have_matrix = false;
int g_legacy_device = 0;
std::vector<uint32_t> g_legacy_counts;   // per-query counts of the current db chunk
bool g_legacy_searched = false;
uint32_t g_legacy_first = 0, g_legacy_end = 0;
uint32_t g_legacy_n_seqs_dummy[1] = {0};

void legacy_die(const char *what) {  // aligner_gpu.h:135-141
  fprintf(stderr, "Cuda error in ghostm_b200 (%s) : %s.\n", what, gm_last_error());
  exit(EXIT_FAILURE);
}

void legacy_ensure() {
  if (!g_legacy && gm_create(g_legacy_device, &g_legacy) != 0) legacy_die("gm_create");
}

}  // namespace

extern "C" int InitGpu(void) { return 0; }

extern "C" size_t GetNeededGPUMemorySize(uint32_t seed, uint32_t shift_size, uint32_t max_list_length,
                                         uint32_t max_query_length, uint32_t max_number_queries,
                                         uint32_t max_db_length) {
  (void)shift_size;
  uint32_t weight = 0;
  for (uint32_t s = seed; s; s >>= 1) weight += s & 1;
  size_t bytes = 0;
  bytes += (size_t)max_db_length * 5;                              // residues + positions
  bytes += ((size_t)1 << (kCharBits * weight)) * 4 + 4;            // keys_count
  bytes += (size_t)max_query_length * max_number_queries;          // queries
  bytes += (size_t)max_number_queries * 4 * 3;                     // per-query counts/offsets/prefix
  bytes += (size_t)std::max<uint32_t>(max_list_length, 1u << 20) * 4 * 3;  // starts, scores, ends
  bytes += (size_t)64 << 20;                                       // staging + slack
  return bytes;
}

extern "C" int CheckGpuMemory(uint32_t seed, uint32_t shift_size, uint32_t max_list_length,
                              uint32_t max_query_length, uint32_t max_number_queries,
                              uint32_t max_db_length) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  return GetNeededGPUMemorySize(seed, shift_size, max_list_length, max_query_length,
                                max_number_queries, max_db_length) > free_b;
}

extern "C" int SetOptionGpu(uint32_t max_list_length, int score_matrix[], int device) {
  g_legacy_device = device;
  if (g_legacy && g_legacy->device != device) { gm_destroy(g_legacy); g_legacy = nullptr; }
  legacy_ensure();
  memset(&g_legacy_opt, 0, sizeof(g_legacy_opt));
  g_legacy_opt.max_list_length = max_list_length;
  memcpy(g_legacy_opt.score_matrix, score_matrix, sizeof(g_legacy_opt.score_matrix));
  g_legacy_have_matrix = true;
  // the candidate store must hold every candidate of one (query chunk, db chunk) pair
  const uint64_t cap = std::min<uint64_t>(std::max<uint64_t>((uint64_t)max_list_length * 2, 1u << 20),
                                          (1ull << 32) - 1);
  if (gm_set_candidate_capacity(g_legacy, cap) != 0) legacy_die("gm_set_candidate_capacity");
  return 0;
}

extern "C" void printGpuInfo(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return;
  printf("  GPU [%d] %s, %d SMs, %.1f GiB, sm_%d%d (ghostm_b200)\n", device, prop.name,
         prop.multiProcessorCount, prop.totalGlobalMem / 1073741824.0, prop.major, prop.minor);
}

namespace {
// queries are kept on the host until the search options arrive with SearchNextGpu
std::vector<uint8_t> g_legacy_queries;
uint32_t g_legacy_nq = 0, g_legacy_ql = 0;
bool g_legacy_queries_dirty = false;
}  // namespace

extern "C" int SetQueryGpu(uint8_t sequences[], uint32_t number_sequences, uint32_t sequence_length) {
  legacy_ensure();
  g_legacy_queries.assign(sequences, sequences + (size_t)number_sequences * sequence_length);
  g_legacy_nq = number_sequences;
  g_legacy_ql = sequence_length;
  g_legacy_queries_dirty = true;
  return 0;
}

extern "C" int SetDbGpu(uint8_t sequences[], uint32_t sequences_length, uint32_t keys_count[],
                        uint32_t keys_count_length, uint32_t positions[], uint32_t positions_length) {
  legacy_ensure();
  // the legacy boundary does not pass the .pos table; Merge stays on the caller's side
  if (gm_db_upload(g_legacy, 0, sequences, sequences_length, keys_count, keys_count_length, positions,
                   positions_length, g_legacy_n_seqs_dummy, 1) != 0)
    legacy_die("SetDbGpu");
  g_legacy_searched = false;
  return 0;
}

extern "C" uint32_t SearchNextGpu(uint32_t query_sequence_length, uint32_t number_query_sequences,
                                  uint32_t seed, uint32_t threshold, uint32_t shift_size,
                                  uint32_t log_region_size, uint32_t max_number_alignments,
                                  uint32_t start_query_id, uint32_t *alignment_count_list,
                                  uint32_t *starts) {
  legacy_ensure();
  if (start_query_id >= number_query_sequences) return 0;               // aligner_gpu.cu:787
  if (start_query_id == 0 || !g_legacy_searched) {                      // aligner_gpu.cu:814
    g_legacy_opt.seed = seed;
    g_legacy_opt.threshold = threshold;
    g_legacy_opt.shift = shift_size;
    g_legacy_opt.log_region = log_region_size;
    g_legacy_opt.max_list_length = max_number_alignments;
    g_legacy_opt.best = 1;
    g_legacy_opt.open_gap = -11;   // refined by CalculateScoreGpu, which receives the real values
    g_legacy_opt.extend_gap = -1;
    if (gm_set_options(g_legacy, &g_legacy_opt) != 0) legacy_die("SearchNextGpu/options");
    if (g_legacy_queries_dirty || g_legacy->n_queries != number_query_sequences ||
        g_legacy->query_len != query_sequence_length) {
      if (g_legacy_nq != number_query_sequences || g_legacy_ql != query_sequence_length) {
        g_error = "SearchNextGpu: query shape differs from SetQueryGpu";
        legacy_die("SearchNextGpu");
      }
      if (gm_query_upload(g_legacy, g_legacy_queries.data(), g_legacy_nq, g_legacy_ql, nullptr) != 0)
        legacy_die("SearchNextGpu/queries");
      g_legacy_queries_dirty = false;
    }
    g_legacy_counts.assign(number_query_sequences, 0);
    if (gm_search(g_legacy, 0, g_legacy_counts.data(), nullptr, nullptr) != 0)
      legacy_die("SearchNextGpu/search");
    g_legacy_searched = true;
  }
  // The reference caller advances start_query_id by the returned count (aligner.cpp:376), so
  // the overflowing query is simply the first query of the next call; its candidates are
  // counted against the next budget exactly like the CPU path's carried list.
  uint64_t n = 0;
  int last = 0;
  const uint32_t end = gm_chunk_rule(g_legacy_counts.data(), number_query_sequences, start_query_id,
                                     max_number_alignments, &n, &last);
  if (n == 0) return 0;
  if (n > max_number_alignments) {  // one query alone exceeds the caller's starts[] buffer
    g_error = "SearchNextGpu: a single query has more candidates than max_number_alignments";
    legacy_die("SearchNextGpu");
  }
  g_legacy_first = start_query_id;
  g_legacy_end = end;
  uint32_t cum = 0;
  alignment_count_list[0] = 0;
  for (uint32_t q = start_query_id; q < end; ++q) {
    cum += g_legacy_counts[q];
    alignment_count_list[q - start_query_id + 1] = cum;
  }
  if (gm_candidates_download(g_legacy, start_query_id, end, nullptr, starts) != 0)
    legacy_die("SearchNextGpu/download");
  return end - start_query_id;
}

extern "C" void CalculateScoreGpu(uint32_t db_length, uint32_t query_sequence_length,
                                  uint32_t number_alignment_list, uint32_t scores[], uint32_t ends[],
                                  uint32_t base_search_length, uint32_t offset, int open_gap,
                                  int extend_gap) {
  (void)db_length;
  (void)query_sequence_length;
  legacy_ensure();
  if (number_alignment_list == 0) return;
  // base_search_length = L + 2*extend + 2*2^r (aligner.cpp:527) is re-derived from the options
  (void)base_search_length;
  g_legacy_opt.extend = offset;
  g_legacy_opt.open_gap = open_gap;
  g_legacy_opt.extend_gap = extend_gap;
  if (gm_set_options(g_legacy, &g_legacy_opt) != 0) legacy_die("CalculateScoreGpu/options");
  if (gm_score(g_legacy, g_legacy_first, g_legacy_end, scores, ends, nullptr) != 0)
    legacy_die("CalculateScoreGpu");
}

extern "C" int FreeGpu(void) {
  if (g_legacy) gm_destroy(g_legacy);
  g_legacy = nullptr;
  g_legacy_searched = false;
  return 0;
}
