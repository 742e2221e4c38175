// Shared declarations of the ghostm_b200 device code (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ghostm_b200.h"

namespace gm {

constexpr int kAlphabet = 32;     // reference common.h:31
constexpr int kCharBits = 5;      // common.h:32
constexpr uint8_t kSeqEnd = 25;   // common.h:34
constexpr uint8_t kBaseX = 23;    // common.h:35
constexpr uint32_t kNoId = 0xFFFFFFFFu;

// Opt-in dynamic shared memory of a kernel: always raised to the device maximum (minus the kernel's
// static part), never to the size of one launch - several contexts launch the same kernels from
// different host threads with different sizes, and a per-launch value would race.
template <typename K>
inline cudaError_t allow_max_dynamic_smem(K kern) {
  cudaFuncAttributes attr;
  cudaError_t err = cudaFuncGetAttributes(&attr, kern);
  if (err != cudaSuccess) return err;
  int dev = 0, optin = 0;
  if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
  if ((err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return err;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)attr.sharedSizeBytes);
}

// ---- SW extension ---------------------------------------------------------------------
constexpr int kSwThreads = 256;          // 8 warps per CTA, one CTA per SM (register bound)
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwCandPerTask = 64;       // one warp task = 64 candidates of ONE query (2 per lane)
constexpr int kSwMaxRows = 80;           // rows (query residues) held in registers per strip

struct SwParams {
  const uint8_t *db;
  uint32_t db_len;
  const uint8_t *queries;      // [n_queries][query_len]
  uint32_t query_len;
  uint32_t first_query;        // tasks cover queries [first_query, first_query + n_q)
  uint32_t n_q;
  const uint32_t *task_prefix; // [n_q + 1] exclusive prefix of ceil(cnt/64)
  const uint32_t *cand_off;    // [n_queries] first candidate of each query in cand_* arrays
  const uint32_t *cand_cnt;    // [n_queries]
  const uint32_t *cand_start;  // candidate db offsets (region starts)
  uint32_t *cand_score;        // out
  uint32_t *cand_end;          // out
  const int32_t *matrix;       // [32*32], row = db residue, column = query residue
  int open_gap, extend_gap;    // negative
  uint32_t extend;             // -e
  uint32_t base_len;           // L + 2*extend + 2*2^r  (aligner.cpp:549)
  uint32_t n_strips;           // ceil(query_len / R)
  uint32_t *task_counter;      // dynamic task fetch
  uint32_t *strip_scratch;     // boundary rows between strips (n_strips > 1)
  unsigned long long *cells;   // += L * clipped window per candidate
};

// ---- seed search ----------------------------------------------------------------------
struct SearchParams {
  const uint8_t *queries;
  uint32_t query_len, n_queries;
  const uint32_t *keys_count;
  const uint32_t *positions;
  uint32_t seed, seed_len, shift, log_region, threshold, list_len;
  uint32_t n_regions;          // regions of the chunk: (seq_len >> log_region) + 1
  uint32_t tile_regions;       // regions per shared-memory tile (multiple of 1024)
  uint32_t *cand_off, *cand_cnt;
  uint32_t *cand_start;
  unsigned long long cand_capacity;
  unsigned long long *cand_cursor;   // global append cursor
  uint32_t *staging;           // [gridDim.x][staging_cap]
  uint32_t staging_cap;
  uint32_t *query_counter;     // dynamic query fetch
  unsigned long long *positions_visited;
  int *overflow;               // set when cand_capacity is exceeded
  const uint32_t *query_list;  // sweep kernels: process query_list[0..n_queries) instead of 0..n_queries
  // bucket kernel (threshold 2)
  uint32_t tile_bits;          // log2 regions per tile
  uint32_t bucket_cap;         // marks per tile bucket
  uint16_t *buckets;           // [gridDim.x][n_tiles][bucket_cap], L2-resident scratch
  uint32_t *fallback_list;     // queries that exceed a capacity, redone by the sweep kernel
  uint32_t *fallback_n;
  // tile kernel (threshold 2, seed_search_tile.cu)
  const uint32_t *split;       // [n_keys * tl_tiles + 1] per-key tile boundaries inside positions[]
  uint32_t tl_nw, tl_hc;       // bitmap words per tile (31 regions each) + carried halo words
  uint32_t tl_tiles;
  uint32_t tl_ring;            // words of the shared-memory summary ring (power of two)
  uint32_t *tl_emap;           // [gridDim.x][tl_emap_stride] emit bitmaps (absolute region words), all zero
  size_t tl_emap_stride;
};

// Geometry of the tiled seed search for one db chunk (search_tile_geometry).
struct TileGeometry {
  uint32_t nw, hc, n_tiles;
  uint32_t tile_pos;           // positions per tile: (31 * nw) << log_region
  uint32_t ring;               // summary ring words
};

// ---- merge / traceback --------------------------------------------------------------
struct MergeParams {
  // queries / runs
  const uint8_t *queries;
  uint32_t query_len, n_queries;
  const uint32_t *run_first;   // [n_runs] first query of each same-name run
  const uint32_t *run_last;    // [n_runs] last query of the run (where the list lives)
  uint32_t n_runs;
  uint32_t first_query, end_query;   // queries with new candidates in this call
  // scored candidates
  const uint32_t *cand_off, *cand_cnt, *cand_start, *cand_score, *cand_end;
  // db chunk
  const uint8_t *db;
  uint32_t db_len;
  const uint32_t *seq_starts;
  uint32_t n_seqs, db_chunk;
  // hit lists (old -> new), cap records per query
  const gm_hit *old_hits;
  const uint32_t *old_cnt;
  gm_hit *new_hits;
  uint32_t *new_cnt;
  uint32_t cap, best;
  // traceback job queue
  uint32_t *jobs;              // slot index = query * cap + k
  uint32_t *n_jobs;
  // overflow area for runs that do not fit the shared-memory list
  unsigned long long *big_scratch;
  unsigned long long big_capacity;
  unsigned long long *big_cursor;
  int *error;                  // set on scratch exhaustion
  uint32_t *run_counter;
  uint32_t smem_elems;         // list capacity per warp in shared memory (u32 records)
  uint32_t serial;             // id of this Merge call: marks hits accepted by it
  int deferred;                // 1: TraceBack runs later on the survivors (no job queue here)
  int wide_scores;             // 1: a score may need more than 16 bits: 64-bit sort records only
};

// A hit whose TraceBack has not run yet carries aln_match == kNoId, aln_len == serial of the
// Merge call that accepted it, db_start == candidate region start and the ABSOLUTE db_end.
struct ChunkRef {
  const uint8_t *seq;
  const uint32_t *seq_starts;
};

struct TracebackParams {
  const uint8_t *queries;
  uint32_t query_len;
  const ChunkRef *chunks;      // [GM_MAX_DB_CHUNKS]; seq == nullptr: chunk not resident here
  gm_hit *hits;
  const uint32_t *jobs;
  const uint32_t *n_jobs;
  const int32_t *matrix;
  int open_gap, extend_gap;
  uint32_t base_len;           // L + 2*extend*2*2^r  (aligner.cpp:775)
  int *work;                   // [threads][4][L+1] column state
};


}  // namespace gm
