/*
 * ghostm_b200 - C ABI of the B200-native `ghostm aln` search hot path.
 *
 * Plain C, plain pointers and sizes.  Two layers:
 *
 *  (1) LEGACY DROP-IN: the ten `extern "C"` symbols of the reference's GPU boundary,
 *      reference aligner_gpu.h:32-117, with identical names, argument meaning and
 *      return conventions.  The reference's own aligner.cpp links against this
 *      library unchanged; `ghostm aln -D <device>` then runs
 *      its search and scoring stages on the B200 kernels.
 *
 *  (2) EXTENDED API (gm_*): handle based, re-entrant per context, keeps db chunks,
 *      queries, candidates and hit lists resident in HBM, and adds the stages the
 *      reference leaves on the host (Merge aligner.cpp:687-769, TraceBack :771-949).
 *
 * All functions are synchronous unless stated.  Errors: gm_* return 0 on success and
 * a negative gm_status otherwise, gm_last_error() gives the text; the legacy symbols
 * follow the reference (CUDA failure -> message on stderr and exit(EXIT_FAILURE),
 * aligner_gpu.h:135-141).  There is no CPU fallback anywhere behind this header.
 */
#ifndef GHOSTM_B200_H_
#define GHOSTM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * (1) Legacy drop-in: reference aligner_gpu.h:32-117
 * ---------------------------------------------------------------------------------- */

/* aligner_gpu.h:36 - called once before anything else (aligner.cpp:77). Returns 0. */
int InitGpu(void);

/* aligner_gpu.h:38-46 - bytes of device memory the given limits need. */
size_t GetNeededGPUMemorySize(uint32_t seed, uint32_t shift_size, uint32_t max_list_length,
                              uint32_t max_query_length, uint32_t max_number_queries,
                              uint32_t max_db_length);

/* aligner_gpu.h:48-56 - 1 = does not fit (caller throws "Out of GPU memory",
 * aligner.cpp:81-84), 0 = fits. */
int CheckGpuMemory(uint32_t seed, uint32_t shift_size, uint32_t max_list_length,
                   uint32_t max_query_length, uint32_t max_number_queries,
                   uint32_t max_db_length);

/* aligner_gpu.h:58-63 - candidate budget (-l), 32x32 score matrix (row = db residue
 * code, column = query residue code), CUDA device ordinal (-D). */
int SetOptionGpu(uint32_t max_list_length, int score_matrix[], int device);

/* aligner_gpu.h:65 */
void printGpuInfo(int device);

/* aligner_gpu.h:67-72 - queries: uint8[number_sequences][sequence_length], X padded. */
int SetQueryGpu(uint8_t sequences[], uint32_t number_sequences, uint32_t sequence_length);

/* aligner_gpu.h:74-82 - one db chunk: residues with SEQUENCE_END separators, CSR k-mer
 * index (keys_count[32^w + 1], positions[]). */
int SetDbGpu(uint8_t sequences[], uint32_t sequences_length, uint32_t keys_count[],
             uint32_t keys_count_length, uint32_t positions[], uint32_t positions_length);

/* aligner_gpu.h:84-96 - next candidate chunk starting at query start_query_id.
 * Returns the number of queries covered (0 = done).  alignment_count_list[0..n] receives
 * the exclusive prefix sums of the per-query candidate counts, starts[] the candidate db
 * offsets, queries ascending, regions ascending inside a query.  The chunk cut follows
 * the reference CPU rule (aligner.cpp:511-516), which is the oracle; see DESIGN.md. */
uint32_t SearchNextGpu(uint32_t query_sequence_length, uint32_t number_query_sequences,
                       uint32_t seed, uint32_t threshold, uint32_t shift_size,
                       uint32_t log_region_size, uint32_t max_number_alignments,
                       uint32_t start_query_id, uint32_t *alignment_count_list,
                       uint32_t *starts);

/* aligner_gpu.h:98-110 - scores the candidates of the preceding SearchNextGpu call
 * (device state, like the reference).  Gap penalties are negative. */
void CalculateScoreGpu(uint32_t db_length, uint32_t query_sequence_length,
                       uint32_t number_alignment_list, uint32_t scores[], uint32_t ends[],
                       uint32_t base_search_length, uint32_t offset, int open_gap,
                       int extend_gap);

/* aligner_gpu.h:112 */
int FreeGpu(void);

/* ------------------------------------------------------------------------------------
 * (2) Extended API
 * ---------------------------------------------------------------------------------- */

typedef struct gm_context gm_context;

typedef enum {
  GM_OK = 0,
  GM_ERR_CUDA = -1,          /* a CUDA runtime call failed */
  GM_ERR_ARGUMENT = -2,      /* invalid argument or call order */
  GM_ERR_UNSUPPORTED = -3,   /* option outside what the kernels implement */
  GM_ERR_CAPACITY = -4       /* a device buffer limit was hit (never silently truncated) */
} gm_status;

/* AlignerOption (aligner.h:41-60) restricted to what the device path needs. */
typedef struct {
  uint32_t seed;             /* Index seed mask (index.h:51), e.g. 0xF for -k 4 */
  uint32_t shift;            /* -s, aligner.cpp:309-311 */
  uint32_t log_region;       /* log2 of -r, aligner.cpp:302-307 */
  uint32_t threshold;        /* -t */
  uint32_t extend;           /* -e */
  uint32_t best;             /* -b */
  uint32_t max_list_length;  /* -l in candidates, aligner.cpp:291-293 */
  int32_t open_gap;          /* negative, aligner.cpp:277-279 */
  int32_t extend_gap;        /* negative, aligner.cpp:273-275 */
  int32_t score_matrix[32 * 32];
} gm_options;

/* Alignment (alignment.h:35-146) as a POD record; db_chunk replaces the name string. */
typedef struct {
  uint32_t query_id;
  uint32_t db_id;            /* sequence index inside db_chunk */
  uint32_t db_chunk;
  uint32_t score;
  uint32_t db_start;         /* relative to the sequence start after Merge */
  uint32_t db_end;
  uint32_t aln_len;
  uint32_t aln_match;
  float seq_id;
} gm_hit;

/* Per-call counters of gm_align_chunk / gm_search. */
typedef struct {
  uint64_t candidates;       /* candidates scored */
  uint64_t cells;            /* SW cells = sum over candidates of L x clipped window */
  uint64_t seed_positions;   /* index positions visited by the seed search */
  uint64_t tracebacks;       /* hits sent through TraceBack */
  uint32_t candidate_chunks; /* Merge calls (aligner.cpp:131-171 iterations) */
  uint32_t kernel_launches;  /* kernels launched by this call */
  float ms_search, ms_score, ms_merge, ms_traceback; /* CUDA-event times */
} gm_stats;

const char *gm_version(void);
const char *gm_last_error(void);

int gm_device_count(void);
/* Free and total bytes of the context's device (drivers decide with it whether all db chunks stay
 * resident or are streamed chunk by chunk like the reference does, aligner.cpp:115-173). */
int gm_device_memory(gm_context *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
int gm_create(int device, gm_context **ctx);
void gm_destroy(gm_context *ctx);

int gm_set_options(gm_context *ctx, const gm_options *opt);

/* Capacity (entries) of the device candidate store: it holds every candidate of one
 * (query chunk, db chunk) pair, i.e. possibly several candidate chunks of max_list_length.
 * Exceeding it is reported as GM_ERR_CAPACITY, never truncated.  Default 2^26. */
int gm_set_candidate_capacity(gm_context *ctx, uint64_t n_candidates);

/* Upload a db chunk and keep it resident under slot `chunk_id` (0..GM_MAX_DB_CHUNKS-1).
 * seq_starts/n_seqs are the .pos file (db.cpp:64-77), needed by Merge (DB::GetID). */
#define GM_MAX_DB_CHUNKS 256
int gm_db_upload(gm_context *ctx, uint32_t chunk_id, const uint8_t *seq, uint32_t seq_len,
                 const uint32_t *keys_count, uint32_t keys_count_len, const uint32_t *positions,
                 uint32_t positions_len, const uint32_t *seq_starts, uint32_t n_seqs);
int gm_db_release(gm_context *ctx, uint32_t chunk_id);

/* Upload one query chunk; name_break[i] != 0 iff the name of query i differs from that of
 * query i-1 (may be NULL = all names distinct).  Clears the hit lists. */
int gm_query_upload(gm_context *ctx, const uint8_t *seqs, uint32_t n_queries, uint32_t query_len,
                    const uint8_t *name_break);

/* The whole path for the resident queries against one resident db chunk: seed search,
 * candidate chunking, SW extension, Merge and TraceBack, everything on the device.
 * Chunks must be presented in ascending db order (Merge carries state, aligner.cpp:114-174). */
int gm_align_chunk(gm_context *ctx, uint32_t chunk_id, gm_stats *stats);

/* Asynchronous layer: the same work enqueued on the context's stream, completion = gm_wait.
 *   gm_query_upload_async      H2D of a query chunk (pin the host buffer to overlap; keep it alive
 *                              until gm_wait); waits for a previous batch that is still open
 *   gm_align_chunk_async       seed search + SW extension + Merge of one db chunk, no host round trip:
 *                              the decisions the host takes from the per-query counts (candidate budget
 *                              -l, buffer limits; aligner.cpp:511-516) are recorded on the device
 *   gm_results_download_async  deferred TraceBack of the survivors + D2H of the hit lists
 *   gm_wait                    synchronises; if any enqueued chunk needed a host decision the whole
 *                              batch is redone through the synchronous calls (identical results) and an
 *                              enqueued download is repeated; fills stats (may be NULL)
 * Chunks must still be presented in ascending db order. */
int gm_query_upload_async(gm_context *ctx, const uint8_t *seqs, uint32_t n_queries, uint32_t query_len,
                          const uint8_t *name_break);
int gm_align_chunk_async(gm_context *ctx, uint32_t chunk_id);
int gm_results_download_async(gm_context *ctx, gm_hit *hits, uint32_t *counts);
int gm_wait(gm_context *ctx, gm_stats *stats);

/* gm_align_chunk in two halves, for pipelines that receive the carried hit lists from another
 * GPU while this one is already searching: prepare = seed search + SW extension of every
 * candidate chunk (independent of the hit lists), merge = the Merge (+TraceBack) calls. */
int gm_align_prepare(gm_context *ctx, uint32_t chunk_id, gm_stats *stats);
int gm_align_merge(gm_context *ctx, gm_stats *stats);

/* Hit lists of the resident query chunk: hits[n_queries][cap] with cap = max(best,1),
 * counts[n_queries].  Lists live at the last query of each same-name run (aligner.cpp:701). */
int gm_results_download(gm_context *ctx, gm_hit *hits, uint32_t *counts);
/* Replace the device hit lists (used to hand carried lists from one GPU to the next). */
int gm_results_upload(gm_context *ctx, const gm_hit *hits, const uint32_t *counts);

/* TraceBack scheduling.  Selection in Merge never reads TraceBack's output (the dedupe key is
 * DB::GetID of the forward end, aligner.cpp:706-709), so by default (on = 1) TraceBack is
 * deferred: hits stay "pending" (absolute db_end, aln_match == 0xFFFFFFFF) until
 * gm_traceback_pending runs it once for the survivors whose db chunk is resident in this
 * context.  gm_results_download calls it implicitly.  on = 0 restores the reference order
 * (TraceBack inside every Merge). */
int gm_set_deferred_traceback(gm_context *ctx, int on);
/* Kernel variants: 4 (default) = tiled seed-search kernel (per-key tile boundaries in the index,
 * one-pass detection in a 31+1-bit occupancy bitmap) when threshold == 2 and list_len <= 64, else
 * as 2; 2 = bucket seed-search kernel under the same conditions (queries beyond its capacities are
 * redone by the sweep kernel), else as 1;
 * 3 = hash seed-search kernel under the same conditions (an independent algorithm, ~20 % slower);
 * 1 = balanced register-resident sweep kernel whenever list_len <= 64 and register-resident
 * TraceBack whenever L <= 80, else the generic kernels; 0 = always the generic kernels (tests).
 * 5 = as 4; 6 = variant 4 with a 16-entry staging area, so that every query takes the rare
 * "count first, then write straight into the output" path (tests).  All variants produce identical
 * candidates. */
int gm_set_search_variant(gm_context *ctx, int variant);
int gm_traceback_pending(gm_context *ctx, uint64_t *n_done, gm_stats *stats);
/* Empty the device hit lists without re-uploading the queries. */
int gm_results_clear(gm_context *ctx);
/* Device addresses of the current hit lists (gm_hit[n_queries][cap], uint32[n_queries]) for
 * peer-to-peer exchange between GPUs (NCCL send/recv on these buffers); valid until the next
 * gm_merge / gm_align_chunk / gm_query_upload. */
int gm_results_device(gm_context *ctx, void **hits, void **counts);
/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on. */
void *gm_stream(gm_context *ctx);
/* INT/DPX issue-rate roofline: runs a register-only VIADDMNMX.S16x2 loop on every SM and
 * returns lane-instructions per second (x4 = packed add+max operations per second). */
int gm_measure_dpx_peak(gm_context *ctx, double *lane_instr_per_s);

/* ---- stage level (parity tests, the legacy symbols, multi-GPU drivers) ---- */

/* Seed search of all resident queries against chunk_id: per-query candidate counts and the
 * candidate db offsets stay on the device.  counts may be NULL. */
int gm_search(gm_context *ctx, uint32_t chunk_id, uint32_t *counts, uint64_t *total,
              gm_stats *stats);
/* Reference chunk rule (aligner.cpp:383-389,511-519) on host counts: given the first query
 * of a chunk, returns the one-past-last query whose candidates belong to it and whether the
 * driver loop ends (the reference stops on an empty list, aligner.cpp:136-139). */
uint32_t gm_chunk_rule(const uint32_t *counts, uint32_t n_queries, uint32_t first_query,
                       uint32_t max_list_length, uint64_t *n_candidates, int *last);
/* Candidates of queries [first_query, end_query) in reference order. */
int gm_candidates_download(gm_context *ctx, uint32_t first_query, uint32_t end_query,
                           uint32_t *query_ids, uint32_t *starts);
/* SW extension of the candidates of queries [first_query, end_query); results stay on the
 * device; scores/ends may be NULL, else they receive them in reference order. */
int gm_score(gm_context *ctx, uint32_t first_query, uint32_t end_query, uint32_t *scores,
             uint32_t *ends, gm_stats *stats);
/* Merge + TraceBack of the scored candidates of [first_query, end_query) into the hit lists. */
int gm_merge(gm_context *ctx, uint32_t first_query, uint32_t end_query, gm_stats *stats);

/* ---- db-sharded runs: front (search + SW, one db chunk per GPU) / back (Merge + TraceBack of
 * one query slice per GPU) with one exchange of scored candidates between them.  The reference
 * has no counterpart (single GPU); what is preserved is its Merge order: every query slice sees
 * the db chunks in ascending order with exactly the candidates Aligner::Execute would hand to
 * Merge (aligner.cpp:131-171). ---- */

/* Sequence-only db chunk for the back side: residues + .pos table (DB::GetID, TraceBack
 * windows), no index.  gm_search on such a chunk fails with GM_ERR_ARGUMENT. */
int gm_db_upload_seq(gm_context *ctx, uint32_t chunk_id, const uint8_t *seq, uint32_t seq_len,
                     const uint32_t *seq_starts, uint32_t n_seqs);
/* Pack the scored candidates of the searched chunk by query slice.  bounds[n_parts + 1] are
 * ascending query indices (bounds[0] = 0, bounds[n_parts] = n_queries).  DEVICE outputs:
 * counts_dev[n_queries] = per-query candidate counts; data_dev = for every part p a block
 * [score[m_p] | end[m_p]] in reference order (query, then region, ascending), blocks in part
 * order, 2 * sum(m_p) words in total (data_capacity_words is checked).  The region starts stay
 * behind: Merge only copies them into the hits and TraceBack overwrites them (aligner.cpp:941).
 * HOST output: part_totals[n_parts] = m_p.  data_dev may be NULL (totals only). */
int gm_candidates_pack(gm_context *ctx, uint32_t n_parts, const uint32_t *bounds,
                       uint32_t *counts_dev, uint32_t *data_dev, uint64_t data_capacity_words,
                       uint64_t *part_totals);
/* Install scored candidates of db chunk chunk_id (resident, possibly sequence-only) for ALL
 * resident queries from DEVICE buffers laid out like one gm_candidates_pack part: counts_dev
 * [n_queries] and data_dev = [score[total] | end[total]].  Afterwards gm_merge behaves as after
 * gm_search + gm_score on that chunk (hits carry db_start = 0 until TraceBack has run). */
int gm_candidates_import(gm_context *ctx, uint32_t chunk_id, const uint32_t *counts_dev,
                         const uint32_t *data_dev, uint64_t total);

/* Single process, several devices: pack + exchange + import in one call.  The scored candidates
 * of src's searched chunk chunk_id for queries [first_query, end_query) go to dst, whose resident
 * queries must be exactly that slice; device to device (cudaMemcpyPeerAsync: NVLink when peer
 * access exists).  Not thread safe per context: serialise calls that share a src or a dst. */
int gm_candidates_transfer(gm_context *src, gm_context *dst, uint32_t chunk_id,
                           uint32_t first_query, uint32_t end_query);

/* Device-side synthetic data and index construction (bench only; db_creator.cpp:167-241). */
int gm_db_build_index(gm_context *ctx, uint32_t chunk_id, const uint8_t *seq, uint32_t seq_len,
                      const uint32_t *seq_starts, uint32_t n_seqs, uint32_t seed);
int gm_db_download_index(gm_context *ctx, uint32_t chunk_id, uint32_t *keys_count,
                         uint32_t *positions, uint32_t *positions_len);

#ifdef __cplusplus
}
#endif
#endif /* GHOSTM_B200_H_ */
