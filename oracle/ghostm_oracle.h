/*
 * TEST INFRASTRUCTURE ONLY - the oracle is the checker, never the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under ghostm_b200/ includes or links this file.
 *
 * ghostm_oracle: plain-C, single-threaded restatement of the `ghostm aln` search
 * hot path of the reference (jakewendt/ghostm), function by function.  Every
 * function cites the reference file:line it follows.  Parity is PINNED: the test
 * suite checks this restatement against
 *   - the README known answer (README.rdoc:138-149, tests/golden/readme_*.txt),
 *   - stage dumps of the unmodified reference built by oracle/Makefile
 *     (oracle/_ref/ref_probe; fixtures under tests/golden/, generator
 *     tests/golden/make_golden.py),
 *   - and, where oracle/_ref is present, fresh reference runs on seeded inputs.
 */
#ifndef GHOSTM_ORACLE_H_
#define GHOSTM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMO_ALPHABET_SIZE 32 /* common.h:31 */
#define GMO_CHARACTER_SIZE 5 /* common.h:32 */
#define GMO_SEQUENCE_END 25  /* common.h:34 */
#define GMO_BASE_X 23        /* common.h:35 */

/* index.h:86-101, :137-161 */
uint32_t gmo_get_key(const uint8_t *sequence, uint32_t seed);
uint32_t gmo_seed_length(uint32_t seed);
uint32_t gmo_seed_weight(uint32_t seed);

/* db_creator.cpp:167-241: counting-sort index of one db chunk.  keys_count must hold
 * 32^weight + 1 entries, positions at least seq_len entries; returns positions_len. */
uint32_t gmo_build_index(const uint8_t *seq, uint32_t seq_len, const uint32_t *seq_starts,
                         uint32_t n_seqs, uint32_t seed, uint32_t *keys_count, uint32_t *positions);

/* aligner.cpp:418-509 - the region-count filter for ONE query.  Writes at most cap
 * starts (db offsets, ascending) and returns the full count. */
uint64_t gmo_search_query(const uint8_t *query, uint32_t query_len, uint32_t seed, uint32_t shift,
                          uint32_t log_region, uint32_t threshold, const uint32_t *keys_count,
                          const uint32_t *positions, uint32_t *starts, uint64_t cap);

/* aligner.cpp:383-521 - SearchNextCpu including the candidate-chunk rule and its carry
 * state (Aligner::next_query_id_, next_alignment_list_).  The state object lives across
 * the calls of one (query chunk, db chunk) pair. */
typedef struct gmo_search_state gmo_search_state;
gmo_search_state *gmo_search_begin(void);
void gmo_search_free(gmo_search_state *st);
/* Returns the number of candidates of this call (0 = done); *query_ids / *starts point
 * to buffers owned by the state, valid until the next call. */
uint64_t gmo_search_next(gmo_search_state *st, const uint8_t *queries, uint32_t n_queries,
                         uint32_t query_len, uint32_t seed, uint32_t shift, uint32_t log_region,
                         uint32_t threshold, uint32_t max_list_length, const uint32_t *keys_count,
                         const uint32_t *positions, const uint32_t **query_ids,
                         const uint32_t **starts);

/* aligner.cpp:545-685 - CalculateScoreCpu over a candidate list. */
void gmo_calculate_score(const uint8_t *db, uint32_t db_len, const uint8_t *queries,
                         uint32_t query_len, uint64_t n, const uint32_t *query_ids,
                         const uint32_t *starts, const int *score_matrix, int open_gap,
                         int extend_gap, uint32_t extend, uint32_t log_region, uint32_t *scores,
                         uint32_t *ends);

/* aligner.cpp:771-949 - TraceBack of one accepted hit. */
void gmo_traceback(const uint8_t *db, const uint8_t *query, uint32_t query_len, uint32_t db_end,
                   const int *score_matrix, int open_gap, int extend_gap, uint32_t extend,
                   uint32_t log_region, uint32_t *db_start, uint32_t *aln_len,
                   uint32_t *aln_match, float *seq_id);

/* db.h:94-120 - DB::GetID. */
uint32_t gmo_db_get_id(const uint32_t *seq_starts, uint32_t n_seqs, uint32_t seq_len,
                       uint32_t position);

/* alignment.h:35-146 as a plain record.  db_id == UINT32_MAX marks "not merged yet"
 * (Alignment() default, alignment.h).  db_chunk is ours: it replaces the db_name string. */
typedef struct {
  uint32_t query_id;
  uint32_t db_id;
  uint32_t db_chunk;
  uint32_t score;
  uint32_t db_start;
  uint32_t db_end;
  uint32_t aln_len;
  uint32_t aln_match;
  float seq_id;
} gmo_hit;

/* libstdc++ 13 std::sort (bits/stl_algo.h: introsort + final insertion sort) with the
 * reference comparator (aligner.cpp:52-63: score descending), restated for gmo_hit. */
void gmo_std_sort_hits(gmo_hit *first, size_t n);

/* aligner.cpp:687-769 - Merge of one scored candidate chunk into result lists.
 * results: n_queries x result_cap records, result_counts[n_queries]; name_break[i] != 0
 * iff query i's name differs from query i-1's (name_break[0] ignored). */
void gmo_merge(gmo_hit *results, uint32_t *result_counts, uint32_t result_cap, uint64_t n,
               const uint32_t *query_ids, const uint32_t *starts, const uint32_t *scores,
               const uint32_t *ends, const uint8_t *queries, uint32_t n_queries,
               uint32_t query_len, const uint8_t *name_break, const uint8_t *db, uint32_t db_len,
               const uint32_t *seq_starts, uint32_t n_seqs, uint32_t db_chunk,
               const int *score_matrix, int open_gap, int extend_gap, uint32_t extend,
               uint32_t log_region, uint32_t best);

/* score_matrix_reader.cpp:44-113 with the built-in BLOSUM62 text (:41-42) -> int[32*32]. */
void gmo_blosum62(int *matrix);

/* aligner.cpp:951-976 + statistics.cpp:40-59 - one output row of WriteOutput (style 0),
 * formatted like the reference's ostream (default %g precision 6).  Returns bytes written. */
int gmo_format_row(char *buf, size_t buflen, const char *query_name, const char *db_name,
                   const gmo_hit *hit, uint32_t query_length_no_x, uint64_t db_length,
                   float lambda, float K);
/* aligner.cpp:956-963 - query length with trailing X removed. */
uint32_t gmo_query_length(const uint8_t *query, uint32_t query_len);

#ifdef __cplusplus
}
#endif
#endif /* GHOSTM_ORACLE_H_ */
