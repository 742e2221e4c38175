/*
 * TEST INFRASTRUCTURE ONLY - the oracle is the checker, never the product.
 * See ghostm_oracle.h for the contract and for how parity is pinned.
 *
 * Plain C, single thread, written to follow the reference loop by loop so that a
 * reader can hold the two side by side.  Speed is irrelevant here.
 */
#include "ghostm_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GMO_UINT_MAX 0xFFFFFFFFu

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) {
    fprintf(stderr, "ghostm_oracle: out of memory (%zu bytes)\n", n);
    abort();
  }
  return p;
}

static void *xrealloc(void *q, size_t n) {
  void *p = realloc(q, n ? n : 1);
  if (!p) {
    fprintf(stderr, "ghostm_oracle: out of memory (%zu bytes)\n", n);
    abort();
  }
  return p;
}

/* ------------------------------------------------------------------ index.h */

/* index.h:86-101  Index::GetKey */
uint32_t gmo_get_key(const uint8_t *sequence, uint32_t seed) {
  uint32_t i, s, key;
  for (i = 0, s = seed, key = 0; s != 0; ++i, s >>= 1) {
    if (s & 1) {
      key = key << GMO_CHARACTER_SIZE;
      key = key | sequence[i];
    }
  }
  return key;
}

/* index.h:137-147  Index::GetSeedLength */
uint32_t gmo_seed_length(uint32_t seed) {
  uint32_t length = 0;
  for (; seed != 0; seed >>= 1) ++length;
  return length;
}

/* index.h:149-161  Index::GetSeedWeight */
uint32_t gmo_seed_weight(uint32_t seed) {
  uint32_t weight = 0;
  for (; seed != 0; seed >>= 1)
    if (seed & 1) ++weight;
  return weight;
}

/* db_creator.cpp:167-241  DBCreator::ConstructIndex */
uint32_t gmo_build_index(const uint8_t *seq, uint32_t seq_len, const uint32_t *seq_starts,
                         uint32_t n_seqs, uint32_t seed, uint32_t *keys_count,
                         uint32_t *positions) {
  uint32_t seed_length = gmo_seed_length(seed);
  uint32_t seed_weight = gmo_seed_weight(seed);
  uint32_t keys_count_length = 1;
  uint32_t i, j, k;
  uint32_t *keys = (uint32_t *)xmalloc((size_t)seq_len * sizeof(uint32_t));
  uint32_t *counts;
  for (i = 0; i < seed_weight; ++i) keys_count_length *= GMO_ALPHABET_SIZE;
  keys_count_length += 1;
  for (i = 0; i < seq_len; ++i) keys[i] = GMO_UINT_MAX;
  for (i = 0; i < keys_count_length; ++i) keys_count[i] = 0;

  for (i = 0; i < n_seqs; ++i) {
    uint32_t end = (i + 1 < n_seqs) ? seq_starts[i + 1] : seq_len;
    uint32_t length = end - seq_starts[i] - 1; /* residues, without the END separator */
    if (length > seed_length) { /* db_creator.cpp:197: strictly longer than the seed */
      for (j = seq_starts[i]; seq[j + seed_length - 1] != GMO_SEQUENCE_END; ++j) {
        int contain_x = 0;
        for (k = 0; k < seed_length; ++k)
          if (seq[j + k] == GMO_BASE_X) contain_x = 1;
        if (!contain_x) {
          uint32_t key = gmo_get_key(seq + j, seed);
          keys[j] = key;
          ++keys_count[key + 1];
        }
      }
    }
  }
  for (i = 1; i < keys_count_length; ++i) keys_count[i] = keys_count[i - 1] + keys_count[i];

  counts = (uint32_t *)xmalloc((size_t)keys_count_length * sizeof(uint32_t));
  memset(counts, 0, (size_t)keys_count_length * sizeof(uint32_t));
  for (i = 0; i < n_seqs; ++i) {
    for (j = seq_starts[i]; seq[j] != GMO_SEQUENCE_END; ++j) {
      uint32_t key = keys[j];
      if (key != GMO_UINT_MAX) {
        positions[keys_count[key] + counts[key]] = j;
        ++counts[key];
      }
    }
  }
  free(keys);
  free(counts);
  return keys_count[keys_count_length - 1];
}

/* ------------------------------------------------------- seed search (a1-a4) */

/* aligner.cpp:418-509: the body of the per-query loop of SearchNextCpu.  The variable
 * names are the reference's. */
uint64_t gmo_search_query(const uint8_t *query, uint32_t query_len, uint32_t seed, uint32_t shift,
                          uint32_t log_region, uint32_t threshold_option,
                          const uint32_t *keys_count, const uint32_t *positions, uint32_t *starts,
                          uint64_t cap) {
  uint32_t seed_length = gmo_seed_length(seed);
  uint32_t list_length = (query_len - seed_length) / shift + 1; /* aligner.cpp:399 */
  uint32_t *distance_list = (uint32_t *)xmalloc(list_length * sizeof(uint32_t));
  const uint32_t **positions_list =
      (const uint32_t **)xmalloc(list_length * sizeof(const uint32_t *));
  uint32_t *positions_length_list = (uint32_t *)xmalloc(list_length * sizeof(uint32_t));
  uint32_t *id_list = (uint32_t *)xmalloc(list_length * sizeof(uint32_t));
  uint32_t threshold = threshold_option - 1; /* aligner.cpp:416 */
  uint32_t d, distance, count, next_distance, next_count, j, k;
  uint64_t emitted = 0;

  for (j = 0; j < list_length; ++j) { /* aligner.cpp:422-444 */
    uint32_t key = gmo_get_key(&query[j * shift], seed);
    positions_length_list[j] = keys_count[key + 1] - keys_count[key]; /* index.h:105-114 */
    positions_list[j] = &positions[keys_count[key]];
    for (k = 0; k < positions_length_list[j] && positions_list[j][k] < j * shift; ++k)
      ;
    id_list[j] = k;
    distance_list[j] = GMO_UINT_MAX;
    if (k < positions_length_list[j]) {
      distance_list[j] = (positions_list[j][k] - j * shift) >> log_region;
      ++id_list[j];
    }
  }

  distance = 0;
  count = 0;
  while (1) { /* aligner.cpp:453-498 */
    next_count = 0;
    next_distance = distance_list[0];
    for (j = 1; j < list_length; ++j)
      if (distance_list[j] < next_distance) next_distance = distance_list[j];
    if (next_distance == GMO_UINT_MAX) break;

    for (j = 0; j < list_length; ++j) {
      if (next_distance == distance_list[j]) {
        ++next_count;
        distance_list[j] = GMO_UINT_MAX;
        for (k = id_list[j]; k < positions_length_list[j]; ++k) {
          d = (positions_list[j][k] - j * shift) >> log_region;
          if (d != next_distance) {
            distance_list[j] = d;
            id_list[j] = k + 1;
            break;
          }
        }
      }
    }

    if ((next_distance - distance) == 1) count += next_count;
    if (count > threshold) {
      distance = distance << log_region;
      if (emitted < cap) starts[emitted] = distance;
      ++emitted;
    }
    count = next_count;
    distance = next_distance;
  }
  if (count > threshold) { /* aligner.cpp:501-508 "last check" */
    distance = distance << log_region;
    if (emitted < cap) starts[emitted] = distance;
    ++emitted;
  }

  free(distance_list);
  free(positions_list);
  free(positions_length_list);
  free(id_list);
  return emitted;
}

struct gmo_search_state {
  uint32_t next_query_id; /* Aligner::next_query_id_ (aligner.h:71) */
  /* Aligner::next_alignment_list_ (aligner.h:72) */
  uint32_t *next_ids, *next_starts;
  uint64_t next_n, next_cap;
  /* alignment_list of the current call */
  uint32_t *ids, *starts;
  uint64_t n, cap;
};

gmo_search_state *gmo_search_begin(void) {
  gmo_search_state *st = (gmo_search_state *)xmalloc(sizeof(*st));
  memset(st, 0, sizeof(*st)); /* aligner.cpp:127-128: next_query_id_ = 0, list cleared */
  return st;
}

void gmo_search_free(gmo_search_state *st) {
  if (!st) return;
  free(st->next_ids);
  free(st->next_starts);
  free(st->ids);
  free(st->starts);
  free(st);
}

static void append(uint32_t **ids, uint32_t **starts, uint64_t *n, uint64_t *cap,
                   const uint32_t *src_ids, const uint32_t *src_starts, uint64_t m) {
  if (*n + m > *cap) {
    uint64_t c = *cap ? *cap : 1024;
    while (c < *n + m) c *= 2;
    *ids = (uint32_t *)xrealloc(*ids, c * sizeof(uint32_t));
    *starts = (uint32_t *)xrealloc(*starts, c * sizeof(uint32_t));
    *cap = c;
  }
  memcpy(*ids + *n, src_ids, m * sizeof(uint32_t));
  memcpy(*starts + *n, src_starts, m * sizeof(uint32_t));
  *n += m;
}

/* aligner.cpp:383-521  Aligner::SearchNextCpu */
uint64_t gmo_search_next(gmo_search_state *st, const uint8_t *queries, uint32_t n_queries,
                         uint32_t query_len, uint32_t seed, uint32_t shift, uint32_t log_region,
                         uint32_t threshold, uint32_t max_list_length, const uint32_t *keys_count,
                         const uint32_t *positions, const uint32_t **query_ids,
                         const uint32_t **starts) {
  uint64_t alignment_count;
  uint32_t i;
  st->n = 0; /* aligner.cpp:348-350 (SearchNext clears the list) */
  *query_ids = NULL;
  *starts = NULL;
  if (st->next_query_id == n_queries) return 0; /* aligner.cpp:385-386 */
  /* aligner.cpp:388-389: the carried candidates of the overflowing query come first */
  append(&st->ids, &st->starts, &st->n, &st->cap, st->next_ids, st->next_starts, st->next_n);
  st->next_n = 0;
  alignment_count = st->n;

  for (i = st->next_query_id; i < n_queries; ++i) {
    uint64_t c, cap = 1024, k;
    uint32_t *tmp = (uint32_t *)xmalloc(cap * sizeof(uint32_t));
    c = gmo_search_query(queries + (size_t)i * query_len, query_len, seed, shift, log_region,
                         threshold, keys_count, positions, tmp, cap);
    if (c > cap) {
      cap = c;
      tmp = (uint32_t *)xrealloc(tmp, cap * sizeof(uint32_t));
      gmo_search_query(queries + (size_t)i * query_len, query_len, seed, shift, log_region,
                       threshold, keys_count, positions, tmp, cap);
    }
    { /* next_alignment_list_.push_back(...) for every emitted candidate */
      uint32_t *qid = (uint32_t *)xmalloc(c * sizeof(uint32_t));
      for (k = 0; k < c; ++k) qid[k] = i;
      append(&st->next_ids, &st->next_starts, &st->next_n, &st->next_cap, qid, tmp, c);
      free(qid);
    }
    free(tmp);
    alignment_count += c;
    if (alignment_count > max_list_length) { /* aligner.cpp:511-514 */
      st->next_query_id = i + 1;
      *query_ids = st->ids;
      *starts = st->starts;
      return st->n;
    }
    append(&st->ids, &st->starts, &st->n, &st->cap, st->next_ids, st->next_starts, st->next_n);
    st->next_n = 0;
  }
  st->next_query_id = n_queries; /* aligner.cpp:519 */
  *query_ids = st->ids;
  *starts = st->starts;
  return st->n;
}

/* ------------------------------------------------------- SW extension (a5) */

/* aligner.cpp:545-685  Aligner::CalculateScoreCpu */
void gmo_calculate_score(const uint8_t *db_sequence, uint32_t db_length, const uint8_t *query_sequences,
                         uint32_t query_sequence_length, uint64_t n, const uint32_t *query_ids,
                         const uint32_t *starts, const int *score_matrix, int open_gap,
                         int extend_gap, uint32_t extend, uint32_t log_region, uint32_t *scores,
                         uint32_t *ends) {
  uint32_t base_search_length = query_sequence_length + 2 * extend + 2 * (1u << log_region);
  uint32_t offset = extend;
  uint32_t query_search_length = query_sequence_length + 1;
  int *dp_column = (int *)xmalloc(query_search_length * sizeof(int));
  int *insertion_column = (int *)xmalloc(query_search_length * sizeof(int));
  uint32_t max_end = 0; /* aligner.cpp:560: NOT reset per candidate */
  uint64_t it;

  for (it = 0; it < n; ++it) {
    int db_offset = (int)(starts[it] - offset); /* aligner.cpp:576-579 */
    uint32_t db_search_length = base_search_length;
    long query_offset;
    int max_score = 0;
    uint32_t j, k;
    if (db_offset < 0) db_offset = 0;
    if ((uint32_t)db_offset + db_search_length > db_length)
      db_search_length = db_length - (uint32_t)db_offset;
    query_offset = (long)query_ids[it] * query_sequence_length - 1;

    for (j = 0; j < query_search_length; ++j) {
      dp_column[j] = 0;
      insertion_column[j] = 0;
    }
    for (j = 0; j < db_search_length; ++j) {
      uint8_t db_character = db_sequence[(uint32_t)db_offset + j];
      if (db_character != GMO_SEQUENCE_END) {
        int score_matrix_offset = db_character * GMO_ALPHABET_SIZE;
        int temp_score = 0;
        int deletion_score = 0;
        for (k = 1; k < query_search_length; ++k) {
          int local_score = 0;
          int score = temp_score + score_matrix[score_matrix_offset + query_sequences[query_offset + k]];
          if (score > 0) local_score = score;
          if (insertion_column[k] + extend_gap < dp_column[k] + open_gap)
            insertion_column[k] = dp_column[k] + open_gap;
          else
            insertion_column[k] += extend_gap;
          if (insertion_column[k] > local_score) local_score = insertion_column[k];
          if (deletion_score + extend_gap < dp_column[k - 1] + open_gap)
            deletion_score = dp_column[k - 1] + open_gap;
          else
            deletion_score += extend_gap;
          if (deletion_score > local_score) local_score = deletion_score;
          temp_score = dp_column[k];
          dp_column[k] = local_score;
          if (local_score >= max_score) { /* ">=": the LAST maximum wins */
            max_score = local_score;
            max_end = j;
          }
        }
      } else { /* aligner.cpp:664-669: columns reset, running max survives */
        for (k = 0; k < query_search_length; ++k) {
          dp_column[k] = 0;
          insertion_column[k] = 0;
        }
      }
    }
    scores[it] = (uint32_t)max_score;
    ends[it] = (uint32_t)db_offset + max_end;
  }
  free(dp_column);
  free(insertion_column);
}

/* --------------------------------------------------------- TraceBack (a7) */

/* aligner.cpp:771-949  Aligner::TraceBack */
void gmo_traceback(const uint8_t *db_sequence, const uint8_t *query, uint32_t query_sequence_length,
                   uint32_t db_end, const int *score_matrix, int open_gap, int extend_gap,
                   uint32_t extend, uint32_t log_region, uint32_t *db_start_out,
                   uint32_t *aln_len_out, uint32_t *aln_match_out, float *seq_id_out) {
  /* aligner.cpp:775: a product, not a sum */
  uint32_t base_search_length = query_sequence_length + 2 * extend * 2 * (1u << log_region);
  int query_search_length = (int)query_sequence_length + 1;
  int *dp_column = (int *)xmalloc((size_t)query_search_length * sizeof(int));
  int *insertion_column = (int *)xmalloc((size_t)query_search_length * sizeof(int));
  uint32_t *qry_aln_len = (uint32_t *)xmalloc((size_t)query_search_length * sizeof(uint32_t));
  uint32_t *aln_match = (uint32_t *)xmalloc((size_t)query_search_length * sizeof(uint32_t));
  uint32_t db_offset = db_end;
  uint32_t db_search_length = base_search_length;
  uint32_t max_start = 0, max_aln_match = 0, max_qry_aln_len = 0;
  int max_score = 0;
  uint32_t j;
  int k;
  if (db_offset < db_search_length) db_search_length = db_offset + 1;
  for (k = 0; k < query_search_length; ++k) {
    dp_column[k] = 0;
    insertion_column[k] = 0;
    aln_match[k] = 0;
    qry_aln_len[k] = 0;
  }
  for (j = 0; j < db_search_length; ++j) {
    uint8_t db_character = db_sequence[db_offset - j];
    int score_matrix_offset, temp_score = 0, deletion_score = 0;
    uint32_t temp_match = 0, new_match = 0, ins_match, del_match;
    uint32_t temp_aln_len = 0, new_aln_len = 0, ins_aln_len, del_aln_len;
    if (db_character == GMO_SEQUENCE_END) break; /* aligner.cpp:927-929 */
    score_matrix_offset = db_character * GMO_ALPHABET_SIZE;
    for (k = query_search_length - 2; 0 <= k; --k) {
      int local_score = 0;
      int score = temp_score + score_matrix[score_matrix_offset + query[k]];
      new_match = 0;
      new_aln_len = 0;
      if (score > 0) {
        local_score = score;
        new_match = (db_character == query[k]) ? temp_match + 1 : temp_match;
        new_aln_len = temp_aln_len + 1;
      }
      if (insertion_column[k] + extend_gap < dp_column[k] + open_gap)
        insertion_column[k] = dp_column[k] + open_gap;
      else
        insertion_column[k] += extend_gap;
      ins_match = aln_match[k];
      ins_aln_len = qry_aln_len[k] + 1;
      if (insertion_column[k] > local_score) {
        local_score = insertion_column[k];
        new_match = ins_match;
        new_aln_len = ins_aln_len;
      }
      if (deletion_score + extend_gap < dp_column[k + 1] + open_gap)
        deletion_score = dp_column[k + 1] + open_gap;
      else
        deletion_score += extend_gap;
      del_match = aln_match[k + 1];
      del_aln_len = qry_aln_len[k + 1] + 1;
      if (deletion_score > local_score) {
        local_score = deletion_score;
        new_match = del_match;
        new_aln_len = del_aln_len;
      }
      temp_score = dp_column[k];
      dp_column[k] = local_score;
      temp_match = aln_match[k];
      aln_match[k] = new_match;
      temp_aln_len = qry_aln_len[k];
      qry_aln_len[k] = new_aln_len;
      if (local_score > max_score) { /* ">": the FIRST maximum wins */
        max_score = local_score;
        max_start = j;
        max_aln_match = new_match;
        max_qry_aln_len = new_aln_len;
      }
    }
  }
  *db_start_out = db_offset - max_start;
  *seq_id_out = (float)max_aln_match / (float)(int)max_qry_aln_len; /* aligner.cpp:936-945: aln_len is int */
  *aln_len_out = max_qry_aln_len;
  *aln_match_out = max_aln_match;
  free(dp_column);
  free(insertion_column);
  free(qry_aln_len);
  free(aln_match);
}

/* ------------------------------------------------------------ Merge (a6) */

/* db.h:94-120  DB::GetID */
uint32_t gmo_db_get_id(const uint32_t *positions_, uint32_t number_sequences_,
                       uint32_t sequences_length_, uint32_t position) {
  uint32_t left, right, mid;
  if (positions_[number_sequences_ - 1] <= position && position < sequences_length_)
    return number_sequences_ - 1;
  left = 0;
  right = number_sequences_ - 2;
  while (left <= right) {
    mid = (left + right) / 2;
    if (positions_[mid] <= position && position < positions_[mid + 1]) {
      return mid;
    } else if (positions_[mid] < position) {
      left = mid + 1;
    } else {
      if (mid == 0) break; /* reference wraps to UINT_MAX and leaves the loop the same way */
      right = mid - 1;
    }
  }
  return GMO_UINT_MAX;
}

/* aligner.cpp:52-63  AlignmentComp */
static int hit_comp(const gmo_hit *a, const gmo_hit *b) { return a->score > b->score; }

static void hit_swap(gmo_hit *a, gmo_hit *b) {
  gmo_hit t = *a;
  *a = *b;
  *b = t;
}

/* libstdc++ 13.3 bits/stl_algo.h, restated.  _S_threshold = 16 (:1848). */
enum { GMO_S_THRESHOLD = 16 };

/* bits/stl_heap.h __push_heap */
static void push_heap_(gmo_hit *first, long hole, long top, gmo_hit value) {
  long parent = (hole - 1) / 2;
  while (hole > top && hit_comp(first + parent, &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

/* bits/stl_heap.h __adjust_heap */
static void adjust_heap_(gmo_hit *first, long hole, long len, gmo_hit value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (hit_comp(first + second, first + (second - 1))) second--;
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

/* __partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap */
static void heap_sort_(gmo_hit *first, gmo_hit *last) {
  long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    while (1) {
      gmo_hit value = first[parent];
      adjust_heap_(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    gmo_hit value;
    --last;
    value = *last; /* __pop_heap(first, last, last) */
    *last = *first;
    adjust_heap_(first, 0, last - first, value);
  }
}

static void move_median_to_first_(gmo_hit *result, gmo_hit *a, gmo_hit *b, gmo_hit *c) {
  if (hit_comp(a, b)) {
    if (hit_comp(b, c))
      hit_swap(result, b);
    else if (hit_comp(a, c))
      hit_swap(result, c);
    else
      hit_swap(result, a);
  } else if (hit_comp(a, c))
    hit_swap(result, a);
  else if (hit_comp(b, c))
    hit_swap(result, c);
  else
    hit_swap(result, b);
}

static gmo_hit *unguarded_partition_(gmo_hit *first, gmo_hit *last, gmo_hit *pivot) {
  while (1) {
    while (hit_comp(first, pivot)) ++first;
    --last;
    while (hit_comp(pivot, last)) --last;
    if (!(first < last)) return first;
    hit_swap(first, last);
    ++first;
  }
}

static void introsort_loop_(gmo_hit *first, gmo_hit *last, long depth_limit) {
  while (last - first > GMO_S_THRESHOLD) {
    gmo_hit *mid, *cut;
    if (depth_limit == 0) {
      heap_sort_(first, last);
      return;
    }
    --depth_limit;
    mid = first + (last - first) / 2;
    move_median_to_first_(first, first + 1, mid, last - 1);
    cut = unguarded_partition_(first + 1, last, first);
    introsort_loop_(cut, last, depth_limit);
    last = cut;
  }
}

static void unguarded_linear_insert_(gmo_hit *last) {
  gmo_hit val = *last;
  gmo_hit *next = last - 1;
  while (hit_comp(&val, next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

static void insertion_sort_(gmo_hit *first, gmo_hit *last) {
  gmo_hit *i;
  if (first == last) return;
  for (i = first + 1; i != last; ++i) {
    if (hit_comp(i, first)) {
      gmo_hit val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(gmo_hit));
      *first = val;
    } else {
      unguarded_linear_insert_(i);
    }
  }
}

void gmo_std_sort_hits(gmo_hit *first, size_t n) {
  gmo_hit *last = first + n;
  long lg = 0;
  size_t m;
  if (n == 0) return;
  for (m = n; m > 1; m >>= 1) ++lg; /* std::__lg */
  introsort_loop_(first, last, lg * 2);
  if (last - first > GMO_S_THRESHOLD) { /* __final_insertion_sort */
    gmo_hit *i;
    insertion_sort_(first, first + GMO_S_THRESHOLD);
    for (i = first + GMO_S_THRESHOLD; i != last; ++i) unguarded_linear_insert_(i);
  } else {
    insertion_sort_(first, last);
  }
}

typedef struct {
  gmo_hit *results;
  uint32_t *result_counts;
  uint32_t result_cap;
  uint32_t *overlap;
  const uint8_t *queries;
  uint32_t query_len;
  const uint8_t *db;
  uint32_t db_len;
  const uint32_t *seq_starts;
  uint32_t n_seqs;
  uint32_t db_chunk;
  const int *score_matrix;
  int open_gap, extend_gap;
  uint32_t extend, log_region, best;
} merge_ctx;

/* aligner.cpp:701-725 (and the identical tail :745-768) */
static void merge_flush(merge_ctx *c, gmo_hit *l, size_t n, uint32_t id) {
  size_t it;
  gmo_std_sort_hits(l, n);
  for (it = 0; it < n; ++it) {
    gmo_hit *h = &l[it];
    uint32_t db_id = h->db_id;
    if (db_id == GMO_UINT_MAX) {
      db_id = gmo_db_get_id(c->seq_starts, c->n_seqs, c->db_len, h->db_end);
      if (c->overlap[db_id] != id) {
        uint32_t db_position;
        c->overlap[db_id] = id;
        gmo_traceback(c->db, c->queries + (size_t)h->query_id * c->query_len, c->query_len,
                      h->db_end, c->score_matrix, c->open_gap, c->extend_gap, c->extend,
                      c->log_region, &h->db_start, &h->aln_len, &h->aln_match, &h->seq_id);
        db_position = c->seq_starts[db_id];
        h->db_id = db_id;
        h->db_chunk = c->db_chunk;
        h->db_start -= db_position;
        h->db_end -= db_position;
        if (c->result_counts[id] >= c->result_cap) abort();
        c->results[(size_t)id * c->result_cap + c->result_counts[id]++] = *h;
      }
    } else {
      if (c->result_counts[id] >= c->result_cap) abort();
      c->results[(size_t)id * c->result_cap + c->result_counts[id]++] = *h;
    }
    if (c->result_counts[id] >= c->best) break;
  }
}

/* aligner.cpp:687-769  Aligner::Merge */
void gmo_merge(gmo_hit *results, uint32_t *result_counts, uint32_t result_cap, uint64_t n,
               const uint32_t *query_ids, const uint32_t *starts, const uint32_t *scores,
               const uint32_t *ends, const uint8_t *queries, uint32_t n_queries,
               uint32_t query_len, const uint8_t *name_break, const uint8_t *db, uint32_t db_len,
               const uint32_t *seq_starts, uint32_t n_seqs, uint32_t db_chunk,
               const int *score_matrix, int open_gap, int extend_gap, uint32_t extend,
               uint32_t log_region, uint32_t best) {
  merge_ctx c;
  gmo_hit *l = NULL;
  size_t ln = 0, lcap = 0;
  uint64_t cand = 0;
  uint32_t i, s;
  c.results = results;
  c.result_counts = result_counts;
  c.result_cap = result_cap;
  c.overlap = (uint32_t *)xmalloc((size_t)n_seqs * sizeof(uint32_t));
  for (s = 0; s < n_seqs; ++s) c.overlap[s] = GMO_UINT_MAX;
  c.queries = queries;
  c.query_len = query_len;
  c.db = db;
  c.db_len = db_len;
  c.seq_starts = seq_starts;
  c.n_seqs = n_seqs;
  c.db_chunk = db_chunk;
  c.score_matrix = score_matrix;
  c.open_gap = open_gap;
  c.extend_gap = extend_gap;
  c.extend = extend;
  c.log_region = log_region;
  c.best = best;

  for (i = 0; i < n_queries; ++i) {
    uint32_t r;
    if (i > 0 && name_break[i]) { /* prev_query_name != query_name */
      merge_flush(&c, l, ln, i - 1);
      ln = 0;
    }
    for (; cand < n; ++cand) { /* aligner.cpp:732-737 */
      gmo_hit h;
      if (i != query_ids[cand]) break;
      h.query_id = query_ids[cand];
      h.db_id = GMO_UINT_MAX;
      h.db_chunk = GMO_UINT_MAX;
      h.score = scores[cand];
      h.db_start = starts[cand];
      h.db_end = ends[cand];
      h.aln_len = GMO_UINT_MAX;
      h.aln_match = GMO_UINT_MAX;
      h.seq_id = 0.0f;
      if (ln == lcap) {
        lcap = lcap ? lcap * 2 : 256;
        l = (gmo_hit *)xrealloc(l, lcap * sizeof(gmo_hit));
      }
      l[ln++] = h;
    }
    for (r = 0; r < result_counts[i]; ++r) { /* aligner.cpp:738-740 */
      if (ln == lcap) {
        lcap = lcap ? lcap * 2 : 256;
        l = (gmo_hit *)xrealloc(l, lcap * sizeof(gmo_hit));
      }
      l[ln++] = results[(size_t)i * result_cap + r];
    }
    result_counts[i] = 0; /* aligner.cpp:741 */
  }
  if (n_queries > 0) merge_flush(&c, l, ln, n_queries - 1);
  free(l);
  free(c.overlap);
}

/* ------------------------------------------------------ scoring + output */

static const char *kBlosum62Letters = "ARNDCQEGHILKMFPSTWYVBZX*";
static const signed char kBlosum62[24][24] = {
    {4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0, -4},
    {-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1, -4},
    {-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1, -4},
    {-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2, -4},
    {-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1, -4},
    {-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1, -4},
    {-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1, -4},
    {-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1, -4},
    {-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1, -4},
    {-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1, -4},
    {-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1, -4},
    {-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1, -4},
    {-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2, -4},
    {1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0, -4},
    {0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0, -4},
    {-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2, -4},
    {-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1, -4},
    {0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1, -4},
    {-2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1, -4},
    {-1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1, -4},
    {0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1, -4},
    {-4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, -4, 1}};

/* sequence.cpp:63-87 restricted to the letters of the matrix header */
static int protein_code(char ch) {
  static const char *letters = "ARNDCQEGHILKMFPSTWYVBJZX*"; /* codes 0..24 */
  const char *p = strchr(letters, ch);
  return p ? (int)(p - letters) : GMO_BASE_X;
}

/* score_matrix_reader.cpp:80-113 applied to the built-in text (:41-42): entries for
 * letters absent from the header (J=21, END=25, codes 26..31) stay 0. */
void gmo_blosum62(int *matrix) {
  int r, c;
  for (r = 0; r < GMO_ALPHABET_SIZE * GMO_ALPHABET_SIZE; ++r) matrix[r] = 0;
  for (r = 0; r < 24; ++r)
    for (c = 0; c < 24; ++c)
      matrix[protein_code(kBlosum62Letters[r]) * GMO_ALPHABET_SIZE +
             protein_code(kBlosum62Letters[c])] = kBlosum62[r][c];
}

/* aligner.cpp:956-963 */
uint32_t gmo_query_length(const uint8_t *query, uint32_t query_len) {
  uint32_t start = 0, end = query_len - 1, offset;
  for (offset = end; offset > start && query[offset] == GMO_BASE_X; --offset)
    ;
  return offset - start + 1;
}

/* aligner.cpp:964-976 with statistics.cpp:40-59.  The reference streams float/double
 * through a default ostream: that is printf's %g with precision 6. */
int gmo_format_row(char *buf, size_t buflen, const char *query_name, const char *db_name,
                   const gmo_hit *hit, uint32_t query_length, uint64_t db_length, float lambda,
                   float K) {
  uint64_t search_space = (uint64_t)query_length * db_length; /* statistics.cpp:57-59 */
  /* statistics.cpp:40-44.  statistics.cpp includes <math.h> under g++, so log(float)
   * resolves to the float overload: every operand and the result are float. */
  float bit_score = ((((float)(int)hit->score * lambda) - logf(K)) / (float)log(2.0));
  /* statistics.cpp:51-55: uint64*float is a FLOAT product; -1.0*int*float is double. */
  float space_k = (float)search_space * K;
  /* aligner.cpp:966: the double result is stored into a FLOAT before it is streamed. */
  float e_value = (float)(space_k * exp((double)(-1.0 * (int)hit->score * lambda)));
  return snprintf(buf, buflen, "%s\t%s\t%g\t%u\t%u\t%u\t%u\t%g\t%g\t\n", query_name, db_name,
                  (double)(hit->seq_id * 100), hit->aln_len, hit->aln_match, hit->db_start + 1,
                  hit->db_end + 1, (double)e_value, (double)bit_score);
}
