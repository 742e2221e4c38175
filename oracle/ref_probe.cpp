/*
 * TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * ref_probe: drives the UNMODIFIED reference Aligner stage by stage and dumps
 * the intermediate parity points that `ghostm aln` itself only exposes through
 * `#if 0` blocks (reference aligner.cpp:140-154):
 *
 *   (1) candidates  (query_id, db_start)        after Aligner::SearchNext      (aligner.cpp:347-381)
 *   (2) scores      (score, db_end)             after Aligner::CalculateScore  (aligner.cpp:523-543)
 *   (3) hit lists   after the last Aligner::Merge of a query chunk             (aligner.cpp:687-769)
 *
 * It is compiled by oracle/Makefile against the reference objects where they lie
 * under /root/reference (only main.o is replaced).  The stage methods are private
 * (aligner.h:74-91); the probe reaches them with the usual test-only
 * `#define private public` around the reference header.  The loop below follows
 * Aligner::Execute (aligner.cpp:98-204), CPU mode only.
 *
 * usage: [GMPROBE_MAX_LIST_LENGTH=n] ref_probe <dump.bin> <aln options...>   (options of `ghostm aln`)
 *
 * dump format (little endian u32 unless noted):
 *   "GMPROBE1"
 *   records:  tag=1  qchunk dbchunk cchunk n  then n x {query_id, start, score, end}
 *             tag=2  qchunk nq  then per query: nhits, nhits x {db_chunk_local_id, score,
 *                    db_start, db_end, aln_len, aln_match, seq_id(float bits), name_len, name bytes}
 *             tag=0  end
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <limits.h>
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <list>
#include <algorithm>
#include <map>
#include <stdexcept>

#define private public
#include "aligner.h"
#undef private
#include "query.h"
#include "query_reader.h"
#include "db.h"
#include "db_reader.h"
#include "score_matrix.h"
#include "common.h"

static void put32(FILE *f, uint32_t v) { fwrite(&v, 4, 1, f); }

int main(int argc, char *argv[]) {
  if (argc < 3) {
    fprintf(stderr, "usage: ref_probe <dump.bin> <aln options>\n");
    return 2;
  }
  FILE *dump = fopen(argv[1], "wb");
  if (!dump) { perror(argv[1]); return 2; }
  fwrite("GMPROBE1", 1, 8, dump);

  Aligner aligner;
  AlignerOption option;
  try {
    aligner.SetOption(argc - 1, argv + 1, option);
  } catch (std::exception &e) {
    fprintf(stderr, "ref_probe: %s\n", e.what());
    return 2;
  }
  // The CLI can only express the candidate budget in units of 2^20 (aligner.cpp:291-293);
  // the probe lets tests set AlignerOption::max_list_length directly so that the
  // candidate-chunk rule (aligner.cpp:511-516) is exercised at small sizes.
  if (getenv("GMPROBE_MAX_LIST_LENGTH"))
    option.max_list_length = (uint32_t)strtoul(getenv("GMPROBE_MAX_LIST_LENGTH"), NULL, 10);
  std::ofstream out;
  if (!option.output_file_name.empty()) out.open(option.output_file_name.c_str());

  QueryReader query_reader(option.query_file_prefix);
  Query *query = (option.start_query_file_id == UINT_MAX)
                     ? query_reader.Read()
                     : query_reader.Read(option.start_query_file_id);
  uint32_t qchunk = (option.start_query_file_id == UINT_MAX) ? 0 : option.start_query_file_id;
  while (query != NULL) {
    std::vector<std::vector<Alignment> > result_list(query->GetNumberSequences());
    DBReader db_reader(option.db_file_prefix);
    uint32_t dbchunk = 0;
    for (DB *db = db_reader.Read(); db != NULL; db = db_reader.Read(), ++dbchunk) {
      aligner.next_query_id_ = 0;
      aligner.next_alignment_list_.clear();
      std::vector<Alignment> alignment_list;
      for (uint32_t cchunk = 0;; ++cchunk) {
        aligner.SearchNext(query, db, option, alignment_list);
        if (alignment_list.empty()) break;
        std::vector<uint32_t> starts(alignment_list.size());
        for (size_t i = 0; i < alignment_list.size(); ++i) starts[i] = alignment_list[i].GetDbStart();
        aligner.CalculateScore(query, db, alignment_list, option);
        put32(dump, 1); put32(dump, qchunk); put32(dump, dbchunk); put32(dump, cchunk);
        put32(dump, (uint32_t)alignment_list.size());
        for (size_t i = 0; i < alignment_list.size(); ++i) {
          put32(dump, alignment_list[i].GetQueryId());
          put32(dump, starts[i]);
          put32(dump, alignment_list[i].GetScore());
          put32(dump, alignment_list[i].GetDbEnd());
        }
        aligner.Merge(result_list, alignment_list, query, db, option);
      }
      delete db;
    }
    put32(dump, 2); put32(dump, qchunk); put32(dump, query->GetNumberSequences());
    for (uint32_t i = 0; i < query->GetNumberSequences(); ++i) {
      put32(dump, (uint32_t)result_list[i].size());
      for (size_t h = 0; h < result_list[i].size(); ++h) {
        Alignment &a = result_list[i][h];
        float sid = a.GetSeqId();
        uint32_t sid_bits;
        memcpy(&sid_bits, &sid, 4);
        std::string name = a.GetDbName();
        put32(dump, a.GetDbId()); put32(dump, a.GetScore());
        put32(dump, a.GetDbStart()); put32(dump, a.GetDbEnd());
        put32(dump, a.GetAlnLen()); put32(dump, a.GetAlnMatch());
        put32(dump, sid_bits); put32(dump, (uint32_t)name.size());
        fwrite(name.data(), 1, name.size(), dump);
      }
    }
    if (out.is_open()) {
      DBReader sum_reader(option.db_file_prefix);
      switch (option.output_style) {
        case 1: aligner.WriteOutputV1(out, result_list, query); break;
        case 2: aligner.WriteOutputV2(out, result_list, query); break;
        default:
          aligner.WriteOutput(out, result_list, query, sum_reader.GetSumDbLength(),
                              option.statistics_parameters);
      }
    }
    db_reader.Close();
    delete query;
    query = NULL;
    ++qchunk;
    if (query_reader.GetNextId() <= option.end_query_file_id) query = query_reader.Read();
  }
  put32(dump, 0);
  fclose(dump);
  return 0;
}
