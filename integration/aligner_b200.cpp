// The "fast" integration of ghostm_b200 behind the reference's own host.
//
// This translation unit is the ONLY thing a maintainer of GHOSTM adds: it defines
// Aligner::Execute (reference aligner.cpp:65-223) on the extended gm_* C ABI of
// include/ghostm_b200.h.  Everything else of the reference is linked unchanged - main, command,
// DBReader / QueryReader / DB / Query / Index, ScoreMatrixReader, Statistics, and from aligner.cpp
// itself SetOption and the three WriteOutput* methods (its own Execute is kept under the name
// ExecuteReference and still serves `aln` without -D).  See oracle/Makefile, target `fast`.
//
// Differences to the reference's GPU branch (aligner.cpp:76-93,108-124,356-375,529-543):
//   * db chunks are uploaded once and stay resident in HBM while they fit; a chunk that does not
//     fit is streamed (upload, align, release) for every query chunk like the reference does
//     (SetDbGpu per iteration, aligner.cpp:115-124);
//   * SearchNext + CalculateScore + Merge + TraceBack of a db chunk are ONE call, gm_align_chunk;
//     only the <= best surviving records per query come back (gm_results_download);
//   * -v prints the per-phase times of aligner.cpp:132-159 from the CUDA-event counters.
#include <climits>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include <unistd.h>

#include "aligner.h"
#include "common.h"
#include "db.h"
#include "db_reader.h"
#include "index.h"
#include "query.h"
#include "query_reader.h"
#include "score_matrix.h"

#include "ghostm_b200.h"

using namespace std;

namespace {

void Check(int status, const char *what) {
  if (status != 0) throw runtime_error(string("error:") + what + ": " + gm_last_error());
}

struct ChunkNames {           // Alignment::db_name_ is filled from here (aligner.cpp:757)
  vector<string> names;
  bool known;
  ChunkNames() : known(false) {}
};

}  // namespace

int Aligner::Execute(int argc, char *argv[]) {
  AlignerOption option;
  SetOption(argc, argv, option);
  if (option.device == CPU) {          // no -D: the reference's own CPU path, untouched
    if (option.score_matrix != NULL) delete option.score_matrix;
    optind = 1;
    return ExecuteReference(argc, argv);
  }
  ofstream out(option.output_file_name.c_str());
  if (option.verbose) {
    cout << "#     G H O S T M " << endl;
    cout << "# * GPU-base HOmology Search Tool for Metagenomics *" << endl << endl;
  }

  gm_context *ctx = NULL;
  Check(gm_create(option.device, &ctx), "gm_create");
  uint32_t seed = 0;
  {
    DBReader db_reader(option.db_file_prefix);
    seed = db_reader.GetSeed();
    db_reader.Close();
  }
  gm_options go;
  memset(&go, 0, sizeof(go));
  go.seed = seed;
  go.shift = option.shift_size;
  go.log_region = option.log_region_size;
  go.threshold = option.threshold;
  go.extend = option.extend;
  go.best = option.best;
  go.max_list_length = option.max_list_length;
  go.open_gap = option.open_gap;
  go.extend_gap = option.extend_gap;
  memcpy(go.score_matrix, option.score_matrix->GetMatrix(), sizeof(go.score_matrix));
  Check(gm_set_options(ctx, &go), "gm_set_options");
  if (option.verbose) {
    cout << "  Maximun size of the candidates (-l option): " << option.max_list_length << " ("
         << (option.max_list_length >> 20) << "MB)" << endl;
    cout << "  Init GPU device [" << option.device << "] ... ok. " << gm_version() << endl << endl;
  }

  vector<ChunkNames> chunk_names;
  vector<char> resident;               // db chunk c stays in HBM
  const uint32_t cap = option.best > 1 ? option.best : 1;
  clock_t start;
  QueryReader query_reader(option.query_file_prefix);
  Query *query = NULL;
  if (option.start_query_file_id == UINT_MAX) {
    query = query_reader.Read();
  } else {
    query = query_reader.Read(option.start_query_file_id);
  }

  if (query != NULL) {
    while (query != NULL) {
      const uint32_t n = query->GetNumberSequences();
      vector<uint8_t> name_break(n, 0);          // same-name runs, aligner.cpp:697-700
      for (uint32_t i = 1; i < n; ++i) name_break[i] = query->GetName(i) != query->GetName(i - 1);
      uint64_t capacity = (uint64_t)n * 2048 + (1u << 22);
      if (capacity > 0xFFFFFFFFull) capacity = 0xFFFFFFFFull;
      Check(gm_set_candidate_capacity(ctx, capacity), "gm_set_candidate_capacity");
      Check(gm_query_upload(ctx, query->GetSequences(), n, query->GetSequenceLength(), &name_break[0]),
            "gm_query_upload");

      vector<vector<Alignment> > result_list(n);
      DBReader db_reader(option.db_file_prefix);
      DB *db = db_reader.Read();
      if (db != NULL) {
        uint32_t c = 0;
        while (db != NULL) {
          if (chunk_names.size() <= c) {
            chunk_names.resize(c + 1);
            resident.resize(c + 1, 0);
          }
          if (!chunk_names[c].known) {
            string *names = db->GetAllNames();
            chunk_names[c].names.assign(names, names + db->GetNumberSequences());
            chunk_names[c].known = true;
          }
          if (!resident[c]) {
            Index *db_index = db->GetIndex();
            uint64_t free_bytes = 0, total_bytes = 0;
            Check(gm_device_memory(ctx, &free_bytes, &total_bytes), "gm_device_memory");
            // residues + CSR index + .pos table + the per-key tile boundaries the seed search adds
            const uint64_t need = (uint64_t)db->GetSequencesLength() + 4ull * db_index->GetKeysCountLength() * 65
                                + 4ull * db_index->GetPositionsLength() + 4ull * db->GetNumberSequences();
            Check(gm_db_upload(ctx, c, db->GetSequences(), db->GetSequencesLength(), db_index->GetKeysCount(),
                               db_index->GetKeysCountLength(), db_index->GetAllPositions(),
                               db_index->GetPositionsLength(), db->GetAllPositions(), db->GetNumberSequences()),
                  "gm_db_upload");
            // keep it while a comfortable part of the device stays free for candidates and hit lists
            resident[c] = free_bytes > need + 16ull * capacity + (4ull << 30);
          }
          gm_stats st;
          memset(&st, 0, sizeof(st));
          if (option.verbose) cout << "|Search, score, merge db chunk " << c << " ... ";
          start = clock();
          int rc = gm_align_chunk(ctx, c, &st);
          while (rc == GM_ERR_CAPACITY && capacity < 0xFFFFFFFFull) {   // skewed data: grow and redo the chunk
            capacity = capacity * 2 > 0xFFFFFFFFull ? 0xFFFFFFFFull : capacity * 2;
            Check(gm_set_candidate_capacity(ctx, capacity), "gm_set_candidate_capacity");
            memset(&st, 0, sizeof(st));
            rc = gm_align_chunk(ctx, c, &st);
          }
          Check(rc, "gm_align_chunk");
          if (option.verbose) {
            cout << (float)(clock() - start) / (float)CLOCKS_PER_SEC << " sec." << endl;
            cout << "|  Search alignment candidates ... " << st.ms_search * 1e-3f << " sec. (" << st.candidates
                 << " candidates in " << st.candidate_chunks << " chunk(s))" << endl;
            cout << "|  Calculate scores ... " << st.ms_score * 1e-3f << " sec." << endl;
            cout << "|  Merge results ... " << (st.ms_merge + st.ms_traceback) * 1e-3f << "sec" << endl;
          }
          if (!resident[c]) Check(gm_db_release(ctx, c), "gm_db_release");   // traces its pending hits first
          delete db;
          db = db_reader.Read();
          ++c;
        }

        if (option.verbose) cout << "|Write results ... ";
        start = clock();
        vector<gm_hit> hits((size_t)n * cap);
        vector<uint32_t> counts(n, 0);
        Check(gm_results_download(ctx, &hits[0], &counts[0]), "gm_results_download");
        for (uint32_t i = 0; i < n; ++i) {
          for (uint32_t k = 0; k < counts[i]; ++k) {
            const gm_hit &h = hits[(size_t)i * cap + k];
            Alignment a;
            a.SetQueryId(h.query_id);
            a.SetDbId(h.db_id);
            a.SetDbName(chunk_names[h.db_chunk].names[h.db_id]);
            a.SetScore(h.score);
            a.SetDbStart(h.db_start);
            a.SetDbEnd(h.db_end);
            a.SetSeqId(h.seq_id);
            a.SetAlnLen(h.aln_len);
            a.SetAlnMatch(h.aln_match);
            result_list[i].push_back(a);
          }
        }
        DBReader db_reader(option.db_file_prefix);
        switch (option.output_style) {
        case 1:
          WriteOutputV1(out, result_list, query);
          break;
        case 2:
          WriteOutputV2(out, result_list, query);
          break;
        default:
          WriteOutput(out, result_list, query, db_reader.GetSumDbLength(), option.statistics_parameters);
          break;
        }
        if (option.verbose) cout << (float)(clock() - start) / (float)CLOCKS_PER_SEC << " sec." << endl;
      } else {
        cerr << "[Aligner] error: don't find db file." << endl;
        db_reader.Close();
        delete query;
        break;
      }
      db_reader.Close();
      delete query;
      query = NULL;
      if (query_reader.GetNextId() <= option.end_query_file_id) {
        query = query_reader.Read();
      }
    }
  } else {
    cerr << "[Aligner] error: don't find query file." << endl;
  }

  gm_destroy(ctx);
  query_reader.Close();
  if (option.score_matrix != NULL) {
    delete option.score_matrix;
  }
  out.close();
  if (option.verbose) cout << "Complete." << endl;
  return SUCCESS;
}
