// ghostm_b200 - host driver: `ghostm_b200 aln ...`, a drop-in for the reference's `ghostm aln`.
//
// Same command line (reference aligner.cpp:225-345), same db / query file formats (db_reader.cpp,
// db.cpp, query_reader.cpp, query.cpp), same tab-separated hit list (aligner.cpp:951-1012).  The
// search itself - seed lookup, candidate chunking, SW extension, Merge, TraceBack - runs on the
// GPU through the C ABI of include/ghostm_b200.h; nothing here computes an alignment and there is
// no CPU fallback.  `-D` takes one device id like the reference, or a comma separated list: the
// db chunks (index) are then spread round-robin over the devices for seed search + SW extension,
// the scored candidates go device to device by query slice (gm_candidates_transfer) and every
// device runs Merge + TraceBack for its slice of the queries over all chunks in ascending order
// (DESIGN.md section 8), which keeps the result identical to a single-device (and to the
// reference's) run.
#include <getopt.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/ghostm_b200.h"

namespace {

constexpr int kAlphabet = 32;      // common.h:31
constexpr uint8_t kBaseX = 23;     // common.h:35

// ---- residue codes (sequence.cpp:63-87): A0 R1 N2 D3 C4 Q5 E6 G7 H8 I9 L10 K11 M12 F13 P14 S15
// T16 W17 Y18 V19 B20 J21 Z22 X23 *24; everything else is X.
int residue_code(char ch) {
  static const char *letters = "ARNDCQEGHILKMFPSTWYVBJZX*";
  if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 'a' + 'A');
  const char *p = ch ? strchr(letters, ch) : nullptr;
  return p ? (int)(p - letters) : kBaseX;
}

// ---- score matrix (score_matrix_reader.cpp:44-113) -------------------------------------------
const char *kBlosum62 =
    "   A  R  N  D  C  Q  E  G  H  I  L  K  M  F  P  S  T  W  Y  V  B  Z  X  *\n"
    "A  4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4\n"
    "R -1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4\n"
    "N -2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4\n"
    "D -2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4\n"
    "C  0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4\n"
    "Q -1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4\n"
    "E -1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4\n"
    "G  0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4\n"
    "H -2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4\n"
    "I -1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4\n"
    "L -1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4\n"
    "K -1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4\n"
    "M -1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4\n"
    "F -2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4\n"
    "P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4\n"
    "S  1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4\n"
    "T  0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4\n"
    "W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4\n"
    "Y -2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4\n"
    "V  0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4\n"
    "B -2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4\n"
    "Z -1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4\n"
    "X  0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4\n"
    "* -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1\n";

struct ScoreMatrix {
  std::string name;
  int32_t m[kAlphabet * kAlphabet];
};

// NCBI text format; a row/column is taken from the FIRST character of its token, at most 25
// rows and 25 tokens per row are read, '#' lines are comments (score_matrix_reader.cpp:80-113).
void parse_matrix(std::istream &in, ScoreMatrix *out) {
  memset(out->m, 0, sizeof(out->m));
  char cols[kAlphabet] = {0};
  std::string line;
  int line_no = 0;
  while (std::getline(in, line)) {
    if (line.empty() || line[0] == '#' || line_no >= 25) continue;
    std::istringstream ss(line);
    std::string tok;
    char row = 0;
    for (int i = 0; i < 25 && (ss >> tok); ++i) {
      if (line_no == 0) cols[i] = tok[0];
      else if (i == 0) row = tok[0];
      else out->m[residue_code(row) * kAlphabet + residue_code(cols[i - 1])] = atoi(tok.c_str());
    }
    ++line_no;
  }
}

ScoreMatrix load_matrix(const std::string &file) {
  ScoreMatrix sm;
  std::ifstream in(file.c_str());
  if (in) {  // score_matrix_reader.cpp:44-60: the name is the file's basename
    const size_t slash = file.find_last_of('/');
    sm.name = slash == std::string::npos ? file : file.substr(slash + 1);
    parse_matrix(in, &sm);
  } else {   // anything that cannot be opened means the built-in BLOSUM62
    std::istringstream def(kBlosum62);
    sm.name = "BLOSUM62";
    parse_matrix(def, &sm);
  }
  return sm;
}

// ---- statistics (statistics.cpp:40-59, 134-146) -----------------------------------------------
struct Karlin { float lambda, K, H; };

Karlin gapped_karlin(const ScoreMatrix &sm, int open_gap, int extend_gap) {
  if (sm.name == "BLOSUM62" && open_gap == -11 && extend_gap == -1) return Karlin{0.267f, 0.041f, 0.14f};
  if (sm.name == "PAM30" && open_gap == -9 && extend_gap == -1) return Karlin{0.294f, 0.11f, 0.61f};
  throw std::invalid_argument("error: not support score option");
}

float bit_score(int score, const Karlin &k) {  // every operand is float (statistics.cpp:40-44)
  return ((static_cast<float>(score) * k.lambda) - logf(k.K)) / static_cast<float>(log(2.0));
}

double e_value(int score, uint64_t search_space, const Karlin &k) {  // statistics.cpp:51-55
  return (static_cast<float>(search_space) * k.K) * exp(static_cast<double>(-1.0 * score * k.lambda));
}

// ---- formatted files --------------------------------------------------------------------------
template <typename T>
bool read_vec(const std::string &path, size_t count, std::vector<T> *out, size_t skip_bytes = 0) {
  std::ifstream in(path.c_str(), std::ios::binary);
  if (!in) return false;
  in.seekg((std::streamoff)skip_bytes);
  out->resize(count);
  in.read(reinterpret_cast<char *>(out->data()), (std::streamsize)(count * sizeof(T)));
  return true;
}

std::vector<std::string> read_names(const std::string &path, uint32_t n) {  // db.cpp:36-61
  std::vector<std::string> names(n);
  std::ifstream in(path.c_str());
  std::string line;
  uint32_t i = 0;
  for (; i < n && in && !in.eof(); ++i) {
    std::getline(in, line);
    names[i] = line;
  }
  if (i < n) std::cerr << "warning : couldn't read all sequence names" << std::endl;
  return names;
}

struct DbInfo {          // <db>.inf, db_creator.cpp:243-264 / db_reader.cpp:36-50
  int32_t division = 0;
  uint32_t seed = 0, max_chunk_len = 0;
  uint64_t sum_length = 0;
};

bool read_db_info(const std::string &prefix, DbInfo *info) {
  std::ifstream in((prefix + ".inf").c_str(), std::ios::binary);
  if (!in) return false;
  in.read(reinterpret_cast<char *>(&info->division), 4);
  in.read(reinterpret_cast<char *>(&info->seed), 4);
  in.read(reinterpret_cast<char *>(&info->max_chunk_len), 4);
  in.read(reinterpret_cast<char *>(&info->sum_length), 8);
  return true;
}

struct DbChunk {         // <db>_<i>.{inf,seq,pos,nam,ind}
  uint32_t n_seqs = 0, seq_len = 0, seed = 0;
  std::vector<uint8_t> seq;
  std::vector<uint32_t> seq_starts, keys_count, positions;
  std::vector<std::string> names;
};

bool read_db_chunk(const std::string &prefix, int i, DbChunk *c, bool with_arrays) {
  std::ostringstream p;
  p << prefix << "_" << i;
  std::vector<uint32_t> inf;
  if (!read_vec(p.str() + ".inf", 2, &inf)) return false;
  c->n_seqs = inf[0];
  c->seq_len = inf[1];
  c->names = read_names(p.str() + ".nam", c->n_seqs);
  if (!with_arrays) return true;
  std::vector<uint32_t> hdr;
  if (!read_vec(p.str() + ".seq", c->seq_len, &c->seq) || !read_vec(p.str() + ".pos", c->n_seqs, &c->seq_starts) ||
      !read_vec(p.str() + ".ind", 3, &hdr))
    return false;
  c->seed = hdr[0];
  return read_vec(p.str() + ".ind", hdr[1], &c->keys_count, 12) &&
         read_vec(p.str() + ".ind", hdr[2], &c->positions, 12 + (size_t)hdr[1] * 4);
}

struct QueryChunk {      // <q>_<i>.{inf,seq,nam}, query_creator.cpp:346-419
  uint32_t n = 0, length = 0;
  std::vector<uint8_t> seqs;
  std::vector<std::string> names;
};

bool read_query_chunk(const std::string &prefix, uint32_t i, QueryChunk *q) {
  std::ostringstream p;
  p << prefix << "_" << i;
  std::vector<uint32_t> inf;
  if (!read_vec(p.str() + ".inf", 2, &inf)) return false;
  q->n = inf[0];
  q->length = inf[1];
  if (!read_vec(p.str() + ".seq", (size_t)q->n * q->length, &q->seqs)) return false;
  q->names = read_names(p.str() + ".nam", q->n);
  return true;
}

// ---- options (aligner.cpp:225-345) ------------------------------------------------------------
struct Options {
  std::string output, queries, db, matrix_file = "BLOSUM62";
  uint32_t log_region = 4, shift = 2, threshold = 2, max_list_length = 1u << 27, extend = 2, best = 10;
  int open_gap = -11, extend_gap = -1;
  uint32_t start_chunk = UINT32_MAX, end_chunk = UINT32_MAX;
  int style = 0;
  bool verbose = false;
  std::vector<int> devices;
};

Options parse_options(int argc, char **argv) {
  Options o;
  int c;
  while ((c = getopt(argc, argv, "b:d:D:e:E:G:i:l:M:o:r:s:t:S:L:y:v")) >= 0) {
    switch (c) {
      case 'b': o.best = (uint32_t)atoi(optarg); break;
      case 'd': o.db = optarg; break;
      case 'D': {
        std::istringstream ss(optarg);
        std::string tok;
        while (std::getline(ss, tok, ',')) o.devices.push_back(atoi(tok.c_str()));
        break;
      }
      case 'e': o.extend = (uint32_t)atoi(optarg); break;
      case 'E': o.extend_gap = -1 * atoi(optarg); break;
      case 'G': o.open_gap = -1 * atoi(optarg); break;
      case 'i': o.queries = optarg; break;
      case 'S': o.start_chunk = (uint32_t)atoi(optarg); break;
      case 'L': o.end_chunk = (uint32_t)atoi(optarg); break;
      case 'l': o.max_list_length = (uint32_t)(atoi(optarg) * (1 << 20)); break;
      case 'M': o.matrix_file = optarg; break;
      case 'o': o.output = optarg; break;
      case 'r': {
        int lr = (int)log2((double)atoi(optarg));
        o.log_region = lr < 1 ? 1u : (uint32_t)lr;
        break;
      }
      case 's': o.shift = (uint32_t)atoi(optarg); break;
      case 't': o.threshold = (uint32_t)atoi(optarg); break;
      case 'y': o.style = atoi(optarg); break;
      case 'v': o.verbose = true; break;
      default: throw std::invalid_argument("");
    }
  }
  if (o.devices.empty()) o.devices.push_back(0);  // this build has no CPU path: default to GPU 0
  return o;
}

struct DeviceError : std::runtime_error {   // a gm_* call failed: non-zero exit status, no partial output file
  explicit DeviceError(const std::string &m) : std::runtime_error(m) {}
};

void check(int rc, const char *what) {
  if (rc != 0) throw DeviceError(std::string(what) + ": " + gm_last_error());
}

// gm_align_chunk / gm_search stop with GM_ERR_CAPACITY before anything was merged when skewed data
// produces more candidates than the buffer holds: double the buffer and redo the call.
template <typename F>
void with_capacity_retry(gm_context *ctx, uint64_t *capacity, const char *what, F call) {
  int rc = call();
  while (rc == GM_ERR_CAPACITY && *capacity < 0xFFFFFFFFull) {
    *capacity = std::min<uint64_t>(*capacity * 2, 0xFFFFFFFFull);
    check(gm_set_candidate_capacity(ctx, *capacity), "gm_set_candidate_capacity");
    rc = call();
  }
  check(rc, what);
}

// ---- several devices: chunk-parallel front, query-sliced back ------------
struct Segment { uint32_t first, end; };   // one candidate chunk = one Merge call (aligner.cpp:131-171)

// Slice boundaries at same-name run starts, about n / world queries each.
std::vector<uint32_t> slice_bounds(const std::vector<uint8_t> &name_break, uint32_t n, size_t world) {
  std::vector<uint32_t> bounds(world + 1, 0);
  for (size_t r = 1; r < world; ++r) {
    uint32_t b = (uint32_t)(((uint64_t)r * n + world / 2) / world);
    while (b < n && b > 0 && !name_break[b]) ++b;   // next run start at or after the target
    bounds[r] = std::max(b, bounds[r - 1]);
  }
  bounds[world] = n;
  return bounds;
}

// Seed search + SW extension of one db chunk for all queries; the candidate-chunk segments.
std::vector<Segment> front_chunk(gm_context *ctx, uint32_t chunk, uint32_t n_queries, uint32_t max_list_length,
                                 uint64_t *capacity, gm_stats *stats) {
  std::vector<uint32_t> counts(n_queries);
  uint64_t total = 0;
  with_capacity_retry(ctx, capacity, "gm_search",
                      [&] { return gm_search(ctx, chunk, counts.data(), &total, stats); });
  std::vector<Segment> segs;
  uint32_t first = 0;
  while (true) {
    uint64_t n = 0;
    int last = 0;
    const uint32_t end = gm_chunk_rule(counts.data(), n_queries, first, max_list_length, &n, &last);
    if (n == 0) break;                                   // aligner.cpp:136-139
    check(gm_score(ctx, first, end, nullptr, nullptr, stats), "gm_score");
    segs.push_back(Segment{first, end});
    if (last) break;
    first = end;
  }
  return segs;
}

struct Shard {            // one device of a multi-device run
  gm_context *front = nullptr, *back = nullptr;
  uint64_t capacity = 0;
  std::mutex front_mu;    // gm_candidates_transfer is serialised per sending context
  std::string error;
};

template <typename F>
void on_every_device(std::vector<Shard> &shards, F body) {
  std::vector<std::thread> workers;
  for (size_t d = 0; d < shards.size(); ++d)
    workers.emplace_back([&shards, &body, d] {
      try { body(d); } catch (std::exception &e) { shards[d].error = e.what(); }
    });
  for (auto &w : workers) w.join();
  for (auto &sh : shards)
    if (!sh.error.empty()) throw DeviceError(sh.error);   // a worker died on a gm_* call
}

// ---- output (aligner.cpp:951-1012) -------------------------------------------------------------
void write_hits(std::ostream &out, const Options &o, const QueryChunk &q, const std::vector<gm_hit> &hits,
                const std::vector<uint32_t> &counts, uint32_t cap, const std::vector<DbChunk> &names,
                uint64_t db_length, const Karlin &karlin) {
  for (uint32_t i = 0; i < q.n; ++i) {
    const std::string &qname = q.names[i];
    uint64_t search_space = 0;
    if (o.style == 0) {  // query length without the trailing X padding (aligner.cpp:956-963)
      const uint32_t start = i * q.length, end = (i + 1) * q.length - 1;
      uint32_t offset = end;
      for (; offset > start && q.seqs[offset] == kBaseX; --offset) {}
      search_space = (uint64_t)(offset - start + 1) * db_length;
    }
    for (uint32_t k = 0; k < counts[i]; ++k) {
      const gm_hit &h = hits[(size_t)i * cap + k];
      const std::string &dname = names[h.db_chunk].names[h.db_id];
      if (o.style == 1) {
        out << qname << "\t" << dname << "\t" << h.score << "\t" << h.db_start + 1 << "\t" << h.db_end + 1 << std::endl;
      } else if (o.style == 2) {
        out << qname << "\t" << dname << "\t" << h.score << "\t" << h.db_start + 1 << "\t" << h.db_end + 1 << "\t"
            << h.seq_id << "\t" << h.aln_len << "\t" << h.aln_match << std::endl;
      } else {
        const float bits = bit_score((int)h.score, karlin);
        const float ev = (float)e_value((int)h.score, search_space, karlin);  // stored as float, aligner.cpp:966
        out << qname << "\t" << dname << "\t" << h.seq_id * 100 << "\t" << h.aln_len << "\t" << h.aln_match << "\t"
            << h.db_start + 1 << "\t" << h.db_end + 1 << "\t" << ev << "\t" << bits << "\t" << std::endl;
      }
    }
  }
}

std::string g_output_path;   // removed when the run dies on a device error

int run_aln(int argc, char **argv) {
  Options o = parse_options(argc, argv);
  g_output_path = o.output;
  const ScoreMatrix sm = load_matrix(o.matrix_file);
  Karlin karlin = {0, 0, 0};
  if (o.style == 0) karlin = gapped_karlin(sm, o.open_gap, o.extend_gap);
  std::ofstream out(o.output.c_str());
  if (o.verbose) {
    std::cout << "#     G H O S T M  (ghostm_b200, " << gm_version() << ")" << std::endl;
    std::cout << "# * GPU-base HOmology Search Tool for Metagenomics *" << std::endl << std::endl;
  }
  DbInfo info;
  if (!read_db_info(o.db, &info) || info.division <= 0) {
    std::cerr << "[Aligner] error: don't find db file." << std::endl;
    return 0;
  }
  const size_t n_dev = o.devices.size();
  gm_options go;
  memset(&go, 0, sizeof(go));
  go.seed = info.seed;
  go.shift = o.shift;
  go.log_region = o.log_region;
  go.threshold = o.threshold;
  go.extend = o.extend;
  go.best = o.best;
  go.max_list_length = o.max_list_length;
  go.open_gap = o.open_gap;
  go.extend_gap = o.extend_gap;
  memcpy(go.score_matrix, sm.m, sizeof(go.score_matrix));
  // one device: a single context does everything.  Several devices: per device a front context
  // (index of the owned chunks, all queries) and a back context (residues + .pos of every chunk,
  // the device's slice of the queries).
  std::vector<Shard> shards(n_dev);
  for (size_t d = 0; d < n_dev; ++d) {
    check(gm_create(o.devices[d], &shards[d].front), "gm_create");
    check(gm_set_options(shards[d].front, &go), "gm_set_options");
    if (n_dev > 1) {
      check(gm_create(o.devices[d], &shards[d].back), "gm_create");
      check(gm_set_options(shards[d].back, &go), "gm_set_options");
    }
  }
  // db chunks: resident in HBM for the whole run while they fit (the reference re-reads and
  // re-uploads every chunk for every query chunk, aligner.cpp:115-124).  On one device a chunk that
  // does not fit is STREAMED instead - read, upload, align, release, per query chunk - so a db larger
  // than device memory still runs; with several devices everything must fit (checked, loud error).
  std::vector<DbChunk> names(info.division);
  std::vector<char> resident(info.division, 0);
  const uint32_t n_chunks = (uint32_t)info.division;
  auto upload = [&](gm_context *ctx, uint32_t c, const DbChunk &full) {
    check(gm_db_upload(ctx, c, full.seq.data(), full.seq_len, full.keys_count.data(),
                       (uint32_t)full.keys_count.size(), full.positions.data(), (uint32_t)full.positions.size(),
                       full.seq_starts.data(), full.n_seqs), "gm_db_upload");
  };
  for (int c = 0; c < info.division; ++c) {
    DbChunk full;
    if (!read_db_chunk(o.db, c, &full, true)) {
      std::cerr << "[Aligner] error: don't find db file." << std::endl;
      return 0;
    }
    const size_t d = (size_t)c % n_dev;
    uint64_t free_bytes = 0, total_bytes = 0;
    check(gm_device_memory(shards[d].front, &free_bytes, &total_bytes), "gm_device_memory");
    // residues + CSR index + .pos table + the per-key tile boundaries the seed search adds, and a
    // reserve for candidates, hit lists and the query chunk
    const uint64_t need = (uint64_t)full.seq_len + 4ull * full.keys_count.size() * 65 + 4ull * full.positions.size() +
                          4ull * full.n_seqs + (n_dev > 1 ? (uint64_t)full.seq_len * n_dev : 0);
    // GHOSTM_B200_STREAM=1 forces the streaming path (tests)
    const bool fits = free_bytes > need + (12ull << 30) && !(n_dev == 1 && getenv("GHOSTM_B200_STREAM"));
    if (fits) {
      upload(shards[d].front, (uint32_t)c, full);
      resident[c] = 1;
      if (n_dev > 1)
        for (auto &sh : shards)
          check(gm_db_upload_seq(sh.back, (uint32_t)c, full.seq.data(), full.seq_len, full.seq_starts.data(),
                                 full.n_seqs), "gm_db_upload_seq");
    } else if (n_dev > 1) {
      throw DeviceError("db chunk " + std::to_string(c) + " does not fit the devices' memory (several devices keep "
                        "the whole db resident; use one device to stream it)");
    }
    names[c].names.swap(full.names);
    if (o.verbose)
      std::cout << "  db chunk " << c << " -> device " << o.devices[d] << (fits ? "" : " (streamed)") << std::endl;
  }
  // query chunks (query_reader.cpp:50-101, aligner.cpp:98-104,201-203)
  std::ifstream qinf((o.queries + ".inf").c_str(), std::ios::binary);
  int32_t q_division = 0;
  if (qinf) qinf.read(reinterpret_cast<char *>(&q_division), 4);
  uint32_t qi = o.start_chunk == UINT32_MAX ? 0u : o.start_chunk;
  QueryChunk q;
  if (q_division <= 0 || qi >= (uint32_t)q_division || !read_query_chunk(o.queries, qi, &q)) {
    std::cerr << "[Aligner] error: don't find query file." << std::endl;
  } else {
    const uint32_t cap = o.best > 1 ? o.best : 1;
    while (true) {
      std::vector<uint8_t> name_break(q.n, 0);
      for (uint32_t i = 1; i < q.n; ++i) name_break[i] = q.names[i] != q.names[i - 1];
      const uint64_t budget = std::min<uint64_t>((uint64_t)q.n * 2048 + (1u << 22), (1ull << 32) - 1);
      std::vector<gm_hit> hits((size_t)q.n * cap, gm_hit());
      std::vector<uint32_t> counts(q.n, 0);
      auto report = [&](uint32_t c, const gm_stats &st, double wall) {   // aligner.cpp:132-159
        std::cout << "|db chunk " << c << ": " << wall << " sec." << std::endl
                  << "|  Search alignment candidates ... " << st.ms_search * 1e-3f << " sec. (" << st.candidates
                  << " candidates in " << st.candidate_chunks << " chunk(s))" << std::endl
                  << "|  Calculate scores ... " << st.ms_score * 1e-3f << " sec." << std::endl
                  << "|  Merge results ... " << (st.ms_merge + st.ms_traceback) * 1e-3f << "sec" << std::endl;
      };
      if (n_dev == 1) {
        gm_context *ctx = shards[0].front;
        uint64_t capacity = std::max(budget, shards[0].capacity);
        shards[0].capacity = capacity;
        check(gm_set_candidate_capacity(ctx, capacity), "gm_set_candidate_capacity");
        check(gm_query_upload(ctx, q.seqs.data(), q.n, q.length, name_break.data()), "gm_query_upload");
        bool done = false;
        if (!o.verbose && std::find(resident.begin(), resident.end(), (char)0) == resident.end()) {
          // everything resident and no per-chunk report wanted: the whole batch is enqueued without a
          // host round trip (gm_align_chunk_async); gm_wait redoes it synchronously by itself when a
          // chunk needed a host decision.  A buffer limit sends us to the loop below, which grows it.
          for (uint32_t c = 0; c < n_chunks; ++c) check(gm_align_chunk_async(ctx, c), "gm_align_chunk_async");
          const int rc = gm_wait(ctx, nullptr);
          if (rc == 0) done = true;
          else if (rc != GM_ERR_CAPACITY) check(rc, "gm_wait");
          else check(gm_query_upload(ctx, q.seqs.data(), q.n, q.length, name_break.data()), "gm_query_upload");
        }
        for (uint32_t c = 0; c < n_chunks && !done; ++c) {
          if (!resident[c]) {     // streamed chunk: like the reference, read and uploaded per query chunk
            DbChunk full;
            if (!read_db_chunk(o.db, (int)c, &full, true)) throw std::runtime_error("db chunk file vanished");
            upload(ctx, c, full);
          }
          gm_stats st;
          memset(&st, 0, sizeof(st));
          const clock_t t0 = clock();
          with_capacity_retry(ctx, &shards[0].capacity, "gm_align_chunk", [&] {
            memset(&st, 0, sizeof(st));
            return gm_align_chunk(ctx, c, &st);
          });
          if (o.verbose) report(c, st, (double)(clock() - t0) / CLOCKS_PER_SEC);
          if (!resident[c]) check(gm_db_release(ctx, c), "gm_db_release");   // traces its pending hits first
        }
        check(gm_results_download(ctx, hits.data(), counts.data()), "gm_results_download");
      } else {
        const std::vector<uint32_t> bounds = slice_bounds(name_break, q.n, n_dev);
        on_every_device(shards, [&](size_t d) {
          Shard &sh = shards[d];
          const uint32_t base = bounds[d], stop = bounds[d + 1];
          sh.capacity = std::max(budget, sh.capacity);
          check(gm_set_candidate_capacity(sh.front, sh.capacity), "gm_set_candidate_capacity");
          check(gm_query_upload(sh.front, q.seqs.data(), q.n, q.length, name_break.data()), "gm_query_upload");
          if (stop > base) {
            check(gm_set_candidate_capacity(sh.back, sh.capacity), "gm_set_candidate_capacity");
            check(gm_query_upload(sh.back, q.seqs.data() + (size_t)base * q.length, stop - base, q.length,
                                  name_break.data() + base), "gm_query_upload");
          }
        });
        // Pipeline over the rounds of n_dev chunks: per device a FRONT thread (seed search + SW of the
        // owned chunk) and a BACK thread (candidate transfer + Merge calls of the device's query slice).
        // The front of round k+1 starts as soon as every back has FETCHED round k's candidates from
        // this front (gm_candidates_transfer reads the front context's buffers); the Merge calls of
        // round k run under it.  A front context is never used by two threads at once: transfers of a
        // round start after its front has finished and the next front waits for them.
        const uint32_t n_rounds = (n_chunks + (uint32_t)n_dev - 1) / (uint32_t)n_dev;
        std::vector<std::vector<std::vector<Segment>>> segs(n_dev, std::vector<std::vector<Segment>>(n_rounds));
        std::vector<std::vector<gm_stats>> round_stats(n_dev, std::vector<gm_stats>(n_rounds));
        std::vector<uint32_t> front_done(n_dev, 0);        // rounds whose front stage is finished, per device
        std::vector<uint32_t> fetched(n_dev, 0);           // transfers taken from this front, all rounds so far
        uint32_t n_backs = 0;
        for (size_t r = 0; r < n_dev; ++r) n_backs += bounds[r + 1] > bounds[r];
        std::mutex mu;
        std::condition_variable cv;
        bool failed = false;
        std::vector<std::thread> workers;
        auto guarded = [&](size_t d, const std::function<void()> &body) {
          try {
            body();
          } catch (std::exception &e) {
            std::lock_guard<std::mutex> lock(mu);
            if (shards[d].error.empty()) shards[d].error = e.what();
            failed = true;
            cv.notify_all();
          }
        };
        for (size_t d = 0; d < n_dev; ++d) {
          workers.emplace_back([&, d] {                    // front thread of device d
            guarded(d, [&] {
              for (uint32_t k = 0; k < n_rounds; ++k) {
                {
                  std::unique_lock<std::mutex> lock(mu);   // every back has taken round k-1 from this front
                  cv.wait(lock, [&] { return failed || fetched[d] >= k * n_backs; });
                  if (failed) return;
                }
                const uint32_t c = k * (uint32_t)n_dev + (uint32_t)d;
                memset(&round_stats[d][k], 0, sizeof(gm_stats));
                if (c < n_chunks)
                  segs[d][k] = front_chunk(shards[d].front, c, q.n, o.max_list_length, &shards[d].capacity,
                                           &round_stats[d][k]);
                std::lock_guard<std::mutex> lock(mu);
                front_done[d] = k + 1;
                cv.notify_all();
              }
            });
          });
          workers.emplace_back([&, d] {                    // back thread of device d: its slice, chunks ascending
            const size_t r = d;
            const uint32_t base = bounds[r], stop = bounds[r + 1];
            if (stop == base) return;
            guarded(r, [&] {
              for (uint32_t k = 0; k < n_rounds; ++k) {
                for (size_t s = 0; s < n_dev; ++s) {
                  const uint32_t c = k * (uint32_t)n_dev + (uint32_t)s;
                  {
                    std::unique_lock<std::mutex> lock(mu);
                    cv.wait(lock, [&] { return failed || front_done[s] > k; });
                    if (failed) return;
                  }
                  const bool work = c < n_chunks && !segs[s][k].empty();   // empty list: no Merge call (aligner.cpp:136)
                  if (work) {
                    std::lock_guard<std::mutex> lock(shards[s].front_mu);
                    const uint64_t back_cap = std::max(shards[r].capacity, shards[s].capacity);
                    check(gm_set_candidate_capacity(shards[r].back, back_cap), "gm_set_candidate_capacity");
                    check(gm_candidates_transfer(shards[s].front, shards[r].back, c, base, stop),
                          "gm_candidates_transfer");
                  }
                  {
                    std::lock_guard<std::mutex> lock(mu);
                    ++fetched[s];
                    cv.notify_all();
                  }
                  if (!work) continue;
                  for (const Segment &g : segs[s][k]) {      // one Merge call per candidate chunk
                    uint32_t f = std::min(std::max(g.first, base), stop), e = std::max(std::min(g.end, stop), base);
                    if (f >= e) f = e = base;                // carried lists only (aligner.cpp:702)
                    check(gm_merge(shards[r].back, f - base, e - base, nullptr), "gm_merge");
                  }
                }
              }
            });
          });
        }
        for (auto &w : workers) w.join();
        for (auto &sh : shards)
          if (!sh.error.empty()) throw DeviceError(sh.error);
        if (o.verbose)
          for (uint32_t k = 0; k < n_rounds; ++k)
            for (size_t d = 0; d < n_dev; ++d)
              if (k * n_dev + d < n_chunks) report(k * (uint32_t)n_dev + (uint32_t)d, round_stats[d][k], 0.0);
        on_every_device(shards, [&](size_t r) {            // TraceBack of the survivors, lists home
          const uint32_t base = bounds[r], stop = bounds[r + 1];
          if (stop == base) return;
          check(gm_results_download(shards[r].back, hits.data() + (size_t)base * cap, counts.data() + base),
                "gm_results_download");
          for (uint32_t i = base; i < stop; ++i)
            for (uint32_t k = 0; k < counts[i]; ++k) hits[(size_t)i * cap + k].query_id += base;
        });
      }
      write_hits(out, o, q, hits, counts, cap, names, (uint32_t)info.sum_length, karlin);
      ++qi;  // aligner.cpp:201-203
      if (qi > o.end_chunk || qi >= (uint32_t)q_division || !read_query_chunk(o.queries, qi, &q)) break;
    }
  }
  for (auto &sh : shards) {
    gm_destroy(sh.front);
    if (sh.back) gm_destroy(sh.back);
  }
  out.close();
  if (o.verbose) std::cout << "Complete." << std::endl;
  return 0;
}

// ---- `db`: FASTA -> <prefix>.inf, <prefix>_<i>.{inf,nam,seq,pos,ind} ----------------------------
// Same bytes as the reference's `ghostm db` (db_creator.cpp:369-479); the counting sort behind the
// index (db_creator.cpp:167-241, one host thread there) runs on the device (gm_db_build_index).
struct FastaReader {     // fasta_sequence_reader.cpp:34-86
  std::ifstream in;
  std::string last_line;
  explicit FastaReader(const std::string &path) : in(path.c_str()) {}
  bool next(std::string *name, std::string *residues) {
    if (in.eof()) return false;
    std::string line = last_line;
    while (!in.eof() && (line.empty() || line[0] != '>')) std::getline(in, line);
    if (in.eof()) return false;
    if (line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    name->clear();
    const size_t at = line.find_first_not_of("> ");
    if (at != std::string::npos) *name = line.substr(at);
    residues->clear();
    while (!in.eof()) {
      std::getline(in, line);
      if (line.empty()) continue;
      if (line[0] == '>') break;
      if (line[line.size() - 1] == '\r') line.erase(line.size() - 1);
      if (!line.empty() && line[line.size() - 1] == '+') line.erase(line.size() - 1);
      *residues += line;
    }
    last_line = line;
    return true;
  }
};

template <typename T>
void write_raw(std::ofstream &out, const T *p, size_t n) {
  out.write(reinterpret_cast<const char *>(p), (std::streamsize)(n * sizeof(T)));
}

int run_db(int argc, char **argv) {
  std::string input, prefix;
  uint32_t seed = (1u << 4) - 1, max_len = 1u << 27;     // db_creator.cpp:57-58
  int device = 0, c;
  while ((c = getopt(argc, argv, "i:o:k:l:D:")) >= 0) {
    switch (c) {
      case 'i': input = optarg; break;
      case 'o': prefix = optarg; break;
      case 'k': seed = (1u << atoi(optarg)) - 1; break;            // contiguous seed of weight k
      case 'l': max_len = (uint32_t)(atoi(optarg) * (1 << 20)); break;
      case 'D': device = atoi(optarg); break;                      // (not a reference option)
      default: throw std::invalid_argument("");
    }
  }
  std::cerr << "seed : " << seed << std::endl << "max length :" << max_len << std::endl;
  gm_context *ctx = nullptr;
  check(gm_create(device, &ctx), "gm_create");
  uint32_t weight = 0;
  for (uint32_t s = seed; s; s >>= 1) weight += s & 1;
  const uint32_t keys_count_len = (1u << (5 * weight)) + 1;

  FastaReader reader(input);
  std::string name, residues, next_name, next_residues;
  bool have_next = false;
  uint64_t sum_length = 0;
  int division = 0;
  for (;; ++division) {
    // db_creator.cpp:88-132: sequences are added while residues + separators fit max_len; the one
    // that does not fit opens the next chunk
    std::vector<std::string> names;
    std::vector<uint8_t> seq;
    std::vector<uint32_t> starts;
    uint32_t sum = 0;
    if (have_next) {
      sum = (uint32_t)next_residues.size() + 1;
      if (sum > max_len) {
        std::cerr << "error : too small max length." << std::endl;
        gm_destroy(ctx);
        return 1;
      }
      names.push_back(next_name);
      starts.push_back(0);
      for (char ch : next_residues) seq.push_back((uint8_t)residue_code(ch));
      seq.push_back(25);
      have_next = false;
    }
    while (reader.next(&name, &residues)) {
      sum += (uint32_t)residues.size() + 1;
      if (sum > max_len) {
        next_name = name;
        next_residues = residues;
        have_next = true;
        sum -= (uint32_t)residues.size() + 1;
        break;
      }
      names.push_back(name);
      starts.push_back((uint32_t)seq.size());
      for (char ch : residues) seq.push_back((uint8_t)residue_code(ch));
      seq.push_back(25);                                          // SEQUENCE_END, db_creator.cpp:138-159
    }
    std::cerr << "sum length : " << sum << std::endl;
    if (names.empty()) break;
    sum_length += seq.size() - names.size();
    std::ostringstream sub;
    sub << prefix << "_" << division;
    {
      std::ofstream out((sub.str() + ".inf").c_str(), std::ios::binary);
      const uint32_t n = (uint32_t)names.size(), len = (uint32_t)seq.size();
      write_raw(out, &n, 1);
      write_raw(out, &len, 1);
    }
    {
      std::ofstream out((sub.str() + ".nam").c_str());
      for (const std::string &nm : names) out << nm << std::endl;
    }
    {
      std::ofstream out((sub.str() + ".seq").c_str(), std::ios::binary);
      write_raw(out, seq.data(), seq.size());
    }
    {
      std::ofstream out((sub.str() + ".pos").c_str(), std::ios::binary);
      write_raw(out, starts.data(), starts.size());
    }
    // the index: keys, stable counting sort and CSR boundaries on the device
    check(gm_db_build_index(ctx, 0, seq.data(), (uint32_t)seq.size(), starts.data(), (uint32_t)starts.size(), seed),
          "gm_db_build_index");
    std::vector<uint32_t> keys_count(keys_count_len), positions(seq.size());
    uint32_t n_pos = 0;
    check(gm_db_download_index(ctx, 0, keys_count.data(), positions.data(), &n_pos), "gm_db_download_index");
    {
      std::ofstream out((sub.str() + ".ind").c_str(), std::ios::binary);   // db_creator.cpp:332-352
      write_raw(out, &seed, 1);
      write_raw(out, &keys_count_len, 1);
      write_raw(out, &n_pos, 1);
      write_raw(out, keys_count.data(), keys_count.size());
      write_raw(out, positions.data(), n_pos);
    }
    check(gm_db_release(ctx, 0), "gm_db_release");
    std::cerr << "number sequences : " << names.size() << std::endl;
  }
  std::cerr << "[db] division number : " << division << std::endl;
  {
    std::ofstream out((prefix + ".inf").c_str(), std::ios::binary);       // db_creator.cpp:243-264
    const int32_t div = division;
    write_raw(out, &div, 1);
    write_raw(out, &seed, 1);
    write_raw(out, &max_len, 1);
    write_raw(out, &sum_length, 1);
    for (int i = 0; i < 32; ++i) write_raw(out, &div, 1);
  }
  gm_destroy(ctx);
  return 0;
}

void usage() {
  std::cerr << "usage: ghostm_b200 aln [-i queries] [-d database] [-o output] [-D device[,device...]]\n"
               "          [-v] [-b best] [-G openGap] [-E extendGap] [-M scoreMatrix] [-l candidatesMB]\n"
               "          [-s skip] [-t threshold] [-r regionSize] [-e extendSize] [-S first] [-L last] [-y style]\n"
               "       ghostm_b200 db [-i fasta] [-o database] [-k seedWeight] [-l chunkMB] [-D device]\n"
               "   `db` writes the reference's files byte for byte (index built on the GPU); `qry` formatting\n"
               "   stays with the reference tool; the file formats are unchanged.\n";
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2 || (strcmp(argv[1], "aln") != 0 && strcmp(argv[1], "db") != 0)) {
    usage();
    return 1;
  }
  const bool is_db = strcmp(argv[1], "db") == 0;
  try {  // main.cpp:107-121: usage errors are reported, the exit status stays 0 like the reference
    return is_db ? run_db(argc - 1, argv + 1) : run_aln(argc - 1, argv + 1);
  } catch (DeviceError &e) {
    // a device-side failure (CUDA error, out of memory, buffer limit) must not look like a finished
    // run: the partial output file is removed and the status is non-zero
    std::cerr << "error: " << e.what() << std::endl;
    if (!g_output_path.empty()) remove(g_output_path.c_str());
    return 2;
  } catch (std::exception &e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 0;
  }
}
