// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Driver around the UNMODIFIED reference engine (compiled from the sources where
// they lie under /root/reference/src/qq_mem/src, see oracle/Makefile). It only
// calls the reference's own public classes:
//   build     linedoc -> vacuum dir       (VacuumInvertedIndexDumper, DocLengthCharStore,
//                                          ChunkedDocStoreDumper; flash_engine_dumper.h:263-830)
//   buildbloom linedoc (with bloom columns) -> vacuum dir carrying the Bloom-begin/-end sections,
//             the pipeline of tests_18.cc:283-310: QqMemEngineDelta::Serialize + BloomDumper +
//             FlashEngineDumper(dir, true).LoadQqMemDump().Dump()
//   replay    query log -> top-k results  (CreateSearchEngine + VacuumEngine::Search,
//                                          engine_factory.h:33-50, vacuum_engine.h:201-258;
//                                          log parsing = QueryProducerNoLoop, query_pool.h:251-311)
//   dumplists every posting list -> (doc, tf) pairs through VacuumPostingListIterator
//                                          (flash_iterators.h:893-1079)
//   time      multi-threaded timed replay against ONE shared engine (the reference's own
//             concurrency model, grpc_server_impl.h:309-328) -> JSON line
//
// Output formats are ours (documented at each writer) and are consumed by tests/ and bench.py.
#include <atomic>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <thread>

#include <gflags/gflags.h>
#include <glog/logging.h>

int FLAGS_minloglevel = 2;
int FLAGS_logtostderr = 1;

#include "bloom_filter.h"
#include "engine_factory.h"
#include "flash_engine_dumper.h"
#include "query_pool.h"
#include "vacuum_engine.h"

DECLARE_string(lock_memory);
DECLARE_bool(enable_prefetch);

namespace {

// Same steps as FlashEngineDumper::{LoadLocalDocuments,Dump} (flash_engine_dumper.h:686-754),
// re-hosted because VacuumInvertedIndexDumper::Dump() calls utils::FormatThousands, which
// constructs std::locale("en_US.UTF-8") — absent in this image — at the 10 000th posting list.
class IndexDumper : public VacuumInvertedIndexDumper {
 public:
  explicit IndexDumper(const std::string &dir) : VacuumInvertedIndexDumper(dir) {}
  void DumpAll(const std::string &terms_path) {
    DumpHeader();
    FILE *tf = fopen(terms_path.c_str(), "w");
    for (auto it = index_.cbegin(); it != index_.cend(); ++it) {
      DumpPostingList(it->first, it->second);
      if (tf) fprintf(tf, "%s %d\n", it->first.c_str(), (int)it->second.Size());
    }
    if (tf) fclose(tf);
  }
};

int CmdBuild(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: ref_tool build <linedoc> <out_dir> [n_rows]\n");
    return 2;
  }
  const std::string linedoc = argv[2], dir = argv[3];
  const int n_rows = argc > 4 ? atoi(argv[4]) : 100000000;
  utils::PrepareDir(dir);

  IndexDumper index(dir);
  DocLengthCharStore doc_lengths;
  ChunkedDocStoreDumper doc_store(false);

  LineDocParserPosition parser(linedoc, n_rows);
  DocInfo doc_info;
  int doc_id = 0;
  while (parser.Pop(&doc_info)) {
    doc_store.Add(doc_id, doc_info.Body());
    index.AddDocument(doc_id, doc_info);
    doc_lengths.AddLength(doc_id, doc_info.BodyLength());
    doc_id++;
  }
  index.DumpAll(dir + "/terms.txt");
  doc_lengths.Serialize(utils::JoinPath(dir, "my.doc_length"));
  doc_store.Dump(dir + "/my.fdx", dir + "/my.fdt");
  remove((dir + "/fake.vacuum").c_str());
  fprintf(stderr, "built %d docs, %d terms -> %s\n", doc_id, (int)index.Size(), dir.c_str());
  return 0;
}

int CmdBuildBloom(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: ref_tool buildbloom <linedoc with bloom columns> <out_dir>\n");
    return 2;
  }
  const std::string linedoc = argv[2], dir = argv[3], tmp = dir + ".qqmem";
  utils::RemoveDir(tmp);
  utils::PrepareDir(dir);
  auto engine = CreateSearchEngine("qq_mem_compressed");
  engine->LoadLocalDocuments(linedoc, 100000000, "WITH_POSITIONS");
  engine->Serialize(tmp);
  BloomDumper bloom_dumper;
  bloom_dumper.Load(linedoc);
  bloom_dumper.Dump(tmp);
  FlashEngineDumper engine_dumper(dir, true);
  engine_dumper.LoadQqMemDump(tmp);
  engine_dumper.Dump();
  utils::RemoveDir(tmp);
  remove((dir + "/fake.vacuum").c_str());
  fprintf(stderr, "built bloom-enabled index, %d terms -> %s\n", engine->TermCount(), dir.c_str());
  return 0;
}

std::vector<SearchQuery> LoadQueries(const std::string &path, int k) {
  QueryProducerNoLoop producer(path);
  std::vector<SearchQuery> qs;
  while (!producer.IsEnd()) {
    SearchQuery q = producer.NextNativeQuery(0);
    q.n_results = k;
    q.return_snippets = false;
    qs.push_back(q);
  }
  return qs;
}

std::unique_ptr<SearchEngineServiceNew> LoadEngine(const std::string &dir, int bloom_factor) {
  FLAGS_lock_memory = "disabled";
  FLAGS_enable_prefetch = false;
  auto engine = CreateSearchEngine("vacuum:vacuum_dump:" + dir, bloom_factor);
  engine->Load();
  return engine;
}

// Result file, one line per query:
//   <n_entries> <n_doc_freqs> {<doc_id> <score as %a hex double>}* {<df>}*
int CmdReplay(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: ref_tool replay <dir> <query_log> <k> <out> [bloom_factor]\n");
    return 2;
  }
  const int k = atoi(argv[4]);
  const int bloom_factor = argc > 6 ? atoi(argv[6]) : 1;
  auto engine = LoadEngine(argv[2], bloom_factor);
  auto qs = LoadQueries(argv[3], k);
  FILE *out = fopen(argv[5], "w");
  if (!out) return 1;
  for (auto &q : qs) {
    SearchResult r = engine->Search(q);
    fprintf(out, "%zu %zu", r.entries.size(), r.doc_freqs.size());
    for (auto &e : r.entries) fprintf(out, " %d %a", e.doc_id, e.doc_score);
    for (auto df : r.doc_freqs) fprintf(out, " %d", df);
    fprintf(out, "\n");
  }
  fclose(out);
  return 0;
}

// Binary dump: repeated { u32 term_len; bytes term; u32 df; df x { u32 doc; u32 tf } },
// in my.tip order.
int CmdDumpLists(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: ref_tool dumplists <dir> <out.bin>\n");
    return 2;
  }
  const std::string dir = argv[2];
  VacuumInvertedIndex index(dir + "/my.tip", dir + "/my.vacuum");
  utils::FileMap tip;
  tip.Open(dir + "/my.tip");
  const char *p = tip.Addr(), *end = p + tip.Length();
  FILE *out = fopen(argv[3], "wb");
  if (!out) return 1;
  while (p < end) {
    uint32_t len = *(const uint32_t *)p;
    std::string term(p + 4, len);
    p += 4 + len + 8;
    auto iters = index.FindIteratorsSolid({term});
    if (iters.size() != 1) return 3;
    auto &it = iters[0];
    uint32_t df = it.Size();
    fwrite(&len, 4, 1, out);
    fwrite(term.data(), 1, len, out);
    fwrite(&df, 4, 1, out);
    while (!it.IsEnd()) {
      uint32_t rec[2] = {(uint32_t)it.DocId(), (uint32_t)it.TermFreq()};
      fwrite(rec, 4, 2, out);
      it.Advance();
    }
  }
  fclose(out);
  tip.Close();
  return 0;
}

// Timed replay: T threads pulling disjoint slices of the log and calling Search against ONE
// shared engine; best wall time of <reps>. Prints one JSON line on stdout.
int CmdTime(int argc, char **argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: ref_tool time <dir> <query_log> <k> <threads> <reps> [max_queries]\n");
    return 2;
  }
  const int k = atoi(argv[4]), T = atoi(argv[5]), reps = atoi(argv[6]);
  auto engine = LoadEngine(argv[2], 1);
  auto qs = LoadQueries(argv[3], k);
  if (argc > 7 && (size_t)atol(argv[7]) < qs.size()) qs.resize(atol(argv[7]));
  const size_t n = qs.size();
  double best = 1e30;
  uint64_t listed = 0, hits = 0;
  std::string rep_secs;
  for (int rep = 0; rep < reps; rep++) {
    std::vector<uint64_t> listed_t(T, 0), hits_t(T, 0);
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) {
      th.emplace_back([&, t]() {
        // dynamic slices of 16 queries keep all threads busy to the end (query cost varies 10^4x)
        uint64_t l = 0, h = 0;
        for (;;) {
          const size_t b = next.fetch_add(16);
          if (b >= n) break;
          const size_t e = b + 16 < n ? b + 16 : n;
          for (size_t i = b; i < e; i++) {
            SearchResult r = engine->Search(qs[i]);
            for (auto df : r.doc_freqs) l += df;
            h += r.entries.size();
          }
        }
        listed_t[t] = l;
        hits_t[t] = h;
      });
    }
    for (auto &x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (s < best) best = s;
    rep_secs += (rep ? ", " : "") + std::to_string(s);
    listed = hits = 0;
    for (int t = 0; t < T; t++) listed += listed_t[t], hits += hits_t[t];
  }
  // keep the engine's chatter off stdout's last line
  fflush(stdout);
  printf("\nREF_TIME_JSON {\"queries\": %zu, \"threads\": %d, \"seconds\": %.6f, \"qps\": %.3f, "
         "\"listed_postings\": %" PRIu64 ", \"listed_postings_per_s\": %.3f, \"result_entries\": %" PRIu64
         ", \"rep_seconds\": [%s]}\n",
         n, T, best, n / best, listed, listed / best, hits, rep_secs.c_str());
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: ref_tool {build|replay|dumplists|time} ...\n");
    return 2;
  }
  std::string cmd = argv[1];
  if (cmd == "build") return CmdBuild(argc, argv);
  if (cmd == "buildbloom") return CmdBuildBloom(argc, argv);
  if (cmd == "replay") return CmdReplay(argc, argv);
  if (cmd == "dumplists") return CmdDumpLists(argc, argv);
  if (cmd == "time") return CmdTime(argc, argv);
  fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 2;
}
