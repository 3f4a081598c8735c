// wsr_replay — the query-log replay driver for the GPU engine: the counterpart of the
// reference's engine_bench (src/qq_mem/src/engine_bench.cc) in its in-process log modes.
// Same log format and query defaults (one query per line, space-separated analysed terms, a
// quoted line is a phrase; n_results=5, return_snippets=false: query_pool.h:251-335,
// types.h:215-218). The reference's `locallog` loop never terminates (IsEnd() is hard-wired
// false, query_pool.h:357-360 vs engine_bench.cc:261); this driver replays the log once per
// -repeat and stops.
//
//   -engine=gpu:vacuum_dump:<dir>   engine URL (engine_factory.h scheme + gpu:)
//   -query_path=<file>              query log
//   -exp_mode=batchlog|locallog     batchlog: whole log in batches of -batch_size lines through
//                                             wsr_search_log (text -> GPU front end); the pass that
//                                             writes -dump uses wsr_search_batch (it needs doc_freqs)
//                                   locallog: n_threads client threads calling Search()
//                                             (LocalLogTreatmentExecutor, engine_bench.cc:214-290)
//   -n_threads=N  -n_results=K  -batch_size=B  -repeat=R  -dump=<file>
//   -dirs=<p0>,<p1>,..  -devices=0,1,..   instead of -engine: a document-partitioned collection, one
//                                   vacuum directory per partition, spread over the listed GPUs of
//                                   this process (wsr_group_*; batchlog only)
// -dump writes one line per query: <n_entries> <n_doc_freqs> {<doc> <score %a>}* {<df>}*
#include <atomic>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "gpu_vacuum_engine.h"

namespace {

std::string FlagStr(int argc, char **argv, const char *name, const std::string &def) {
  const std::string key = std::string("-") + name + "=";
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a.compare(0, 2, "--") == 0) a = a.substr(1);
    if (a.compare(0, key.size(), key) == 0) return a.substr(key.size());
  }
  return def;
}

bool ReadFile(const std::string &path, std::string *out) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return false;
  std::ostringstream ss;
  ss << in.rdbuf();
  *out = ss.str();
  return true;
}

// QueryProducerByLog::GetTerms / IsPhrase, query_pool.h:337-352
SearchQuery ParseLine(std::string line, int n_results) {
  size_t a = 0, b = line.size();
  while (a < b && isspace((unsigned char)line[a])) a++;
  while (b > a && isspace((unsigned char)line[b - 1])) b--;
  line = line.substr(a, b - a);
  SearchQuery q;
  q.n_results = n_results;
  q.is_phrase = line.size() >= 1 && line.front() == '"' && line.back() == '"';
  if (q.is_phrase) {
    line.erase(line.size() - 1, 1);
    if (!line.empty()) line.erase(0, 1);
  }
  std::string buf;
  for (char c : line) {
    if (c != ' ') buf += c;
    else if (!buf.empty()) { q.terms.push_back(buf); buf.clear(); }
  }
  if (!buf.empty()) q.terms.push_back(buf);
  return q;
}

std::vector<std::string> Split(const std::string &s, char sep) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == sep) { if (!cur.empty()) out.push_back(cur); cur.clear(); }
    else cur += c;
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

// -dirs=<p0>,<p1>,... [-devices=0,1,...]: a document-partitioned collection (one vacuum directory
// per partition, local doc ids) served by one process over one or more GPUs through the group API
// (wsr_group_search_log: per-partition search, on-device merge, NCCL exchange between devices).
// The dump has the format of the single-index one, with global doc ids.
int ReplayGroup(int argc, char **argv, const std::string &dirs_flag, const std::string &text, int n_results,
                int batch_size, int repeat, const std::string &dump) {
  const std::vector<std::string> dirs = Split(dirs_flag, ',');
  std::vector<int> devices;
  for (const std::string &d : Split(FlagStr(argc, argv, "devices", "0"), ',')) devices.push_back(atoi(d.c_str()));
  std::vector<const char *> dptr;
  for (const std::string &d : dirs) dptr.push_back(d.c_str());
  char err[512] = {0};
  const auto t_load = std::chrono::steady_clock::now();
  wsr_group *g = wsr_group_open(dptr.data(), (int)dptr.size(), devices.data(), (int)devices.size(), 0, 0u, nullptr,
                                err, sizeof(err));
  if (!g) { fprintf(stderr, "wsr_group_open: %s\n", err); return 1; }
  const double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
  std::vector<size_t> line_at;
  for (size_t p = 0; p < text.size();) {
    line_at.push_back(p);
    const size_t e = text.find('\n', p);
    p = e == std::string::npos ? text.size() : e + 1;
  }
  const int n = (int)line_at.size();
  line_at.push_back(text.size());
  const int B = batch_size > 0 ? std::min(batch_size, n) : n;
  wsr_hit *hits = (wsr_hit *)wsr_host_alloc((size_t)B * n_results * sizeof(wsr_hit));
  int32_t *n_hits = (int32_t *)wsr_host_alloc((size_t)B * 4);
  uint32_t *dfs = (uint32_t *)wsr_host_alloc((size_t)B * WSR_MAX_TERMS * 4);
  int32_t *ndf = (int32_t *)wsr_host_alloc((size_t)B * 4);
  if (!hits || !n_hits || !dfs || !ndf) { fprintf(stderr, "pinned alloc failed\n"); return 1; }
  FILE *df = dump.empty() ? nullptr : fopen(dump.c_str(), "w");
  uint64_t n_queries = 0, listed = 0, entries = 0;
  const auto t0 = std::chrono::steady_clock::now();
  double best_rep = 1e30;
  for (int rep = 0; rep < repeat; rep++) {
    const auto t_rep = std::chrono::steady_clock::now();
    for (int lo = 0; lo < n; lo += B) {
      const int m = std::min(B, n - lo);
      int got = 0;
      if (wsr_group_search_log(g, text.data() + line_at[lo], line_at[lo + m] - line_at[lo], n_results, hits, n_hits,
                               dfs, ndf, B, &got) != 0 || got != m) {
        fprintf(stderr, "wsr_group_search_log: %s\n", wsr_last_error());
        return 1;
      }
      for (int i = 0; i < m; i++) {
        entries += n_hits[i];
        for (int t = 0; t < ndf[i]; t++) listed += dfs[(size_t)i * WSR_MAX_TERMS + t];
        if (df && rep == 0) {
          fprintf(df, "%d %d", n_hits[i], ndf[i]);
          for (int j = 0; j < n_hits[i]; j++)
            fprintf(df, " %d %a", hits[(size_t)i * n_results + j].doc_id, hits[(size_t)i * n_results + j].score);
          for (int t = 0; t < ndf[i]; t++) fprintf(df, " %u", dfs[(size_t)i * WSR_MAX_TERMS + t]);
          fprintf(df, "\n");
        }
      }
      n_queries += m;
    }
    best_rep = std::min(best_rep, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_rep).count());
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (df) fclose(df);
  wsr_host_free(hits); wsr_host_free(n_hits); wsr_host_free(dfs); wsr_host_free(ndf);
  wsr_group_close(g);
  printf("WSR_REPLAY_JSON {\"mode\": \"group\", \"partitions\": %zu, \"devices\": %zu, \"queries\": %" PRIu64
         ", \"seconds\": %.6f, \"qps\": %.3f, \"listed_postings\": %" PRIu64 ", \"listed_postings_per_s\": %.3f, "
         "\"result_entries\": %" PRIu64 ", \"load_seconds\": %.3f, \"best_pass_seconds\": %.6f, \"best_pass_qps\": %.3f}\n",
         dirs.size(), devices.size(), n_queries, secs, n_queries / secs, listed, listed / secs, entries, load_s,
         best_rep, (double)n / best_rep);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  const std::string engine_url = FlagStr(argc, argv, "engine", "");
  const std::string query_path = FlagStr(argc, argv, "query_path", "");
  const std::string mode = FlagStr(argc, argv, "exp_mode", "batchlog");
  const std::string dump = FlagStr(argc, argv, "dump", "");
  const int n_threads = atoi(FlagStr(argc, argv, "n_threads", "1").c_str());
  const int n_results = atoi(FlagStr(argc, argv, "n_results", "5").c_str());
  const int batch_size = atoi(FlagStr(argc, argv, "batch_size", "65536").c_str());
  const int repeat = atoi(FlagStr(argc, argv, "repeat", "1").c_str());
  if ((engine_url.empty() && FlagStr(argc, argv, "dirs", "").empty()) || query_path.empty()) {
    fprintf(stderr, "usage: wsr_replay -engine=gpu:vacuum_dump:<dir> -query_path=<log> "
                    "[-exp_mode=batchlog|locallog] [-n_threads=N] [-n_results=K] "
                    "[-batch_size=B] [-repeat=R] [-dump=<file>]\n");
    return 2;
  }
  std::string text;
  if (!ReadFile(query_path, &text)) {
    fprintf(stderr, "File may not exist: %s\n", query_path.c_str());   // QueryLogReader, query_pool.h:18-23
    return 1;
  }
  const std::string dirs_flag = FlagStr(argc, argv, "dirs", "");
  if (!dirs_flag.empty() && mode != "locallog")
    return ReplayGroup(argc, argv, dirs_flag, text, n_results, batch_size, repeat, dump);
  std::unique_ptr<SearchEngineServiceNew> engine;
  if (!dirs_flag.empty()) {
    // Search() callers over a partitioned collection: the adapter in group mode
    wsr::GpuEngineOptions opt;
    opt.partition_dirs = Split(dirs_flag, ',');
    for (const std::string &d : Split(FlagStr(argc, argv, "devices", "0"), ',')) opt.devices.push_back(atoi(d.c_str()));
    opt.load_positions = false;
    engine.reset(new wsr::GpuVacuumEngine("group", 1, opt));
  } else {
    engine = wsr::CreateSearchEngine(engine_url);
  }
  const auto t_load = std::chrono::steady_clock::now();
  engine->Load();
  const double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
  auto *gpu = static_cast<wsr::GpuVacuumEngine *>(engine.get());
  wsr_index *idx = gpu->handle();

  uint64_t n_queries = 0, listed = 0, entries = 0;
  double secs = 0;
  FILE *df = dump.empty() ? nullptr : fopen(dump.c_str(), "w");

  if (mode == "batchlog") {
    size_t n_lines = 1;
    for (char c : text) n_lines += c == '\n';
    std::vector<wsr_query> qs(n_lines + 1);
    int n = 0;
    if (wsr_parse_query_log(idx, text.data(), text.size(), n_results, qs.data(), (int)qs.size(), &n) != 0) {
      fprintf(stderr, "parse: %s\n", wsr_last_error());
      return 1;
    }
    const int B = batch_size > 0 ? batch_size : n;
    wsr_hit *hits = (wsr_hit *)wsr_host_alloc((size_t)B * n_results * sizeof(wsr_hit));
    int32_t *n_hits = (int32_t *)wsr_host_alloc((size_t)B * 4);
    std::vector<uint32_t> dfs((size_t)B * WSR_MAX_TERMS);
    std::vector<int32_t> ndf(B);
    if (!hits || !n_hits) { fprintf(stderr, "pinned alloc failed\n"); return 1; }
    // line starts, so that a batch of B queries is a contiguous piece of the log text
    std::vector<size_t> line_at;
    for (size_t p = 0; p < text.size();) {
      line_at.push_back(p);
      const size_t e = text.find('\n', p);
      p = e == std::string::npos ? text.size() : e + 1;
    }
    line_at.push_back(text.size());
    // listed postings of one pass (sum of the df of every term of every answerable query)
    uint64_t listed_per_rep = 0;
    for (int i = 0; i < n; i++) {
      const wsr_query &q = qs[i];
      bool ok = q.k > 0 && q.n_terms > 0;
      for (uint32_t t = 0; ok && t < q.n_terms; t++) ok = q.term_ids[t] != WSR_TERM_ABSENT;
      for (uint32_t t = 0; ok && t < q.n_terms; t++) {
        uint32_t df_t = 0;
        wsr_term_at(idx, q.term_ids[t], nullptr, 0, &df_t);
        listed_per_rep += df_t;
      }
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < repeat; rep++) {
      const bool dumping = df && rep == 0;
      for (int lo = 0; lo < n; lo += B) {
        const int m = std::min(B, n - lo);
        if (dumping) {
          // doc_freqs are wanted: host-parsed queries through wsr_search_batch
          if (wsr_search_batch(idx, qs.data() + lo, m, n_results, hits, n_hits, dfs.data(), ndf.data()) != 0) {
            fprintf(stderr, "search: %s\n", wsr_last_error());
            return 1;
          }
        } else {
          // timing passes: the text itself goes to the GPU (parse + lookup + plan as kernels)
          int got = 0;
          if (wsr_search_log(idx, text.data() + line_at[lo], line_at[lo + m] - line_at[lo], n_results, hits,
                             n_hits, B, &got) != 0 || got != m) {
            fprintf(stderr, "search_log: %s\n", wsr_last_error());
            return 1;
          }
        }
        for (int i = 0; i < m; i++) {
          entries += n_hits[i];
          if (dumping) {
            fprintf(df, "%d %d", n_hits[i], ndf[i]);
            for (int j = 0; j < n_hits[i]; j++)
              fprintf(df, " %d %a", hits[(size_t)i * n_results + j].doc_id, hits[(size_t)i * n_results + j].score);
            for (int t = 0; t < ndf[i]; t++) fprintf(df, " %u", dfs[(size_t)i * WSR_MAX_TERMS + t]);
            fprintf(df, "\n");
          }
        }
        n_queries += m;
      }
      listed += listed_per_rep;
    }
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    wsr_host_free(hits);
    wsr_host_free(n_hits);
  } else if (mode == "locallog") {
    std::vector<SearchQuery> qs;
    {
      std::istringstream in(text);
      std::string line;
      while (std::getline(in, line)) qs.push_back(ParseLine(line, n_results));
    }
    std::vector<SearchResult> results(df ? qs.size() : 0);
    const int T = std::max(1, n_threads);
    std::vector<uint64_t> l_t(T, 0), e_t(T, 0);
    const auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < repeat; rep++) {
      std::vector<std::thread> th;
      for (int t = 0; t < T; t++) {
        th.emplace_back([&, t, rep]() {
          // QueryProducerByLog hands query i to thread i % n_threads (query_pool.h:326-329)
          for (size_t i = t; i < qs.size(); i += T) {
            SearchResult r = engine->Search(qs[i]);
            for (int d : r.doc_freqs) l_t[t] += d;
            e_t[t] += r.entries.size();
            if (df && rep == 0) results[i] = std::move(r);
          }
        });
      }
      for (auto &x : th) x.join();
      n_queries += qs.size();
    }
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < T; t++) { listed += l_t[t]; entries += e_t[t]; }
    if (df)
      for (auto &r : results) {
        fprintf(df, "%zu %zu", r.entries.size(), r.doc_freqs.size());
        for (auto &e : r.entries) fprintf(df, " %d %a", e.doc_id, e.doc_score);
        for (int d : r.doc_freqs) fprintf(df, " %d", d);
        fprintf(df, "\n");
      }
  } else {
    fprintf(stderr, "unknown exp_mode %s\n", mode.c_str());
    return 2;
  }
  if (df) fclose(df);
  printf("WSR_REPLAY_JSON {\"mode\": \"%s\", \"queries\": %" PRIu64 ", \"seconds\": %.6f, \"qps\": %.3f, "
         "\"listed_postings\": %" PRIu64 ", \"listed_postings_per_s\": %.3f, \"result_entries\": %" PRIu64
         ", \"threads\": %d, \"load_seconds\": %.3f}\n",
         mode.c_str(), n_queries, secs, n_queries / secs, listed, listed / secs, entries, n_threads,
         load_s);
  return 0;
}
