// wsr_index_linedoc — native, multi-threaded indexer: a WITH_POSITIONS linedoc file (the input of
// the reference's create_qq_mem_dump) straight to a vacuum index directory in the reference's
// on-disk format, replacing the single-threaded, whole-corpus-in-RAM pipeline
// linedoc -> create_qq_mem_dump -> convert_qq_to_vacuum. The directory is read
// by the unmodified reference engine, the CPU oracle and the GPU loader alike.
//
// Input semantics follow the reference's loader, statement by statement:
//   row     = explode_strict(line, '\t'): title, body, tokenized, offsets, positions
//             (utils.h:69-79, engine_loader.h:84-97); the first line is the column header
//   terms   = explode(tokenized, ' ')  — the document's distinct analysed terms (types.cc:4-6)
//   tf      = number of "start,end;" pairs of the term's offsets group "s,e;s,e;." — NOT the
//             number of positions (qq_mem_engine.h:194-215, utils.cc:105-140)
//   positions group "p;p;." per term (types.cc:17-36); its size must equal tf because the
//             position bags are delimited by tf (flash_iterators.h:619-628)
//   doc length = number of space-separated pieces of the BODY (types.cc:38-40, utils.cc:163-165)
//   doc ids = 0..N-1 in file order (flash_engine_dumper.h:714-721); average length is the
//             running mean of doc_length_store.h:104-112
// Terms are written in bytewise-sorted order (the reference's order is that of an unordered_map;
// readers do not depend on it).
//
//   wsr_index_linedoc --linedoc FILE --out DIR [--rows N] [--threads T] [--positions 0|1]
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "vacuum_writer.h"

namespace {
using namespace wsrw;

struct TermAcc {                 // postings of one term inside one doc range, doc ascending
  std::vector<uint32_t> docs, tfs, pos;
};

struct Part {                    // one contiguous range of documents
  size_t doc_begin = 0, doc_end = 0;
  std::unordered_map<std::string_view, uint32_t> ids;   // views into the mapped file
  std::vector<std::string_view> names;
  std::vector<TermAcc> acc;
  std::vector<uint32_t> doc_len;
  std::string err;
};

// the i-th TAB-separated field of [b, e), like explode_strict
std::string_view Field(const char *b, const char *e, int i) {
  const char *p = b;
  for (int k = 0; k < i; k++) {
    p = (const char *)memchr(p, '\t', e - p);
    if (!p) return std::string_view();
    p++;
  }
  const char *q = (const char *)memchr(p, '\t', e - p);
  return std::string_view(p, (q ? q : e) - p);
}

// explode(s, c): non-empty pieces
template <typename F>
void ForEachPiece(std::string_view s, char c, F fn) {
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && s[i] == c) i++;
    size_t j = i;
    while (j < s.size() && s[j] != c) j++;
    if (j > i) fn(s.substr(i, j - i));
    i = j;
  }
}

// groups "....." terminated by '.', empty groups dropped, an unterminated tail dropped
// (utils::parse_offsets, utils.cc:124-140)
template <typename F>
void ForEachDotGroup(std::string_view s, F fn) {
  size_t i = 0;
  while (i < s.size()) {
    size_t j = i;
    while (j < s.size() && s[j] != '.') j++;
    if (j >= s.size()) break;
    if (j > i) fn(s.substr(i, j - i));
    i = j + 1;
  }
}

bool ParseDoc(const char *b, const char *e, uint32_t doc, bool want_pos, Part *part,
              std::vector<uint32_t> *tf_scratch, std::vector<std::vector<uint32_t>> *pos_scratch) {
  const std::string_view body = Field(b, e, 1), toks = Field(b, e, 2), offs = Field(b, e, 3),
                         poss = Field(b, e, 4);
  uint32_t len = 0;
  ForEachPiece(body, ' ', [&](std::string_view) { len++; });
  part->doc_len.push_back(len);
  // tf per term: pairs terminated by ';' inside each '.'-terminated group
  tf_scratch->clear();
  ForEachDotGroup(offs, [&](std::string_view g) {
    uint32_t n = 0;
    size_t i = 0;
    while (i < g.size()) {
      size_t j = i;
      while (j < g.size() && g[j] != ';') j++;
      if (j >= g.size()) break;           // unterminated pair is dropped (handle_term_offsets)
      if (j > i) n++;
      i = j + 1;
    }
    tf_scratch->push_back(n);
  });
  size_t n_groups = 0;
  if (want_pos) {
    ForEachPiece(poss, '.', [&](std::string_view g) {      // explode(positions, '.')
      if (pos_scratch->size() <= n_groups) pos_scratch->emplace_back();
      std::vector<uint32_t> &v = (*pos_scratch)[n_groups++];
      v.clear();
      ForEachPiece(g, ';', [&](std::string_view p) {
        uint32_t x = 0;
        for (char c : p) {
          if (c < '0' || c > '9') { part->err = "non-numeric position"; return; }
          x = x * 10 + (uint32_t)(c - '0');
        }
        v.push_back(x);
      });
    });
  }
  size_t t = 0;
  bool ok = true;
  ForEachPiece(toks, ' ', [&](std::string_view term) {
    if (!ok) return;
    if (t >= tf_scratch->size()) { part->err = "fewer offset groups than terms"; ok = false; return; }
    const uint32_t tf = (*tf_scratch)[t];
    if (tf == 0) { part->err = "term without occurrences"; ok = false; return; }
    auto it = part->ids.find(term);
    uint32_t id;
    if (it == part->ids.end()) {
      id = (uint32_t)part->names.size();
      part->ids.emplace(term, id);
      part->names.push_back(term);
      part->acc.emplace_back();
    } else {
      id = it->second;
    }
    TermAcc &a = part->acc[id];
    if (!a.docs.empty() && a.docs.back() == doc) { part->err = "term listed twice in one document"; ok = false; return; }
    a.docs.push_back(doc);
    a.tfs.push_back(tf);
    if (want_pos) {
      if (t >= n_groups || (*pos_scratch)[t].size() != tf) {
        part->err = "positions of a term do not match its tf";
        ok = false;
        return;
      }
      const std::vector<uint32_t> &pv = (*pos_scratch)[t];
      for (size_t i = 1; i < pv.size(); i++)
        if (pv[i] <= pv[i - 1]) { part->err = "positions not ascending"; ok = false; return; }
      a.pos.insert(a.pos.end(), pv.begin(), pv.end());
    }
    t++;
  });
  return ok && part->err.empty();
}

}  // namespace

int main(int argc, char **argv) {
  std::string linedoc, out;
  size_t rows = SIZE_MAX;
  int threads = 0;
  bool positions = true;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--linedoc") linedoc = v;
    else if (k == "--out") out = v;
    else if (k == "--rows") rows = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--threads") threads = atoi(v.c_str());
    else if (k == "--positions") positions = atoi(v.c_str()) != 0;
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  if (linedoc.empty() || out.empty()) {
    fprintf(stderr, "usage: wsr_index_linedoc --linedoc FILE --out DIR [--rows N] [--threads T] [--positions 0|1]\n");
    return 2;
  }
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const auto t0 = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };

  int fd = open(linedoc.c_str(), O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "File may not exist: %s\n", linedoc.c_str()); return 1; }
  const size_t size = (size_t)st.st_size;
  const char *data = size ? (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
  if (size && data == MAP_FAILED) { perror("mmap"); return 1; }

  // line starts (std::getline semantics: a last line without '\n' still counts); line 0 is the header
  std::vector<size_t> line_at;
  {
    size_t p = 0;
    bool header = true;
    while (p < size && line_at.size() < rows) {
      const char *nl = (const char *)memchr(data + p, '\n', size - p);
      const size_t e = nl ? (size_t)(nl - data) : size;
      if (!header) line_at.push_back(p);
      header = false;
      p = e + 1;
    }
  }
  const size_t n_docs = line_at.size();
  if (n_docs >= (1ull << 31)) { fprintf(stderr, "too many documents\n"); return 1; }
  auto line_end = [&](size_t i) {
    const char *nl = (const char *)memchr(data + line_at[i], '\n', size - line_at[i]);
    return nl ? nl : data + size;
  };

  const double t_lines = since();
  // ---- pass 1: documents -> per-range term accumulators (parallel over contiguous doc ranges)
  const int n_parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads * 2, (n_docs + 255) / 256));
  std::vector<Part> parts(n_parts);
  for (int i = 0; i < n_parts; i++) {
    parts[i].doc_begin = n_docs * i / n_parts;
    parts[i].doc_end = n_docs * (i + 1) / n_parts;
  }
  {
    std::atomic<int> next{0};
    auto worker = [&]() {
      std::vector<uint32_t> tf_scratch;
      std::vector<std::vector<uint32_t>> pos_scratch;
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n_parts) return;
        Part &p = parts[i];
        for (size_t d = p.doc_begin; d < p.doc_end; d++)
          if (!ParseDoc(data + line_at[d], line_end(d), (uint32_t)d, positions, &p, &tf_scratch, &pos_scratch)) {
            if (p.err.empty()) p.err = "malformed row";
            p.err += " (document " + std::to_string(d) + ")";
            return;
          }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }
  for (const Part &p : parts)
    if (!p.err.empty()) { fprintf(stderr, "%s\n", p.err.c_str()); return 1; }

  const double t_parse = since();
  // ---- global term table, bytewise sorted
  std::vector<std::string_view> terms;
  {
    std::unordered_map<std::string_view, uint32_t> seen;
    for (const Part &p : parts)
      for (std::string_view t : p.names)
        if (seen.emplace(t, 0).second) terms.push_back(t);
    std::sort(terms.begin(), terms.end());
  }
  std::unordered_map<std::string_view, uint32_t> gid;
  gid.reserve(terms.size() * 2);
  for (size_t i = 0; i < terms.size(); i++) gid.emplace(terms[i], (uint32_t)i);
  // per part: local id -> global id (parallel lookups in the read-only table); per global term:
  // the parts that hold it, in doc order
  std::vector<std::vector<uint32_t>> to_global(n_parts);
  {
    std::atomic<int> next{0};
    auto worker = [&]() {
      for (;;) {
        const int pi = next.fetch_add(1);
        if (pi >= n_parts) return;
        const Part &p = parts[pi];
        std::vector<uint32_t> &g = to_global[pi];
        g.resize(p.names.size());
        for (size_t l = 0; l < p.names.size(); l++) g[l] = gid.find(p.names[l])->second;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> where(terms.size());   // (part, local id)
  for (int pi = 0; pi < n_parts; pi++)
    for (uint32_t l = 0; l < to_global[pi].size(); l++) where[to_global[pi][l]].push_back({(uint32_t)pi, l});
  const double t_merge = since();
  // ---- pass 2: term-major encoding in chunks of terms (parallel)
  std::vector<uint64_t> weight(terms.size() + 1, 0);
  for (size_t t = 0; t < terms.size(); t++) {
    uint64_t w = 16;
    for (auto &pl : where[t]) w += parts[pl.first].acc[pl.second].docs.size();
    weight[t + 1] = weight[t] + w;
  }
  const uint64_t per_chunk = std::max<uint64_t>(1 << 16, weight[terms.size()] / (uint64_t)(threads * 8) + 1);
  std::vector<Chunk> chunks;
  for (size_t t = 0; t < terms.size();) {
    size_t u = t + 1;
    while (u < terms.size() && weight[u] - weight[t] < per_chunk) u++;
    Chunk c;
    c.term_begin = t;
    c.term_end = u;
    chunks.push_back(std::move(c));
    t = u;
  }
  {
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      std::vector<uint32_t> docs, tfs, pos, delta, rd, rt;
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= chunks.size()) return;
        Chunk &c = chunks[i];
        for (uint64_t t = c.term_begin; t < c.term_end; t++) {
          docs.clear(); tfs.clear(); pos.clear();
          for (auto &pl : where[t]) {
            const TermAcc &a = parts[pl.first].acc[pl.second];
            docs.insert(docs.end(), a.docs.begin(), a.docs.end());
            tfs.insert(tfs.end(), a.tfs.begin(), a.tfs.end());
            pos.insert(pos.end(), a.pos.begin(), a.pos.end());
          }
          EncodeList(t, docs, tfs, pos, positions, &c, &delta, &rd, &rt);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }
  const double t_encode = since();
  std::vector<uint32_t> doc_len;
  doc_len.reserve(n_docs);
  for (const Part &p : parts) doc_len.insert(doc_len.end(), p.doc_len.begin(), p.doc_len.end());

  mkdir(out.c_str(), 0777);
  uint64_t file_size = 0, n_lists = 0, postings = 0;
  if (!WriteVacuumDir(out, chunks, [&](uint32_t t) { return std::string(terms[t]); }, doc_len, threads,
                      &file_size, &n_lists, &postings))
    return 1;
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"docs\": %zu, \"terms\": %llu, \"postings\": %llu, \"vacuum_bytes\": %llu, \"positions\": %s, "
         "\"seconds\": %.2f, \"threads\": %d, \"stage_seconds\": {\"lines\": %.2f, \"parse\": %.2f, "
         "\"term_table\": %.2f, \"encode\": %.2f, \"write\": %.2f}}\n",
         n_docs, (unsigned long long)n_lists, (unsigned long long)postings, (unsigned long long)file_size,
         positions ? "true" : "false", secs, threads, t_lines, t_parse - t_lines, t_merge - t_parse,
         t_encode - t_merge, secs - t_encode);
  return 0;
}
