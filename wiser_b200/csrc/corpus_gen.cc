// wsr_gen_corpus — native, multi-threaded synthetic Zipf corpus generator that writes a
// vacuum index directory in the REFERENCE's on-disk format, so that the very
// same directory is read by the unmodified reference engine (CPU baseline), the CPU oracle and
// the GPU loader. It replaces, for large synthetic corpora, the reference's single-threaded
// linedoc -> create_qq_mem_dump -> convert_qq_to_vacuum pipeline (the
// writer it mirrors is VacuumInvertedIndexDumper::DumpPostingListNoBloom,
// flash_engine_dumper.h:339-411, with GetCozyBoxWriter :78-104 and SkipListWriter
// flash_containers.h:236-308).
//
// Corpus model, generated TERM-major so no 1B-token shuffle is needed:
//   nominal doc length  L_d ~ clip(lognormal(mu, sigma), min, max)
//   term rank r has probability p_r = (r+1)^-s / H;  its token count is Poisson(p_r * T),
//   each token lands in a uniformly random token slot of the corpus (slot -> doc through the
//   cumulative lengths), i.e. the Poissonised multinomial bag-of-words model. tf = tokens of
//   the term in the doc; the stored doc length is the ACTUAL token count of the doc, and
//   avg length is accumulated as the reference does (doc_length_store.h:102-112).
// The doc-id and tf columns are always written; --positions 1 adds the position column (a
// token's position is its slot inside the document), which phrase queries need. The offset
// column is never written (skip rows carry zeros), so the directory serves queries without
// snippets — exactly the path under test. Deterministic for a given (seed, parameters), independent of thread count.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "vacuum_writer.h"

namespace {
using namespace wsrw;


struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t Next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  double Uni() { return ((Next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double Normal() {
    const double u = Uni(), v = Uni();
    return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
  }
  uint64_t Poisson(double lam) {
    if (lam < 30.0) {
      const double l = exp(-lam);
      uint64_t k = 0;
      double p = Uni();
      while (p > l) { k++; p *= Uni(); }
      return k;
    }
    const double x = lam + sqrt(lam) * Normal() + 0.5;
    return x < 0 ? 0 : (uint64_t)x;
  }
};

inline uint64_t Mix(uint64_t a, uint64_t b) {
  Rng r(a * 0x9e3779b97f4a7c15ull + b + 0x632be59bd9b4e019ull);
  r.Next();
  return r.Next();
}

struct Params {
  std::string out;
  uint64_t docs = 100000, vocab = 200000, seed = 1;
  double zipf = 1.0, mu = 4.3, sigma = 0.6;
  uint32_t min_len = 5, max_len = 2000;
  int threads = 0;
  bool positions = false;   // also write the position column (phrase queries)
};

struct Gen {
  Params P;
  std::vector<uint32_t> len_nominal;
  std::vector<uint64_t> cum;         // cum[d] = first token slot of doc d; cum[N] = T
  std::vector<uint32_t> coarse;      // coarse[k] = doc holding slot k << shift
  int shift = 0;
  uint64_t T = 0;
  double H = 0;
  std::vector<std::atomic<uint32_t>> actual_len;

  uint32_t DocOfSlot(uint64_t slot) const {
    uint32_t d = coarse[slot >> shift];
    while (cum[d + 1] <= slot) d++;
    return d;
  }

  void BuildDocs() {
    const uint64_t N = P.docs;
    len_nominal.resize(N);
    cum.resize(N + 1);
    for (uint64_t d = 0; d < N; d++) {
      Rng r(Mix(P.seed, 0xD0C00000000ull + d));
      double l = exp(P.mu + P.sigma * r.Normal());
      l = std::min<double>(std::max<double>(l, P.min_len), P.max_len);
      len_nominal[d] = (uint32_t)l;
    }
    cum[0] = 0;
    for (uint64_t d = 0; d < N; d++) cum[d + 1] = cum[d] + len_nominal[d];
    T = cum[N];
    shift = 0;
    while ((T >> shift) > 4 * N + 16) shift++;
    coarse.resize((T >> shift) + 2);
    uint32_t d = 0;
    for (uint64_t k = 0; k < coarse.size(); k++) {
      const uint64_t slot = k << shift;
      while (d + 1 < N && cum[d + 1] <= slot) d++;
      coarse[k] = d;
    }
    H = 0;
    for (uint64_t r = 0; r < P.vocab; r++) H += pow((double)(r + 1), -P.zipf);
    actual_len = std::vector<std::atomic<uint32_t>>(N);
    for (auto &a : actual_len) a.store(0, std::memory_order_relaxed);
  }

  // Postings (doc ascending, tf) of one term; with P.positions also the in-document positions
  // of every posting, ascending and distinct, postings back to back.
  void TermPostings(uint64_t r, std::vector<uint32_t> *docs, std::vector<uint32_t> *tfs,
                    std::vector<uint64_t> *scratch, std::vector<uint32_t> *pos) {
    docs->clear();
    tfs->clear();
    pos->clear();
    const double p = pow((double)(r + 1), -P.zipf) / H;
    const double expect = p * (double)T;
    Rng rng(Mix(P.seed, 0x7E4A00000000ull + r));
    const uint64_t N = P.docs;
    if (expect >= 0.5 * (double)N) {
      for (uint64_t d = 0; d < N; d++) {
        uint64_t tf = rng.Poisson(p * len_nominal[d]);
        if (!tf) continue;
        if (P.positions) {
          // tf distinct positions inside the document (rejection; tf is far below the length)
          const uint32_t L = len_nominal[d];
          tf = std::min<uint64_t>(tf, L);
          const size_t at = pos->size();
          while (pos->size() - at < tf) {
            const uint32_t x = (uint32_t)(((unsigned __int128)rng.Next() * L) >> 64);
            bool dup = false;
            for (size_t q = at; q < pos->size(); q++) dup = dup || (*pos)[q] == x;
            if (!dup) pos->push_back(x);
          }
          std::sort(pos->begin() + at, pos->end());
        }
        docs->push_back((uint32_t)d);
        tfs->push_back((uint32_t)tf);
      }
      return;
    }
    const uint64_t n = rng.Poisson(expect);
    if (!n) return;
    scratch->resize(n);
    for (uint64_t i = 0; i < n; i++)   // uniform slot in [0, T): 64x64 -> high 64 multiply
      (*scratch)[i] = (uint64_t)(((unsigned __int128)rng.Next() * T) >> 64);
    std::sort(scratch->begin(), scratch->end());
    uint64_t prev_slot = ~0ull;
    uint32_t d = 0;
    bool first = true;
    for (uint64_t i = 0; i < n; i++) {
      const uint64_t slot = (*scratch)[i];
      if (P.positions && slot == prev_slot) continue;   // one token per slot: positions stay distinct
      prev_slot = slot;
      const uint32_t doc = (!first && slot < cum[d + 1]) ? d : DocOfSlot(slot);
      if (first || doc != d) {
        docs->push_back(doc);
        tfs->push_back(1);
        d = doc;
        first = false;
      } else {
        tfs->back()++;
      }
      if (P.positions) pos->push_back((uint32_t)(slot - cum[doc]));
    }
  }

  void RunChunk(Chunk *c) {
    std::vector<uint32_t> docs, tfs, pos, delta, rd, rt;
    std::vector<uint64_t> scratch;
    for (uint64_t r = c->term_begin; r < c->term_end; r++) {
      TermPostings(r, &docs, &tfs, &scratch, &pos);
      if (docs.empty()) continue;
      for (size_t i = 0; i < docs.size(); i++)
        actual_len[docs[i]].fetch_add(tfs[i], std::memory_order_relaxed);
      EncodeList(r, docs, tfs, pos, P.positions, c, &delta, &rd, &rt);
    }
  }
};

}  // namespace

int main(int argc, char **argv) {
  Params P;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--out") P.out = v;
    else if (k == "--docs") P.docs = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--vocab") P.vocab = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--seed") P.seed = strtoull(v.c_str(), nullptr, 10);
    else if (k == "--zipf") P.zipf = atof(v.c_str());
    else if (k == "--mu") P.mu = atof(v.c_str());
    else if (k == "--sigma") P.sigma = atof(v.c_str());
    else if (k == "--min-len") P.min_len = (uint32_t)atoi(v.c_str());
    else if (k == "--max-len") P.max_len = (uint32_t)atoi(v.c_str());
    else if (k == "--threads") P.threads = atoi(v.c_str());
    else if (k == "--positions") P.positions = atoi(v.c_str()) != 0;
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  if (P.out.empty() || P.docs == 0 || P.vocab == 0 || P.docs >= (1ull << 31)) {
    fprintf(stderr, "usage: wsr_gen_corpus --out DIR --docs N --vocab V [--zipf s] [--mu m] "
                    "[--sigma s] [--min-len a] [--max-len b] [--seed x] [--threads t] [--positions 0|1]\n");
    return 2;
  }
  if (P.max_len >= (1u << 19)) { fprintf(stderr, "max-len must stay below 2^19 (SURVEY §5.1)\n"); return 2; }
  if (P.threads <= 0) P.threads = (int)std::max(1u, std::thread::hardware_concurrency());
  mkdir(P.out.c_str(), 0777);
  const auto t0 = std::chrono::steady_clock::now();

  Gen g;
  g.P = P;
  g.BuildDocs();

  // chunks of ~equal expected token count, in term order
  std::vector<Chunk> chunks;
  {
    const double per_chunk = std::max(2.0e5, (double)g.T / (64.0 * P.threads));
    double acc = 0;
    uint64_t begin = 0;
    for (uint64_t r = 0; r < P.vocab; r++) {
      acc += pow((double)(r + 1), -P.zipf) / g.H * (double)g.T + 2.0;
      if (acc >= per_chunk || r + 1 == P.vocab) {
        Chunk c;
        c.term_begin = begin;
        c.term_end = r + 1;
        chunks.push_back(std::move(c));
        begin = r + 1;
        acc = 0;
      }
    }
  }
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= chunks.size()) return;
      g.RunChunk(&chunks[i]);
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < P.threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }

  std::vector<uint32_t> doc_len(P.docs);
  for (uint64_t d = 0; d < P.docs; d++) doc_len[d] = g.actual_len[d].load(std::memory_order_relaxed);
  uint64_t file_size = 0, n_lists = 0, postings = 0;
  if (!wsrw::WriteVacuumDir(P.out, chunks, [](uint32_t t) { return "t" + std::to_string(t); }, doc_len,
                            P.threads, &file_size, &n_lists, &postings))
    return 1;
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"docs\": %llu, \"vocab\": %llu, \"terms_with_postings\": %llu, \"postings\": %llu, "
         "\"tokens_nominal\": %llu, \"vacuum_bytes\": %llu, \"seconds\": %.2f, \"threads\": %d}\n",
         (unsigned long long)P.docs, (unsigned long long)P.vocab, (unsigned long long)n_lists,
         (unsigned long long)postings, (unsigned long long)g.T, (unsigned long long)file_size, secs,
         P.threads);
  return 0;
}
