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

namespace {

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

inline int VarLen(uint64_t v) { int n = 1; while (v >= 128) { v >>= 7; n++; } return n; }
inline void PutVarint(std::vector<uint8_t> *b, uint64_t v) {   // utils::varint_encode, utils.cc:257-270
  while (v >= 128) { b->push_back((uint8_t)(v & 0x7f) | 0x80); v >>= 7; }
  b->push_back((uint8_t)v);
}
// Fixed-width 7-byte varint (non-canonical, zero high groups): decodes to v with the reference's
// varint_decode_64bit (utils.h:249-266); lets the first skip row be patched in place.
inline void PutVarint7(uint8_t *p, uint64_t v) {
  for (int i = 0; i < 6; i++) { p[i] = (uint8_t)(v & 0x7f) | 0x80; v >>= 7; }
  p[6] = (uint8_t)(v & 0x7f);
}
inline int BitWidth(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

// One column (doc-id deltas or tfs) as a "cozy box": 128-value packs then a VInts tail
// (GeneralTermEntry::GetCozyBoxWriter, flash_engine_dumper.h:78-104). Records the offset
// (relative to `base`) of the blob that holds posting 128*r for every skip row r.
void EncodeColumn(const uint32_t *v, size_t n, std::vector<uint8_t> *buf, size_t base,
                  std::vector<uint32_t> *row_off) {
  const size_t n_packs = n / 128;
  for (size_t p = 0; p < n_packs; p++) {
    row_off->push_back((uint32_t)(buf->size() - base));
    const uint32_t *x = v + p * 128;
    uint32_t m = 0;
    for (int i = 0; i < 128; i++) m |= x[i];
    const int bits = std::max(1, BitWidth(m));                  // LittlePackedIntsWriter::Add
    buf->push_back(0xD6);
    buf->push_back((uint8_t)bits);
    const size_t at = buf->size();
    buf->resize(at + 16 * (size_t)bits, 0);
    uint8_t *d = buf->data() + at;
    uint64_t acc = 0;
    int have = 0;
    size_t w = 0;
    for (int i = 0; i < 128; i++) {                             // LSB-first bitstream
      acc |= (uint64_t)x[i] << have;
      have += bits;
      while (have >= 8) { d[w++] = (uint8_t)acc; acc >>= 8; have -= 8; }
    }
  }
  if (n % 128) {
    row_off->push_back((uint32_t)(buf->size() - base));
    size_t nbytes = 0;
    for (size_t i = n_packs * 128; i < n; i++) nbytes += VarLen(v[i]);
    buf->push_back(0x9B);                                       // VIntsWriter::Serialize
    PutVarint(buf, nbytes);
    for (size_t i = n_packs * 128; i < n; i++) PutVarint(buf, v[i]);
  }
}

struct ListRec {        // what is needed to patch the first skip row once offsets are absolute
  uint32_t term;
  uint32_t df;
  uint64_t rel_start;   // list start inside the chunk buffer
  uint32_t patch_at;    // offset (from list start) of the 7-byte absolute fields
  uint32_t docid0, tf0; // offsets (from list start) of the first doc-id / tf blob
  uint32_t pos0;        // offset of the first position blob (0: no position column)
  uint32_t tf_end;      // offset (from list start) of the end of the tf column
};

struct Chunk {
  uint64_t term_begin = 0, term_end = 0;
  std::vector<uint8_t> buf;
  std::vector<ListRec> lists;
  uint64_t abs_start = 0;
  uint64_t postings = 0;
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

  void EncodeList(uint64_t r, const std::vector<uint32_t> &docs, const std::vector<uint32_t> &tfs,
                  const std::vector<uint32_t> &pos, Chunk *c, std::vector<uint32_t> *delta,
                  std::vector<uint32_t> *rows_d, std::vector<uint32_t> *rows_t) {
    const size_t df = docs.size();
    const size_t n_rows = (df + 127) / 128;
    std::vector<uint8_t> &b = c->buf;
    ListRec rec;
    rec.term = (uint32_t)r;
    rec.df = (uint32_t)df;
    rec.rel_start = b.size();
    const size_t start = b.size();
    b.push_back(0xF4);                       // POSTING_LIST_FIRST_BYTE
    PutVarint(&b, df);
    b.insert(b.end(), 8, 0);                 // Bloom skip-list offsets: none
    // doc-id deltas over the whole list from 0 (utils::EncodeDelta, utils.h:573-584)
    delta->resize(df);
    uint32_t prev = 0;
    for (size_t i = 0; i < df; i++) { (*delta)[i] = docs[i] - prev; prev = docs[i]; }
    // The skip list sits in front of the data; data offsets are needed first -> encode the
    // columns into a side buffer, then emit skip list + columns.
    std::vector<uint8_t> cols;
    rows_d->clear();
    rows_t->clear();
    EncodeColumn(delta->data(), df, &cols, 0, rows_d);
    const size_t tf_col = cols.size();
    EncodeColumn(tfs.data(), df, &cols, 0, rows_t);
    const size_t pos_col = cols.size();
    // position column: per bag deltas (first from 0), all bags concatenated, same cozy-box
    // container; skip row r addresses the first position of posting 128*r as (blob, index)
    std::vector<uint32_t> blobs_p, prow_off, prow_idx;
    if (P.positions) {
      std::vector<uint32_t> pd(pos.size());
      size_t at = 0;
      for (size_t i = 0; i < df; i++) {
        if (i % 128 == 0) { prow_off.push_back((uint32_t)(at / 128)); prow_idx.push_back((uint32_t)(at % 128)); }
        uint32_t prev_p = 0;
        for (uint32_t j = 0; j < tfs[i]; j++, at++) { pd[at] = pos[at] - prev_p; prev_p = pos[at]; }
      }
      EncodeColumn(pd.data(), pd.size(), &cols, 0, &blobs_p);
      for (auto &o : prow_off) o = blobs_p[o];     // blob index -> offset inside cols
    }
    // skip list: 0xA3, n_rows, rows of 7 varints (flash_containers.h:250-299). Row 0's two
    // absolute blob offsets use fixed 7-byte varints patched after layout; later rows are
    // deltas vs the previous row and do not depend on the absolute position.
    const size_t skip_at = b.size();
    b.push_back(0xA3);
    PutVarint(&b, n_rows);
    size_t data_at = 0;   // offset of the columns from list start, known once the skip list is sized
    {
      size_t sz = b.size() - start;
      sz += 1 + 7 + 7 + (P.positions ? 7 : 1) + 3;   // row 0
      for (size_t rr = 1; rr < n_rows; rr++) {
        sz += VarLen(docs[rr * 128 - 1] - (rr >= 2 ? docs[(rr - 1) * 128 - 1] : 0));
        sz += VarLen((*rows_d)[rr] - (*rows_d)[rr - 1]);
        sz += VarLen((*rows_t)[rr] - (*rows_t)[rr - 1]);
        if (P.positions) sz += VarLen(prow_off[rr] - prow_off[rr - 1]) + VarLen(prow_idx[rr]) + 2;
        else sz += 4;
      }
      data_at = sz;
    }
    (void)skip_at;
    b.push_back(0);                          // row 0: previous_doc_id = 0
    rec.patch_at = (uint32_t)(b.size() - start);
    b.insert(b.end(), 14, 0);                // docid blob abs offset, tf blob abs offset (patched)
    b.insert(b.end(), P.positions ? 7 : 1, 0);   // pos blob abs offset (patched) or 0 = no positions
    b.insert(b.end(), 3, 0);                 // pos idx (0), offset blob off, offset idx
    for (size_t rr = 1; rr < n_rows; rr++) {
      PutVarint(&b, docs[rr * 128 - 1] - (rr >= 2 ? docs[(rr - 1) * 128 - 1] : 0));
      PutVarint(&b, (*rows_d)[rr] - (*rows_d)[rr - 1]);
      PutVarint(&b, (*rows_t)[rr] - (*rows_t)[rr - 1]);
      if (P.positions) {
        PutVarint(&b, prow_off[rr] - prow_off[rr - 1]);
        PutVarint(&b, prow_idx[rr]);
        b.insert(b.end(), 2, 0);
      } else {
        b.insert(b.end(), 4, 0);
      }
    }
    if (b.size() - start != data_at) { fprintf(stderr, "internal: skip list size\n"); abort(); }
    rec.docid0 = (uint32_t)data_at;
    rec.tf0 = (uint32_t)(data_at + tf_col);
    rec.pos0 = P.positions ? (uint32_t)(data_at + pos_col) : 0u;
    rec.tf_end = (uint32_t)(data_at + pos_col);
    b.insert(b.end(), cols.begin(), cols.end());
    c->lists.push_back(rec);
    c->postings += df;
  }

  void RunChunk(Chunk *c) {
    std::vector<uint32_t> docs, tfs, pos, delta, rd, rt;
    std::vector<uint64_t> scratch;
    for (uint64_t r = c->term_begin; r < c->term_end; r++) {
      TermPostings(r, &docs, &tfs, &scratch, &pos);
      if (docs.empty()) continue;
      for (size_t i = 0; i < docs.size(); i++)
        actual_len[docs[i]].fetch_add(tfs[i], std::memory_order_relaxed);
      EncodeList(r, docs, tfs, pos, c, &delta, &rd, &rt);
    }
  }
};

bool WriteFile(const std::string &path, const void *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = n == 0 || fwrite(p, 1, n, f) == n;
  fclose(f);
  return ok;
}

// utils::UintToChar4, utils.h:301-315
uint8_t UintToChar4(uint32_t val) {
  if (val < 8) return (uint8_t)val;
  const int nb = BitWidth(val), sh = nb - 4;
  return (uint8_t)(((val >> sh) & 7u) | ((uint32_t)(sh + 1) << 3));
}

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

  // absolute layout: 100-byte header (VacuumInvertedIndexDumper::DumpHeader,
  // flash_engine_dumper.h:288-316: 0x88, two Bloom descriptors of {0,0,0,f32 0}, pad to 100)
  uint64_t off = 100, postings = 0, n_lists = 0;
  for (auto &c : chunks) { c.abs_start = off; off += c.buf.size(); postings += c.postings; n_lists += c.lists.size(); }
  const uint64_t file_size = off;

  // patch first skip rows, build my.tip (term_len, term, (pages << 48) | offset) and terms.txt
  std::vector<uint8_t> tip;
  std::string terms_txt;
  tip.reserve(n_lists * 24);
  for (auto &c : chunks) {
    for (const ListRec &l : c.lists) {
      const uint64_t abs = c.abs_start + l.rel_start;
      uint8_t *p = c.buf.data() + l.rel_start + l.patch_at;
      PutVarint7(p, abs + l.docid0);
      PutVarint7(p + 7, abs + l.tf0);
      if (l.pos0) PutVarint7(p + 14, abs + l.pos0);
      char name[32];
      const int len = snprintf(name, sizeof(name), "t%u", l.term);
      const uint32_t ulen = (uint32_t)len;
      const uint64_t pages = std::min<uint64_t>(l.tf_end / 4096, 0xffff);
      const uint64_t val = (pages << 48) | abs;
      tip.insert(tip.end(), (const uint8_t *)&ulen, (const uint8_t *)&ulen + 4);
      tip.insert(tip.end(), name, name + len);
      tip.insert(tip.end(), (const uint8_t *)&val, (const uint8_t *)&val + 8);
      terms_txt += name;
      terms_txt += ' ';
      terms_txt += std::to_string(l.df);
      terms_txt += '\n';
    }
  }

  // my.vacuum
  {
    const std::string path = P.out + "/my.vacuum";
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0 || ftruncate(fd, (off_t)file_size) != 0) { perror("my.vacuum"); return 1; }
    uint8_t header[100];
    memset(header, 0, sizeof(header));
    header[0] = 0x88;
    if (pwrite(fd, header, 100, 0) != 100) { perror("pwrite"); return 1; }
    next = 0;
    std::atomic<bool> ok{true};
    auto writer = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= chunks.size()) return;
        const Chunk &c = chunks[i];
        size_t done = 0;
        while (done < c.buf.size()) {
          ssize_t w = pwrite(fd, c.buf.data() + done, c.buf.size() - done, (off_t)(c.abs_start + done));
          if (w <= 0) { ok = false; return; }
          done += (size_t)w;
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::min(P.threads, 8); t++) pool.emplace_back(writer);
    writer();
    for (auto &t : pool) t.join();
    close(fd);
    if (!ok) { fprintf(stderr, "write error on my.vacuum\n"); return 1; }
  }
  if (!WriteFile(P.out + "/my.tip", tip.data(), tip.size()) ||
      !WriteFile(P.out + "/terms.txt", terms_txt.data(), terms_txt.size())) {
    perror("my.tip");
    return 1;
  }

  // my.doc_length: i32 count, f64 avg (running mean in doc order), count x {i32 id, i8 char4}
  {
    const uint64_t N = P.docs;
    std::vector<uint8_t> dl(12 + 5 * N);
    double avg = 0;
    for (uint64_t d = 0; d < N; d++) {
      const int len = (int)g.actual_len[d].load(std::memory_order_relaxed);
      avg = avg + (len - avg) / (double)(d + 1);           // DocLengthCharStore::AddLength
      const int32_t id = (int32_t)d;
      memcpy(&dl[12 + 5 * d], &id, 4);
      dl[12 + 5 * d + 4] = UintToChar4((uint32_t)len);
    }
    const int32_t count = (int32_t)N;
    memcpy(&dl[0], &count, 4);
    memcpy(&dl[4], &avg, 8);
    if (!WriteFile(P.out + "/my.doc_length", dl.data(), dl.size())) { perror("my.doc_length"); return 1; }
  }
  // stub doc store: 0 documents (ChunkedDocStoreReader::LoadFdx, doc_store.h:365-392)
  {
    const uint8_t fdx[4] = {0x00, 0x80, 0x80, 0x01};
    const uint8_t fdt[1] = {0};
    if (!WriteFile(P.out + "/my.fdx", fdx, 4) || !WriteFile(P.out + "/my.fdt", fdt, 1)) return 1;
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"docs\": %llu, \"vocab\": %llu, \"terms_with_postings\": %llu, \"postings\": %llu, "
         "\"tokens_nominal\": %llu, \"vacuum_bytes\": %llu, \"seconds\": %.2f, \"threads\": %d}\n",
         (unsigned long long)P.docs, (unsigned long long)P.vocab, (unsigned long long)n_lists,
         (unsigned long long)postings, (unsigned long long)g.T, (unsigned long long)file_size, secs,
         P.threads);
  return 0;
}
