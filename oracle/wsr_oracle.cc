// TEST INFRASTRUCTURE ONLY — see wsr_oracle.h. CPU restatement of the reference's
// conjunctive-query hot path, written to follow the reference's control flow step by step so
// that iteration order, floating-point operation order and heap tie behaviour are the same.
// Every function cites the reference code (paths relative to src/qq_mem/src/) it restates.
// Compiled with -ffp-contract=off and no -march (the reference has none: CMakeLists.txt:6,12),
// so every * + / rounds separately.
#include "wsr_oracle.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

constexpr int kPack = 128;             // PACK_ITEM_CNT, packed_value.h:13
constexpr uint8_t kVacuumMagic = 0x88; // types.h:43-51
constexpr uint8_t kListMagic = 0xF4;
constexpr uint8_t kSkipMagic = 0xA3;
constexpr uint8_t kPackMagic = 0xD6;
constexpr uint8_t kVIntsMagic = 0x9B;

struct Mapped {
  const uint8_t *p = nullptr;
  size_t n = 0;
  bool Open(const std::string &path) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    n = st.st_size;
    if (n) {
      void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { close(fd); return false; }
      p = (const uint8_t *)m;
    }
    close(fd);
    return true;
  }
  ~Mapped() { if (p) munmap((void *)p, n); }
};

// utils::varint_decode_uint32 / varint_decode_64bit, utils.h:230-266 (LEB128, low group first).
inline int VarintDecode(const uint8_t *b, uint64_t *v) {
  uint64_t r = 0;
  int i = 0;
  for (;;) {
    r += (uint64_t)(b[i] & 0x7f) << (7 * i);
    if (!(b[i++] & 0x80)) break;
  }
  *v = r;
  return i;
}

// utils::Char4ToUint, utils.h:317-329.
inline uint32_t Char4ToUint(int c) {
  uint32_t bits = c & 0x07;
  int shift = ((c & 0xff) >> 3) - 1;
  return shift == -1 ? bits : (bits | 0x08) << shift;
}

// turbounpack32 for 128 values, LittleIntPacker/src/turbobitpacking32.c:3863-3868: the pack is
// one LSB-first little-endian bitstream, value i at bit offset i*bits (packed_value.h:205-209).
inline void Unpack128(const uint8_t *in, int bits, uint32_t *out) {
  if (bits == 0) { memset(out, 0, kPack * 4); return; }
  const uint64_t mask = bits == 32 ? 0xffffffffull : ((1ull << bits) - 1);
  for (int i = 0; i < kPack; i++) {
    size_t bit = (size_t)i * bits;
    uint64_t w = 0;
    size_t byte = bit >> 3;
    size_t avail = (size_t)16 * bits - byte;
    memcpy(&w, in + byte, avail < 8 ? avail : 8);
    out[i] = (uint32_t)((w >> (bit & 7)) & mask);
  }
}

// One row of the skip list, flash_containers.h:312-335 (only the two columns this path reads).
struct SkipRow {
  uint32_t prev_doc;
  uint64_t docid_off;
  uint64_t tf_off;
  uint64_t pos_off;   // blob holding the first position of posting 128*r ...
  uint32_t pos_idx;   // ... and its index inside that blob
};

// A decoded blob of <=128 values: either a pack (0xD6,bits,16*bits B) or the VInts tail
// (0x9B, varint n_bytes, varints) — packed_value.h:184-235, 400-460.
struct Blob {
  uint32_t v[kPack];
  int n = 0;       // values present (128 for a pack)
  bool vints = false;
  void Load(const uint8_t *b) {
    if (b[0] == kPackMagic) {
      Unpack128(b + 2, b[1], v);
      n = kPack;
      vints = false;
    } else {
      // VIntsIterator::Reset, packed_value.h:408-424
      uint64_t nbytes;
      int len = VarintDecode(b + 1, &nbytes);
      const uint8_t *p = b + 1 + len, *e = p + nbytes;
      n = 0;
      while (p < e && n < kPack) {
        uint64_t x;
        p += VarintDecode(p, &x);
        v[n++] = (uint32_t)x;
      }
      vints = true;
    }
  }
};

struct Index;

// Restates VacuumPostingListIterator (flash_iterators.h:893-1079) restricted to the doc-id and
// term-frequency columns: DocIdIterator (:121-262) + TermFreqIterator (:43-118).
class ListCursor {
 public:
  // ResetWithZoneInfo, flash_iterators.h:903-956: magic, varint df, 8 reserved bytes, skip list.
  bool Reset(const uint8_t *file, uint64_t list_off) {
    file_ = file;
    const uint8_t *b = file + list_off;
    if (b[0] != kListMagic) return false;
    b += 1;
    uint64_t df;
    b += VarintDecode(b, &df);
    n_postings_ = (int64_t)df;
    b += 8;
    // SkipList::Load, flash_containers.h:354-391: 7 delta-coded varints per row.
    if (b[0] != kSkipMagic) return false;
    uint64_t n_rows;
    b += 1;
    b += VarintDecode(b, &n_rows);
    rows_.resize(n_rows);
    uint64_t pd = 0, po = 0, pt = 0, pp = 0, x, pidx;
    for (uint64_t r = 0; r < n_rows; r++) {
      b += VarintDecode(b, &x); pd += x;
      b += VarintDecode(b, &x); po += x;
      b += VarintDecode(b, &x); pt += x;
      b += VarintDecode(b, &x); pp += x;   // position blob offset (delta vs previous row)
      b += VarintDecode(b, &pidx);         // position in-blob index
      b += VarintDecode(b, &x);            // offset blob off (delta) — snippets only
      b += VarintDecode(b, &x);            // offset in-blob index
      rows_[r] = {(uint32_t)pd, po, pt, pp, (uint32_t)pidx};
    }
    // DocIdIterator::Reset -> SkipTo(0), flash_iterators.h:131-139
    cur_ = 0;
    doc_blob_ = -1;
    tf_blob_ = -1;
    if (n_postings_ > 0) SkipToIndex(0);
    return true;
  }

  int64_t Size() const { return n_postings_; }
  bool IsEnd() const { return cur_ == n_postings_; }       // flash_iterators.h:177-179
  uint32_t DocId() const { return value_; }                // Value(), :205-214
  int64_t PostingIndex() const { return cur_; }

  void Advance() { SkipToIndex(cur_ + 1); }                // :201-203

  // DocIdIterator::SkipForward, flash_iterators.h:181-199, with GetBlobIndexToGo (:218-227):
  // walk skip rows linearly while the NEXT row's previous_doc_id < val, then scan inside the
  // blob (DeltaEncodedPackedIntsIterator::SkipForward, packed_value.h:346-350 /
  // DeltaEncodedVIntsIterator::SkipForward, :483-487).
  void SkipForward(uint32_t val) {
    int64_t blob = cur_ / kPack;
    const int64_t last = (n_postings_ - 1) / kPack;
    int64_t go = blob;
    while (go + 1 <= last && rows_[go + 1].prev_doc < val) go++;
    if (go != blob) SkipToIndex(go * kPack);
    // in-blob scan; in_ is the index inside doc_ (the current blob)
    while (in_ < doc_.n && value_ < val) StepInBlob();
    cur_ = go * kPack + in_;
    // A pack that runs out of values lands on index 128 == start of the next blob's range;
    // the reference leaves the iterator there with IsEnd() decided by cur == df
    // (flash_iterators.h:193-198). Value() is then undefined in the reference; for a
    // non-final blob this cannot happen because the chosen blob's last doc is >= val.
  }

  // VacuumPostingListIterator::TermFreq, flash_iterators.h:989-992: random access by posting
  // index with a one-blob cache.
  uint32_t TermFreq() { return TfAt(cur_); }
  uint32_t TfAt(int64_t idx) {
    int64_t blob = idx / kPack;
    if (blob != tf_blob_) {
      tf_.Load(file_ + rows_[blob].tf_off);
      tf_blob_ = blob;
    }
    return tf_.v[idx % kPack];
  }

  // Positions of the CURRENT posting: AssignPositionBegin + InBagPositionIterator::Pop until
  // IsEnd (flash_iterators.h:1002-1005, 558-661). The position column is one "cozy box": the
  // bags' values (deltas inside each bag, first from 0) concatenated and cut into 128-value
  // packs plus a VInts tail, blobs back to back (CozyBoxIterator, :280-425). Skip row r points at
  // the first value of posting 128*r; reaching posting p from there means advancing by the tfs
  // of postings 128*r .. p-1 (PositionPostingBagIterator::NumCozyEntriesBetween, :619-628).
  void Positions(std::vector<uint32_t> *out) {
    out->clear();
    const int64_t r = cur_ / kPack;
    uint64_t blob_off = rows_[r].pos_off;
    uint64_t in_blob = rows_[r].pos_idx;
    for (int64_t q = r * kPack; q < cur_; q++) in_blob += TfAt(q);
    const uint32_t tf = TfAt(cur_);
    // CozyBoxIterator::AdvanceBy, flash_iterators.h:317-336: whole packs are stepped over by
    // their serialized size (2 + 16*bits); the VInts blob is the last one of the column
    while (file_[blob_off] == kPackMagic && in_blob >= (uint64_t)kPack) {
      in_blob -= kPack;
      blob_off += 2 + 16ull * file_[blob_off + 1];
    }
    Blob blob;
    blob.Load(file_ + blob_off);
    uint32_t prev = 0;
    for (uint32_t i = 0; i < tf; i++) {
      if (!blob.vints && in_blob == (uint64_t)kPack) {      // CozyBoxIterator::Advance, :309-315
        blob_off += 2 + 16ull * file_[blob_off + 1];
        blob.Load(file_ + blob_off);
        in_blob = 0;
      }
      prev += blob.v[in_blob++];
      out->push_back(prev);
    }
  }

 private:
  // DocIdIterator::SkipTo(posting_index), flash_iterators.h:141-166
  void SkipToIndex(int64_t idx) {
    int64_t blob = idx / kPack;
    if (doc_blob_ < 0 || blob != cur_ / kPack || blob != doc_blob_) {
      if (idx >= n_postings_) { cur_ = n_postings_; return; }   // SkipToEnd, :150-152
      // SetupBlob, :233-246
      doc_.Load(file_ + rows_[blob].docid_off);
      doc_blob_ = blob;
      in_ = 0;
      prev_ = rows_[blob].prev_doc;
      value_ = prev_ + doc_.v[0];
    }
    int target = (int)(idx % kPack);
    while (in_ < target) StepInBlob();
    cur_ = idx;
  }
  // DeltaEncodedPackedIntsIterator::Advance, packed_value.h:335-338
  void StepInBlob() {
    prev_ = value_;
    in_++;
    if (in_ < doc_.n) value_ = prev_ + doc_.v[in_];
  }

  const uint8_t *file_ = nullptr;
  std::vector<SkipRow> rows_;
  int64_t n_postings_ = 0;
  int64_t cur_ = 0;
  Blob doc_, tf_;
  int64_t doc_blob_ = -1, tf_blob_ = -1;
  int in_ = 0;
  uint32_t prev_ = 0, value_ = 0;
};

struct Index {
  Mapped vacuum, tip;
  std::unordered_map<std::string, uint64_t> term_to_off;  // TermTrieIndex, term_index.h:100-159
  std::vector<std::string> term_order;
  std::vector<signed char> norms;  // DocLengthCharStore::vec_char_store_
  int doc_count = 0;
  double avg_len = 0;
  double cache[256];               // Bm25Similarity::cache_
  // Partition mode (document-partitioned deployments): a partition directory scored
  // with the COLLECTION's statistics — global N, average length and per-term df — so that its
  // top-k lists merge into what one big reference index would return.
  std::unordered_map<uint64_t, int32_t> global_df;   // list offset -> collection-wide df

  void BuildCache() {
    // Bm25Similarity::BuildCache, scoring.h:85-90: k1*(1 - b + b*len/avg), left to right.
    const double k1 = 1.2, b = 0.75;
    for (int i = 0; i < 256; i++) {
      uint32_t field_length = Char4ToUint(i & 0xff);
      cache[i] = k1 * (1 - b + b * field_length / avg_len);
    }
  }

  bool Load(const std::string &dir, std::string *err) {
    if (!vacuum.Open(dir + "/my.vacuum") || vacuum.n < 100 || vacuum.p[0] != kVacuumMagic) {
      *err = "cannot open my.vacuum or bad magic";
      return false;
    }
    if (!tip.Open(dir + "/my.tip")) { *err = "cannot open my.tip"; return false; }
    // TermTrieIndex::Load/LoadEntry, term_index.h:106-159; value decode flash_containers.h:14-19
    const uint8_t *p = tip.p, *e = tip.p + tip.n;
    while (p < e) {
      uint32_t len;
      memcpy(&len, p, 4);
      std::string term((const char *)p + 4, len);
      int64_t v;
      memcpy(&v, p + 4 + len, 8);
      p += 4 + len + 8;
      term_to_off[term] = (uint64_t)v & 0xffffffffffffull;
      term_order.push_back(term);
    }
    // DocLengthCharStore::Deserialize, doc_length_store.h:164-190
    Mapped dl;
    if (!dl.Open(dir + "/my.doc_length") || dl.n < 12) { *err = "cannot open my.doc_length"; return false; }
    int32_t count;
    memcpy(&count, dl.p, 4);
    memcpy(&avg_len, dl.p + 4, 8);
    const uint8_t *q = dl.p + 12;
    doc_count = 0;
    for (int i = 0; i < count; i++) {
      int32_t id;
      memcpy(&id, q, 4);
      signed char c = (signed char)q[4];
      q += 5;
      if ((size_t)id >= norms.size()) norms.resize(id + 1, 0);
      norms[id] = c;
      doc_count++;
    }
    BuildCache();
    return true;
  }
};

// calc_es_idf, scoring.h:21-25
inline double EsIdf(int doc_count, int doc_freq) {
  return log(1 + (doc_count - doc_freq + 0.5) / (doc_freq + 0.5));
}

struct HeapEntry { int32_t doc; double score; };
// EntryGreater, query_processing.h:510-517 — min-heap on score only.
struct EntryGreater {
  bool operator()(const HeapEntry &a, const HeapEntry &b) const { return a.score > b.score; }
};
using MinHeap = std::priority_queue<HeapEntry, std::vector<HeapEntry>, EntryGreater>;

struct Processor {
  const Index &ix;
  std::vector<ListCursor> &its;
  int k;
  std::vector<double> idfs;
  MinHeap heap;

  // ProcessorBase ctor, query_processing.h:530-548
  Processor(const Index &ix_, std::vector<ListCursor> &its_, int k_, const std::vector<int32_t> &dfs)
      : ix(ix_), its(its_), k(k_) {
    for (size_t i = 0; i < its.size(); i++) idfs.push_back(EsIdf(ix.doc_count, dfs[i]));
  }

  // CalcDocScoreLossy, scoring.h:124-145, with TfNormLossy (:65-69). The reference indexes
  // cache_ with a signed char; norm bytes >= 128 are UB there and not supported here either.
  double Score(int32_t doc) {
    const signed char norm = ix.norms[doc];
    double final_doc_score = 0;
    for (size_t i = 0; i < its.size(); i++) {
      const int freq = (int)its[i].TermFreq();
      double tfnorm = (freq * (1.2 + 1)) / (freq + ix.cache[(unsigned char)norm]);
      double term_doc_score = idfs[i] * tfnorm;
      final_doc_score += term_doc_score;
    }
    return final_doc_score;
  }

  // RankDoc, query_processing.h:588-603 (strict > against the heap minimum).
  void Rank(int32_t doc) {
    double s = Score(doc);
    if (heap.size() < (size_t)k) {
      heap.push({doc, s});
    } else if (s > heap.top().score) {
      heap.pop();
      heap.push({doc, s});
    }
  }

  // SortHeap, query_processing.h:551-562
  std::vector<HeapEntry> Sorted() {
    std::vector<HeapEntry> r;
    int kk = k;
    while (!heap.empty() && kk != 0) {
      r.push_back(heap.top());
      heap.pop();
      kk--;
    }
    std::vector<HeapEntry> rev(r.rbegin(), r.rend());
    return rev;
  }

  // HandleTheFoundDoc for phrase queries, query_processing.h:886-895: rank the doc only if
  // PhraseQueryProcessor2 finds >= 1 match. No Bloom filters here: IsPossibleToPresent answers
  // "possible" when the index has none (flash_iterators.h:1039-1058).
  bool is_phrase = false;
  std::vector<std::vector<uint32_t>> pos;
  bool PhraseMatches() {
    const size_t n = its.size();
    pos.resize(n);
    for (size_t i = 0; i < n; i++) its[i].Positions(&pos[i]);
    if (n == 2) {
      // PhraseQueryProcessor2::ProcessTwoTerm, query_processing.h:282-331
      size_t i0 = 0, i1 = 0;
      long pos0 = -100, pos1 = -200;
      bool tried_pop_end = false;
      int matches = 0;
      while (!tried_pop_end) {
        if (pos0 < pos1) {
          if (i0 < pos[0].size()) pos0 = pos[0][i0++]; else tried_pop_end = true;
        } else if (pos0 > pos1) {
          if (i1 < pos[1].size()) pos1 = (long)pos[1][i1++] - 1; else tried_pop_end = true;
        } else {
          matches++;
          if (i0 < pos[0].size()) pos0 = pos[0][i0++]; else tried_pop_end = true;
          if (i1 < pos[1].size()) pos1 = (long)pos[1][i1++] - 1; else tried_pop_end = true;
        }
      }
      return matches > 0;
    }
    // PhraseQueryProcessor2::ProcessGeneral, query_processing.h:333-362 with
    // InitializeLastPopped / FindMaxAdjustedLastPopped / MovePoppedBeyond / IsPoppedMatch
    std::vector<size_t> nxt(n, 0);
    std::vector<long> last(n);
    for (size_t i = 0; i < n; i++) {
      if (pos[i].empty()) return false;
      last[i] = pos[i][nxt[i]++];
    }
    auto move_beyond = [&](long target) {
      for (size_t i = 0; i < n; i++) {
        while (nxt[i] < pos[i].size() && last[i] - (long)i < target) last[i] = pos[i][nxt[i]++];
        if (nxt[i] >= pos[i].size() && last[i] - (long)i < target) return false;
      }
      return true;
    };
    int matches = 0;
    for (;;) {
      long mx = last[0];
      for (size_t i = 1; i < n; i++) mx = std::max(mx, last[i] - (long)i);
      if (!move_beyond(mx)) break;
      bool match = true;
      for (size_t i = 0; i < n; i++) match = match && (last[i] - (long)i == mx);
      if (match) {
        matches++;
        if (!move_beyond(mx + 1)) break;
      }
    }
    return matches > 0;
  }
  void Found(int32_t doc) {
    if (is_phrase && its.size() > 1) {
      if (PhraseMatches()) Rank(doc);
    } else {
      Rank(doc);
    }
  }

  // SingleTermQueryProcessor::Process, query_processing.h:632-641
  void One() {
    auto &it = its[0];
    while (!it.IsEnd()) {
      Rank((int32_t)it.DocId());
      it.Advance();
    }
  }
  // TwoTermNonPhraseQueryProcessor::Process, query_processing.h:656-677
  void Two() {
    auto &a = its[0], &b = its[1];
    while (!a.IsEnd() && !b.IsEnd()) {
      int32_t d0 = (int32_t)a.DocId(), d1 = (int32_t)b.DocId();
      if (d0 > d1) {
        b.SkipForward(d0);
      } else if (d0 < d1) {
        a.SkipForward(d1);
      } else {
        Found(d0);      // non-phrase: RankDoc (:669); phrase: QueryProcessor::ProcessTwoTerm (:742-763)
        a.Advance();
        b.Advance();
      }
    }
  }
  // QueryProcessor::ProcessMultipleTerms / FindMax / FindMatch, query_processing.h:710-728, 810-852
  void Many() {
    const int n = (int)its.size();
    for (;;) {
      int32_t max_doc = -1;
      for (int i = 0; i < n; i++) {
        if (its[i].IsEnd()) return;
        int32_t d = (int32_t)its[i].DocId();
        if (d > max_doc) max_doc = d;
      }
      for (int i = 0; i < n; i++) {
        its[i].SkipForward(max_doc);
        if (its[i].IsEnd()) return;
        if ((int32_t)its[i].DocId() != max_doc) break;
        if (i == n - 1) {
          Found(max_doc);
          for (int j = 0; j < n; j++) its[j].Advance();
        }
      }
    }
  }
};

// VacuumEngine::Search, vacuum_engine.h:201-258 (non-phrase; snippets never requested).
int Search(const Index &ix, const char *const *terms, const size_t *lens, int n_terms, int k,
           std::vector<HeapEntry> *out, std::vector<int32_t> *dfs, bool is_phrase = false) {
  out->clear();
  dfs->clear();
  if (k == 0) return 0;                                  // :206-208
  std::vector<ListCursor> its;
  its.reserve(n_terms);
  for (int i = 0; i < n_terms; i++) {                    // FindIteratorsSolid, :89-99
    auto f = ix.term_to_off.find(std::string(terms[i], lens[i]));
    if (f == ix.term_to_off.end()) continue;
    its.emplace_back();
    if (!its.back().Reset(ix.vacuum.p, f->second)) return -2;
    // :217-219 — the posting-list size; in partition mode the collection-wide df of the term
    auto g = ix.global_df.find(f->second);
    dfs->push_back(g == ix.global_df.end() ? (int32_t)its.back().Size() : g->second);
  }
  if (its.empty() || (int)its.size() < n_terms) { dfs->clear(); return 0; }  // :213-215
  Processor p(ix, its, k, *dfs);
  p.is_phrase = is_phrase;
  // qq_search::ProcessQueryDelta, query_processing.h:956-979
  if (its.size() == 1) p.One();
  else if (its.size() == 2) p.Two();
  else p.Many();
  *out = p.Sorted();
  return 0;
}

}  // namespace

struct wsr_oracle_index { Index ix; };

extern "C" {

wsr_oracle_index *wsr_oracle_open(const char *dir, char *err, size_t errlen) {
  auto *h = new wsr_oracle_index;
  std::string e;
  if (!h->ix.Load(dir, &e)) {
    if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
    delete h;
    return nullptr;
  }
  return h;
}
void wsr_oracle_close(wsr_oracle_index *h) { delete h; }
int wsr_oracle_num_docs(const wsr_oracle_index *h) { return h->ix.doc_count; }
double wsr_oracle_avg_doc_len(const wsr_oracle_index *h) { return h->ix.avg_len; }
int wsr_oracle_set_global_stats(wsr_oracle_index *h, int n_docs_global, double avg_len_global) {
  if (n_docs_global <= 0 || !(avg_len_global > 0)) return -1;
  h->ix.doc_count = n_docs_global;
  h->ix.avg_len = avg_len_global;
  h->ix.BuildCache();
  return 0;
}
int wsr_oracle_set_global_df(wsr_oracle_index *h, const char *term, size_t len, int32_t df_global) {
  auto f = h->ix.term_to_off.find(std::string(term, len));
  if (f == h->ix.term_to_off.end()) return 1;
  h->ix.global_df[f->second] = df_global;
  return 0;
}
int wsr_oracle_term_count(const wsr_oracle_index *h) { return (int)h->ix.term_to_off.size(); }
int wsr_oracle_term_at(const wsr_oracle_index *h, int i, char *buf, int cap) {
  if (i < 0 || (size_t)i >= h->ix.term_order.size()) return -1;
  const std::string &t = h->ix.term_order[i];
  memcpy(buf, t.data(), t.size() < (size_t)cap ? t.size() : (size_t)cap);
  return (int)t.size();
}
int wsr_oracle_norm_byte(const wsr_oracle_index *h, int doc) {
  if (doc < 0 || (size_t)doc >= h->ix.norms.size()) return -1;
  return (unsigned char)h->ix.norms[doc];
}
int64_t wsr_oracle_term_df(const wsr_oracle_index *h, const char *term, size_t len) {
  auto f = h->ix.term_to_off.find(std::string(term, len));
  if (f == h->ix.term_to_off.end()) return -1;
  ListCursor c;
  if (!c.Reset(h->ix.vacuum.p, f->second)) return -2;
  return c.Size();
}
int64_t wsr_oracle_decode_list(const wsr_oracle_index *h, const char *term, size_t len,
                               uint32_t *docs, uint32_t *tfs, size_t cap) {
  auto f = h->ix.term_to_off.find(std::string(term, len));
  if (f == h->ix.term_to_off.end()) return -1;
  ListCursor c;
  if (!c.Reset(h->ix.vacuum.p, f->second)) return -2;
  size_t i = 0;
  while (!c.IsEnd()) {
    if (i < cap) {
      docs[i] = c.DocId();
      tfs[i] = c.TermFreq();
    }
    i++;
    c.Advance();
  }
  return c.Size();
}
int wsr_oracle_search_ex(const wsr_oracle_index *h, const char *const *terms, const size_t *lens,
                         int n_terms, int k, int is_phrase, int32_t *out_docs, double *out_scores,
                         size_t cap, int *n_hits, int32_t *doc_freqs, int *n_df);
int wsr_oracle_search(const wsr_oracle_index *h, const char *const *terms, const size_t *lens,
                      int n_terms, int k, int32_t *out_docs, double *out_scores, size_t cap,
                      int *n_hits, int32_t *doc_freqs, int *n_df) {
  return wsr_oracle_search_ex(h, terms, lens, n_terms, k, 0, out_docs, out_scores, cap, n_hits,
                              doc_freqs, n_df);
}
int wsr_oracle_search_ex(const wsr_oracle_index *h, const char *const *terms, const size_t *lens,
                         int n_terms, int k, int is_phrase, int32_t *out_docs, double *out_scores,
                         size_t cap, int *n_hits, int32_t *doc_freqs, int *n_df) {
  if (n_terms < 0 || k < 0) return -1;
  std::vector<HeapEntry> r;
  std::vector<int32_t> dfs;
  int rc = Search(h->ix, terms, lens, n_terms, k, &r, &dfs, is_phrase != 0);
  if (rc) return rc;
  size_t n = r.size() < cap ? r.size() : cap;
  for (size_t i = 0; i < n; i++) {
    out_docs[i] = r[i].doc;
    out_scores[i] = r[i].score;
  }
  *n_hits = (int)n;
  *n_df = (int)dfs.size();
  for (size_t i = 0; i < dfs.size(); i++) doc_freqs[i] = dfs[i];
  return 0;
}
double wsr_oracle_time_batch(const wsr_oracle_index *h, const char *const *terms,
                             const size_t *lens, const int64_t *q_off, int64_t n, int k,
                             int threads, uint64_t *listed_postings) {
  std::vector<uint64_t> listed(threads, 0);
  std::vector<std::thread> th;
  auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < threads; t++) {
    th.emplace_back([&, t]() {
      std::vector<HeapEntry> r;
      std::vector<int32_t> dfs;
      uint64_t l = 0;
      for (int64_t i = n * t / threads; i < n * (t + 1) / threads; i++) {
        Search(h->ix, terms + q_off[i], lens + q_off[i], (int)(q_off[i + 1] - q_off[i]), k, &r, &dfs);
        for (auto d : dfs) l += d;
      }
      listed[t] = l;
    });
  }
  for (auto &x : th) x.join();
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t tot = 0;
  for (auto l : listed) tot += l;
  if (listed_postings) *listed_postings = tot;
  return s;
}

}  // extern "C"
