// Index loader: vacuum directory -> HostIndex (see host_index.h for the layout).
// File formats follow SURVEY.md §5.1; the reference readers they replace are cited inline
// (paths relative to the reference's src/qq_mem/src/).
#include "host_index.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

namespace wsr {
namespace {

struct FileView {
  const uint8_t *data = nullptr;
  size_t size = 0;
  ~FileView() { if (data) munmap((void *)data, size); }
  bool Map(const std::string &path) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    bool ok = fstat(fd, &st) == 0;
    if (ok && st.st_size > 0) {
      void *m = mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      ok = m != MAP_FAILED;
      if (ok) { data = (const uint8_t *)m; size = st.st_size; }
    }
    close(fd);
    return ok;
  }
};

// LEB128 reader with bounds (utils.h:230-266 is the reference decoder).
struct ByteCursor {
  const uint8_t *p, *end;
  bool ok = true;
  uint64_t Varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
};

inline int BitWidth(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

// Reads one blob of the doc-id or tf column into out[0..n): a 128-value pack
// (0xD6, bits, 16*bits bytes: LittlePackedIntsReader, packed_value.h:184-235) or the VInts
// tail (0x9B, varint n_bytes, LEB128 values: VIntsIterator, packed_value.h:400-460).
bool ReadBlob(const uint8_t *file, size_t file_size, uint64_t off, int n, uint32_t *out) {
  if (off + 2 > file_size) return false;
  const uint8_t *b = file + off;
  if (b[0] == 0xD6) {
    const int bits = b[1];
    if (bits > 32 || off + 2 + 16ull * bits > file_size) return false;
    const uint8_t *s = b + 2;
    const size_t nbytes = 16ull * bits;
    const uint64_t mask = bits == 32 ? 0xffffffffull : ((1ull << bits) - 1);
    uint64_t bitpos = 0;
    for (int i = 0; i < n; i++, bitpos += bits) {
      const size_t byte = bitpos >> 3;
      uint64_t w;
      if (byte + 8 <= nbytes) {
        memcpy(&w, s + byte, 8);
      } else {
        w = 0;
        memcpy(&w, s + byte, nbytes - byte);
      }
      out[i] = (uint32_t)((w >> (bitpos & 7)) & mask);
    }
    return true;
  }
  if (b[0] == 0x9B) {
    ByteCursor c{b + 1, file + file_size};
    uint64_t nbytes = c.Varint();
    if (!c.ok || c.p + nbytes > file + file_size) return false;
    c.end = c.p + nbytes;
    for (int i = 0; i < n; i++) out[i] = (uint32_t)c.Varint();
    return c.ok;
  }
  return false;
}

// Appends the doc records of one block: per lane [f:w0][d1:b][d2:b][d3:b] LSB first in R words.
void PackDocRecords(const uint32_t *first, const uint32_t (*deltas)[3], int nl, const BlockShape &sh,
                    std::vector<uint8_t> *out) {
  const size_t at = out->size();
  out->resize(at + sh.doc_bytes(), 0);
  uint32_t *w = reinterpret_cast<uint32_t *>(out->data() + at);
  const int R = sh.rec_words();
  for (int l = 0; l < nl; l++) {
    unsigned __int128 x = first[l];
    int sft = sh.w0;
    for (int i = 0; i < 3; i++) { x |= (unsigned __int128)deltas[l][i] << sft; sft += sh.b; }
    for (int k = 0; k < R; k++) w[l * R + k] = (uint32_t)(x >> (32 * k));
  }
}

// Appends the tf records of one block: 4 tfs per lane at 4 / 8 / 32 bits each.
void PackTfRecords(const uint32_t (*tf)[4], int nl, const BlockShape &sh, std::vector<uint8_t> *out) {
  const size_t at = out->size();
  out->resize(at + sh.tf_bytes(), 0);
  uint8_t *d = out->data() + at;
  for (int l = 0; l < nl; l++) {
    if (sh.tcode == 0) {
      const uint16_t v = (uint16_t)(tf[l][0] | tf[l][1] << 4 | tf[l][2] << 8 | tf[l][3] << 12);
      memcpy(d + 2 * l, &v, 2);
    } else if (sh.tcode == 1) {
      const uint32_t v = tf[l][0] | tf[l][1] << 8 | tf[l][2] << 16 | tf[l][3] << 24;
      memcpy(d + 4 * l, &v, 4);
    } else {
      memcpy(d + 16 * l, tf[l], 16);
    }
  }
}

struct Chunk {               // output of one slice of terms
  size_t term_begin = 0, term_end = 0;
  std::vector<ListInfo> lists;
  std::vector<uint32_t> filters;
  std::vector<uint64_t> list_flt;
  std::vector<uint64_t> list_alg_bytes;
  std::vector<BlockInfo> blk_info;
  std::vector<uint32_t> blk_last;
  std::vector<uint16_t> blk_heads;   // 8 per block
  std::vector<uint8_t> payload;
  std::vector<uint32_t> positions;   // optional position column, absolute in-document positions
  std::vector<uint32_t> blk_pos;     // per block: index of its first position
  std::vector<uint16_t> grp_pos;     // per block 8 entries: positions before record 4g
  bool no_positions = false;         // some list of the chunk has no position column
  int64_t postings = 0, postings_global = 0;
  std::string err;
};

struct Builder {
  uint32_t filter_ppw = kFilterPostingsPerWord;   // per 32-bit filter word (WSR_FILTER_PPW overrides, 1..8)
  bool want_positions = false;
  const FileView &vac;
  const std::vector<uint64_t> &list_offs;
  const HostIndex &ix;      // norms / cache already filled
  const float *tfn_tab;     // [64][256] exact-rounded-up tfn for tf < 64
  uint32_t doc_lo, doc_hi;
  uint32_t filter_min_df = kFilterMinDf;   // shorter lists get no filter (WSR_FILTER_MIN_DF overrides)

  float TfnUpper(uint32_t tf, uint8_t norm) const {
    if (tf < 64) return tfn_tab[tf * 256 + norm];
    double x = (tf * (1.2 + 1)) / (tf + ix.cache[norm]);
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
  }

  bool BuildList(uint64_t off, Chunk *c, std::vector<uint32_t> *docs, std::vector<uint32_t> *tfs,
                 std::vector<uint32_t> &pos_scratch) {
    // Posting-list header: VacuumPostingListIterator::ResetWithZoneInfo, flash_iterators.h:903-956
    if (off + 10 > vac.size || vac.data[off] != 0xF4) { c->err = "bad posting-list magic"; return false; }
    ByteCursor cur{vac.data + off + 1, vac.data + vac.size};
    const uint64_t df = cur.Varint();
    cur.p += 8;  // two reserved Bloom skip-list offsets
    // Skip list: SkipList::Load, flash_containers.h:354-391
    if (!cur.ok || cur.p >= cur.end || *cur.p != 0xA3) { c->err = "bad skip-list magic"; return false; }
    cur.p++;
    const uint64_t n_rows = cur.Varint();
    if (n_rows != (df + kBlock - 1) / kBlock) { c->err = "skip rows != ceil(df/128)"; return false; }
    docs->resize(df);
    tfs->resize(df);
    uint64_t prev_doc = 0, docid_off = 0, tf_off = 0, pos_col_off = 0;
    for (uint64_t r = 0; r < n_rows; r++) {
      prev_doc += cur.Varint();
      docid_off += cur.Varint();
      tf_off += cur.Varint();
      const uint64_t pos_off = cur.Varint();      // position blob offset (delta vs previous row)
      cur.Varint();                               // position in-blob index
      cur.Varint(); cur.Varint();                 // offset column (snippets only)
      if (r == 0) pos_col_off = pos_off;          // the column starts at posting 0's blob
      if (!cur.ok) { c->err = "truncated skip list"; return false; }
      const int n = (int)std::min<uint64_t>(kBlock, df - r * kBlock);
      uint32_t *d = docs->data() + r * kBlock;
      if (!ReadBlob(vac.data, vac.size, docid_off, n, d) ||
          !ReadBlob(vac.data, vac.size, tf_off, n, tfs->data() + r * kBlock)) {
        c->err = "bad doc-id/tf blob";
        return false;
      }
      // deltas are taken over the whole list; each block restarts from the skip row's
      // previous_doc_id (DeltaEncodedPackedIntsIterator::Reset, packed_value.h:328-333)
      uint32_t run = (uint32_t)prev_doc;
      for (int i = 0; i < n; i++) { run += d[i]; d[i] = run; }
    }
    // Position column (phrase queries): all bags' values back to back, delta-coded inside each
    // bag, cut into 128-value packs + a VInts tail (GeneralTermEntry::GetCozyBoxWriter,
    // flash_engine_dumper.h:78-104; read side CozyBoxIterator, flash_iterators.h:280-425).
    std::vector<uint32_t> &pos_vals = pos_scratch;
    pos_vals.clear();
    // a column offset of 0 means the index was written without positions (wsr_gen_corpus
    // without --positions): phrase queries are then refused at planning time
    if (want_positions && pos_col_off == 0) c->no_positions = true;
    if (want_positions && pos_col_off != 0) {
      uint64_t total = 0;
      for (uint64_t i = 0; i < df; i++) total += (*tfs)[i];
      pos_vals.resize(total);
      uint64_t off = pos_col_off, done = 0;
      while (done < total) {
        const int n = (int)std::min<uint64_t>(kBlock, total - done);
        if (!ReadBlob(vac.data, vac.size, off, n, pos_vals.data() + done)) { c->err = "bad position blob"; return false; }
        if (vac.data[off] == 0xD6) off += 2 + 16ull * vac.data[off + 1];
        done += n;
      }
      uint64_t at = 0;
      for (uint64_t i = 0; i < df; i++) {          // deltas -> absolute positions inside each bag
        uint32_t run = 0;
        for (uint32_t j = 0; j < (*tfs)[i]; j++) { run += pos_vals[at]; pos_vals[at++] = run; }
      }
    }
    // shard filter: contiguous doc-id range
    size_t a = 0, b = df;
    if (ix.n_shards > 1) {
      a = std::lower_bound(docs->begin(), docs->end(), doc_lo) - docs->begin();
      b = std::lower_bound(docs->begin(), docs->end(), doc_hi) - docs->begin();
    }
    ListInfo li;
    li.first_block = (uint32_t)c->blk_info.size();
    li.df_shard = (uint32_t)(b - a);
    li.df_global = (uint32_t)df;
    li.n_blocks = (uint32_t)((b - a + kBlock - 1) / kBlock);
    uint64_t alg = 0;
    uint32_t base = doc_lo;  // shard 0: 0, as in the reference
    uint64_t pos_at = 0;     // index into pos_vals of posting s's first position
    const bool keep_pos = want_positions && pos_col_off != 0;
    if (keep_pos)
      for (size_t i = 0; i < a; i++) pos_at += (*tfs)[i];
    for (size_t s = a; s < b; s += kBlock) {
      const int n = (int)std::min<size_t>(kBlock, b - s);
      const int nl = (n + 3) / 4;
      uint32_t first[32], deltas[32][3], tfr[32][4];
      uint32_t fmax = 0, dmax = 0, tmax = 0, refmax = 0, p = base;
      float mx = 0.f;
      for (int i = 0; i < n; i++) {
        const uint32_t doc = (*docs)[s + i], tf = (*tfs)[s + i];
        const bool very_first = (s == a && i == 0);
        if (doc >= ix.norms.size() || (very_first ? doc < p : doc <= p)) {
          c->err = "doc ids not strictly increasing or out of range";
          return false;
        }
        refmax |= doc - p;                      // the reference packs consecutive deltas
        if ((i & 3) == 0) { first[i >> 2] = doc - base; fmax |= doc - base; }
        else { deltas[i >> 2][(i & 3) - 1] = doc - p; dmax |= doc - p; }
        tfr[i >> 2][i & 3] = tf;
        tmax |= tf;
        p = doc;
        mx = std::max(mx, TfnUpper(tf, ix.norms[doc]));
      }
      for (int i = n; i < 4 * nl; i++) {         // pad the last record with the last posting
        deltas[i >> 2][(i & 3) - 1] = 0;
        tfr[i >> 2][i & 3] = (*tfs)[s + n - 1];
      }
      BlockShape sh;
      sh.w0 = std::max(1, BitWidth(fmax));
      sh.b = std::max(1, BitWidth(dmax));
      sh.n = n;
      const int rec_bits = sh.w0 + 3 * sh.b;
      sh.rcode = rec_bits <= 32 ? 0 : rec_bits <= 64 ? 1 : rec_bits <= 96 ? 3 : 2;
      sh.tcode = tmax < 16 ? 0 : tmax < 256 ? 1 : 2;
      sh.ref_dbits = std::max(1, BitWidth(refmax));
      sh.ref_tbits = std::max(1, BitWidth(tmax));
      BlockInfo bi;
      bi.base_doc = base;
      bi.payload_off16 = (uint32_t)(c->payload.size() / 16);
      bi.bits = PackShape(sh);
      bi.max_tfn = mx;
      PackDocRecords(first, deltas, nl, sh, &c->payload);
      PackTfRecords(tfr, nl, sh, &c->payload);
      c->blk_info.push_back(bi);
      c->blk_last.push_back(p);
      // group heads: first doc (relative to base) of records 0, 4, .., 28 when they fit 16 bits
      for (int t = 0; t < 8; t++)
        c->blk_heads.push_back(sh.w0 <= 16 && 4 * t < nl ? (uint16_t)first[4 * t] : (uint16_t)0xFFFF);
      if (want_positions && !keep_pos) {
        c->blk_pos.push_back(0u);
        c->grp_pos.insert(c->grp_pos.end(), 8, (uint16_t)0);
      }
      if (keep_pos) {
        if (c->positions.size() > 0xFFFFFFF0ull) { c->err = "more than 2^32 positions"; return false; }
        c->blk_pos.push_back((uint32_t)c->positions.size());
        uint64_t cnt = 0;
        uint16_t gp[8];
        for (int i = 0; i < 4 * 32; i++) {
          if ((i & 15) == 0) gp[i >> 4] = (uint16_t)std::min<uint64_t>(cnt, 0xFFFF);
          if (i < n) cnt += (*tfs)[s + i];
        }
        // a block with 65535+ positions keeps the sentinel everywhere: the kernel then sums tfs
        if (cnt >= 0xFFFF)
          for (int g = 1; g < 8; g++) gp[g] = 0xFFFF;
        c->grp_pos.insert(c->grp_pos.end(), gp, gp + 8);
        c->positions.insert(c->positions.end(), pos_vals.begin() + pos_at, pos_vals.begin() + pos_at + cnt);
        pos_at += cnt;
      }
      alg += AlgorithmicBytes(sh);
      base = p;
    }
    // doc-range-partitioned Bloom filter of this (shard-local) list
    uint64_t flt = 0xFFFFFFFFull << 32;
    if (b - a >= filter_min_df) {
      const uint64_t range = (uint64_t)doc_hi - doc_lo;
      uint32_t g = 0;
      while (g < 31 && (range >> (g + 1)) >= (uint64_t)(b - a) / filter_ppw + 1) g++;   // postings per word
      const size_t words = (size_t)(range >> g) + 1;
      const size_t at = c->filters.size();
      c->filters.resize(at + words, 0u);
      uint32_t *w = c->filters.data() + at;
      for (size_t i = a; i < b; i++) {
        const uint32_t doc = (*docs)[i];
        w[(doc - doc_lo) >> g] |= FilterBits(doc);
      }
      flt = ((uint64_t)g << 32) | (uint32_t)at;
    }
    c->list_flt.push_back(flt);
    c->lists.push_back(li);
    c->list_alg_bytes.push_back(alg);
    c->postings += (int64_t)(b - a);
    c->postings_global += (int64_t)df;
    return true;
  }

  void BuildChunk(Chunk *c) {
    std::vector<uint32_t> docs, tfs, pos_scratch;
    for (size_t t = c->term_begin; t < c->term_end; t++)
      if (!BuildList(list_offs[t], c, &docs, &tfs, pos_scratch)) return;
  }
};

}  // namespace

uint64_t TermDict::Hash(const char *s, size_t len) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < len; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001b3ull; }
  return h ^ (h >> 29);
}

void TermDict::Build(const std::vector<char> *arena, const std::vector<uint64_t> *offs) {
  arena_ = arena;
  offs_ = offs;
  const size_t n = offs->size() - 1;
  size_t cap = 16;
  while (2 * cap < 3 * n) cap <<= 1;   // load factor <= 2/3 (the device copy is 8 B per slot)
  mask_ = cap - 1;
  slots_.assign(cap, 0xFFFFFFFFu);
  for (size_t t = 0; t < n; t++) {
    const char *s = arena->data() + (*offs)[t];
    const size_t len = (*offs)[t + 1] - (*offs)[t];
    uint64_t h = Hash(s, len) & mask_;
    while (slots_[h] != 0xFFFFFFFFu) h = (h + 1) & mask_;   // later duplicates never shadow
    slots_[h] = (uint32_t)t;
  }
}

uint32_t TermDict::Find(const char *s, size_t len) const {
  if (!offs_) return 0xFFFFFFFFu;
  uint64_t h = Hash(s, len) & mask_;
  for (;;) {
    const uint32_t t = slots_[h];
    if (t == 0xFFFFFFFFu) return t;
    const uint64_t a = (*offs_)[t], b = (*offs_)[t + 1];
    if (b - a == len && memcmp(arena_->data() + a, s, len) == 0) return t;
    h = (h + 1) & mask_;
  }
}

bool LoadVacuumDir(const std::string &dir, int shard, int n_shards, int threads,
                   HostIndex *out, std::string *err, int flags) {
  HostIndex &ix = *out;
  if (n_shards < 1 || shard < 0 || shard >= n_shards) { *err = "bad shard/n_shards"; return false; }
  ix.shard = shard;
  ix.n_shards = n_shards;

  // ---- my.doc_length: DocLengthCharStore::Deserialize, doc_length_store.h:164-190
  {
    FileView f;
    if (!f.Map(dir + "/my.doc_length") || f.size < 12) { *err = "cannot read my.doc_length"; return false; }
    int32_t count;
    memcpy(&count, f.data, 4);
    memcpy(&ix.avg_len, f.data + 4, 8);
    if (count < 0 || 12 + 5ull * count > f.size) { *err = "my.doc_length truncated"; return false; }
    ix.norms.assign(count, 0);
    ix.n_docs = 0;
    const uint8_t *p = f.data + 12;
    for (int i = 0; i < count; i++, p += 5) {
      int32_t id;
      memcpy(&id, p, 4);
      if (id < 0) { *err = "negative doc id in my.doc_length"; return false; }
      if ((size_t)id >= ix.norms.size()) ix.norms.resize(id + 1, 0);
      ix.norms[id] = p[4];
      ix.n_docs++;
    }
    // The reference indexes its 256-entry cache with a SIGNED char (scoring.h:65-69): norm
    // bytes >= 128 (doc length >= 2^19 tokens) are undefined behaviour there; reject them.
    for (uint8_t b : ix.norms)
      if (b >= 128) { *err = "doc length >= 2^19 tokens is undefined in the reference"; return false; }
  }
  // Bm25Similarity::BuildCache, scoring.h:85-90 — k1*(1 - b + b*len/avg), left to right.
  {
    const double k1 = 1.2, b = 0.75;
    for (int i = 0; i < 256; i++) {
      const uint32_t m = i & 7;
      const int sh = (i >> 3) - 1;
      const uint32_t len = sh < 0 ? m : (m | 8u) << sh;   // Char4ToUint, utils.h:317-329
      ix.cache[i] = k1 * (1 - b + b * len / ix.avg_len);
    }
  }
  const int64_t N = (int64_t)ix.norms.size();
  ix.doc_lo = (int32_t)(N * shard / n_shards);
  ix.doc_hi = (int32_t)(N * (shard + 1) / n_shards);

  // ---- my.tip: TermTrieIndex::Load, term_index.h:106-159; value = zone pages << 48 | offset
  std::vector<uint64_t> list_offs;
  {
    FileView f;
    if (!f.Map(dir + "/my.tip")) { *err = "cannot read my.tip"; return false; }
    const uint8_t *p = f.data, *e = f.data + f.size;
    ix.term_off.assign(1, 0);
    while (p < e) {
      if (p + 4 > e) { *err = "my.tip truncated"; return false; }
      uint32_t len;
      memcpy(&len, p, 4);
      if (p + 4 + len + 8 > e) { *err = "my.tip truncated"; return false; }
      ix.term_arena.insert(ix.term_arena.end(), p + 4, p + 4 + len);
      ix.term_off.push_back(ix.term_arena.size());
      uint64_t v;
      memcpy(&v, p + 4 + len, 8);
      list_offs.push_back(v & 0xffffffffffffull);
      p += 4 + len + 8;
    }
  }
  ix.dict.Build(&ix.term_arena, &ix.term_off);
  const size_t n_terms = list_offs.size();

  // ---- my.vacuum
  FileView vac;
  if (!vac.Map(dir + "/my.vacuum") || vac.size < 100 || vac.data[0] != 0x88) {
    *err = "cannot read my.vacuum (or bad magic 0x88)";
    return false;
  }

  std::vector<float> tfn_tab(64 * 256, 0.f);
  for (int tf = 1; tf < 64; tf++)
    for (int nb = 0; nb < 256; nb++) {
      double x = (tf * (1.2 + 1)) / (tf + ix.cache[nb]);
      float fl = (float)x;
      if ((double)fl < x) fl = nextafterf(fl, INFINITY);
      tfn_tab[tf * 256 + nb] = fl;
    }

  // Slice the terms into chunks of roughly equal FILE bytes (lists are laid out in my.tip
  // order), processed by a dynamic pool so one huge list does not serialise the load.
  std::vector<Chunk> chunks;
  {
    const uint64_t target = std::max<uint64_t>(1 << 20, vac.size / 1024);
    size_t begin = 0;
    while (begin < n_terms) {
      size_t end = begin + 1;
      while (end < n_terms && list_offs[end] - list_offs[begin] < target && end - begin < 65536) end++;
      Chunk c;
      c.term_begin = begin;
      c.term_end = end;
      chunks.push_back(std::move(c));
      begin = end;
    }
  }
  Builder builder{kFilterPostingsPerWord, false, vac, list_offs, ix, tfn_tab.data(), (uint32_t)ix.doc_lo, (uint32_t)ix.doc_hi};
  builder.want_positions = (flags & kLoadPositions) != 0;
  ix.has_positions = builder.want_positions;
  if (const char *e = getenv("WSR_FILTER_PPW")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 8) builder.filter_ppw = (uint32_t)v;
  }
  if (const char *e = getenv("WSR_FILTER_MIN_DF")) {
    const int v = atoi(e);
    if (v >= 1) builder.filter_min_df = (uint32_t)v;
  }
  if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  threads = (int)std::min<size_t>(threads, std::max<size_t>(1, chunks.size()));
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= chunks.size()) return;
      builder.BuildChunk(&chunks[i]);
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
  }
  size_t tot_blocks = 0, tot_payload = 0, tot_flt = 0, tot_pos = 0;
  for (auto &c : chunks) {
    if (!c.err.empty()) { *err = c.err; return false; }
    tot_blocks += c.blk_info.size();
    tot_payload += c.payload.size();
    tot_flt += c.filters.size();
    tot_pos += c.positions.size();
    if (c.no_positions) ix.has_positions = false;
  }
  if (tot_pos >= 0xFFFFFFF0ull) { *err = "positions exceed 2^32 entries on one shard"; return false; }
  if (tot_flt >= 0xFFFFFFF0ull) { *err = "filters exceed 2^32 words"; return false; }
  if (tot_payload / 16 > 0xFFFFFFF0ull) { *err = "payload exceeds 64 GiB addressable by u32 offsets"; return false; }
  if (tot_blocks >= (1ull << 25)) { *err = "more than 2^25 blocks on one shard (hit records pack block<<7|slot)"; return false; }

  // ---- concatenate chunk outputs (block indices and payload offsets re-based)
  ix.lists.resize(n_terms);
  ix.list_alg_bytes.resize(n_terms);
  ix.blk_info.resize(tot_blocks);
  ix.blk_last.resize(tot_blocks);
  ix.blk_heads.resize(tot_blocks * 8 + 8);
  ix.payload.assign(tot_payload + 1024, 0);  // tail pad: prefetchers read up to 512 B past a block
  ix.filters.assign(tot_flt + 1, 0u);
  ix.list_flt.resize(n_terms);
  if (ix.has_positions) {
    ix.positions.assign(tot_pos + 1, 0u);
    ix.blk_pos.assign(tot_blocks + 1, 0u);
    ix.grp_pos.assign((tot_blocks + 1) * 8, (uint16_t)0);
  }
  std::vector<size_t> blk_base(chunks.size()), pay_base(chunks.size()), flt_base(chunks.size()),
      pos_base(chunks.size());
  size_t bb = 0, pb = 0, fb = 0, qb = 0;
  for (size_t i = 0; i < chunks.size(); i++) {
    blk_base[i] = bb;
    pay_base[i] = pb;
    flt_base[i] = fb;
    pos_base[i] = qb;
    qb += chunks[i].positions.size();
    bb += chunks[i].blk_info.size();
    pb += chunks[i].payload.size();
    fb += chunks[i].filters.size();
    ix.n_postings += chunks[i].postings;
    ix.n_postings_global += chunks[i].postings_global;
  }
  next = 0;
  auto merger = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= chunks.size()) return;
      Chunk &c = chunks[i];
      const uint32_t b0 = (uint32_t)blk_base[i], p0 = (uint32_t)(pay_base[i] / 16);
      for (size_t j = 0; j < c.lists.size(); j++) {
        ListInfo li = c.lists[j];
        li.first_block += b0;
        ix.lists[c.term_begin + j] = li;
        ix.list_alg_bytes[c.term_begin + j] = c.list_alg_bytes[j];
        uint64_t f = c.list_flt[j];
        if ((f >> 32) != 0xFFFFFFFFull) f = (f & 0xFFFFFFFF00000000ull) | (uint32_t)((f & 0xFFFFFFFFull) + flt_base[i]);
        ix.list_flt[c.term_begin + j] = f;
      }
      if (!c.filters.empty())
        memcpy(ix.filters.data() + flt_base[i], c.filters.data(), c.filters.size() * 4);
      std::vector<uint32_t>().swap(c.filters);
      if (ix.has_positions) {
        for (size_t j = 0; j < c.blk_pos.size(); j++) ix.blk_pos[b0 + j] = c.blk_pos[j] + (uint32_t)pos_base[i];
        if (!c.grp_pos.empty()) memcpy(ix.grp_pos.data() + b0 * 8, c.grp_pos.data(), c.grp_pos.size() * 2);
        std::vector<uint16_t>().swap(c.grp_pos);
        if (!c.positions.empty())
          memcpy(ix.positions.data() + pos_base[i], c.positions.data(), c.positions.size() * 4);
        std::vector<uint32_t>().swap(c.positions);
      }
      for (size_t j = 0; j < c.blk_info.size(); j++) {
        BlockInfo bi = c.blk_info[j];
        bi.payload_off16 += p0;
        ix.blk_info[b0 + j] = bi;
      }
      if (!c.blk_last.empty())
        memcpy(ix.blk_last.data() + b0, c.blk_last.data(), c.blk_last.size() * 4);
      if (!c.blk_heads.empty())
        memcpy(ix.blk_heads.data() + b0 * 8, c.blk_heads.data(), c.blk_heads.size() * 2);
      std::vector<uint16_t>().swap(c.blk_heads);
      if (!c.payload.empty()) memcpy(ix.payload.data() + pay_base[i], c.payload.data(), c.payload.size());
      Chunk().payload.swap(c.payload);
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(merger);
    merger();
    for (auto &t : pool) t.join();
  }
  return true;
}

}  // namespace wsr
