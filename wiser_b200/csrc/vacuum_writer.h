// Writer of a vacuum index directory in the REFERENCE's on-disk format: the
// posting-list encoder (mirrors VacuumInvertedIndexDumper::DumpPostingListNoBloom,
// flash_engine_dumper.h:339-411, with GetCozyBoxWriter :78-104 and SkipListWriter
// flash_containers.h:236-308) and the directory assembly (my.vacuum, my.tip, my.doc_length, stub
// doc store). Shared by the synthetic corpus generator (corpus_gen.cc) and the linedoc indexer
// (linedoc_index.cc). The doc-id and tf columns are always written, the position column on
// request; the offset column is never written (skip rows carry zeros), so the directory serves
// queries without snippets — exactly the path under test.
#ifndef WSR_VACUUM_WRITER_H
#define WSR_VACUUM_WRITER_H
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace wsrw {

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

// Appends one posting list (term id `r`; docs ascending, tfs, and with `positions` the absolute
// in-document positions of every posting back to back) to the chunk buffer.
inline void EncodeList(uint64_t r, const std::vector<uint32_t> &docs, const std::vector<uint32_t> &tfs,
                     const std::vector<uint32_t> &pos, bool positions, Chunk *c,
                     std::vector<uint32_t> *delta, std::vector<uint32_t> *rows_d,
                     std::vector<uint32_t> *rows_t) {
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
  if (positions) {
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
    sz += 1 + 7 + 7 + (positions ? 7 : 1) + 3;   // row 0
    for (size_t rr = 1; rr < n_rows; rr++) {
      sz += VarLen(docs[rr * 128 - 1] - (rr >= 2 ? docs[(rr - 1) * 128 - 1] : 0));
      sz += VarLen((*rows_d)[rr] - (*rows_d)[rr - 1]);
      sz += VarLen((*rows_t)[rr] - (*rows_t)[rr - 1]);
      if (positions) sz += VarLen(prow_off[rr] - prow_off[rr - 1]) + VarLen(prow_idx[rr]) + 2;
      else sz += 4;
    }
    data_at = sz;
  }
  (void)skip_at;
  b.push_back(0);                          // row 0: previous_doc_id = 0
  rec.patch_at = (uint32_t)(b.size() - start);
  b.insert(b.end(), 14, 0);                // docid blob abs offset, tf blob abs offset (patched)
  b.insert(b.end(), positions ? 7 : 1, 0);   // pos blob abs offset (patched) or 0 = no positions
  b.insert(b.end(), 3, 0);                 // pos idx (0), offset blob off, offset idx
  for (size_t rr = 1; rr < n_rows; rr++) {
    PutVarint(&b, docs[rr * 128 - 1] - (rr >= 2 ? docs[(rr - 1) * 128 - 1] : 0));
    PutVarint(&b, (*rows_d)[rr] - (*rows_d)[rr - 1]);
    PutVarint(&b, (*rows_t)[rr] - (*rows_t)[rr - 1]);
    if (positions) {
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
  rec.pos0 = positions ? (uint32_t)(data_at + pos_col) : 0u;
  rec.tf_end = (uint32_t)(data_at + pos_col);
  b.insert(b.end(), cols.begin(), cols.end());
  c->lists.push_back(rec);
  c->postings += df;
}


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


// Lays the chunks out behind the 100-byte header, patches the absolute offsets of every list's
// first skip row, and writes my.vacuum, my.tip, terms.txt ("term df" per line), my.doc_length and
// the stub doc store. name(term id) gives the term bytes; doc_len[d] is the token count of doc d.
inline bool WriteVacuumDir(const std::string &out, std::vector<Chunk> &chunks,
                           const std::function<std::string(uint32_t)> &name,
                           const std::vector<uint32_t> &doc_len, int threads, uint64_t *file_size_out,
                           uint64_t *n_lists_out, uint64_t *postings_out) {
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
      const std::string nm = name(l.term);
      const int len = (int)nm.size();
      const uint32_t ulen = (uint32_t)len;
      const uint64_t pages = std::min<uint64_t>(l.tf_end / 4096, 0xffff);
      const uint64_t val = (pages << 48) | abs;
      tip.insert(tip.end(), (const uint8_t *)&ulen, (const uint8_t *)&ulen + 4);
      tip.insert(tip.end(), nm.begin(), nm.end());
      tip.insert(tip.end(), (const uint8_t *)&val, (const uint8_t *)&val + 8);
      terms_txt += nm;
      terms_txt += ' ';
      terms_txt += std::to_string(l.df);
      terms_txt += '\n';
    }
  }

  // my.vacuum
  {
    const std::string path = out + "/my.vacuum";
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0 || ftruncate(fd, (off_t)file_size) != 0) { perror("my.vacuum"); return false; }
    uint8_t header[100];
    memset(header, 0, sizeof(header));
    header[0] = 0x88;
    if (pwrite(fd, header, 100, 0) != 100) { perror("pwrite"); return false; }
    std::atomic<size_t> next{0};
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
    for (int t = 1; t < std::min(threads, 8); t++) pool.emplace_back(writer);
    writer();
    for (auto &t : pool) t.join();
    close(fd);
    if (!ok) { fprintf(stderr, "write error on my.vacuum\n"); return false; }
  }
  if (!WriteFile(out + "/my.tip", tip.data(), tip.size()) ||
      !WriteFile(out + "/terms.txt", terms_txt.data(), terms_txt.size())) {
    perror("my.tip");
    return false;
  }

  // my.doc_length: i32 count, f64 avg (running mean in doc order), count x {i32 id, i8 char4}
  {
    const uint64_t N = doc_len.size();
    std::vector<uint8_t> dl(12 + 5 * N);
    double avg = 0;
    for (uint64_t d = 0; d < N; d++) {
      const int len = (int)doc_len[d];
      avg = avg + (len - avg) / (double)(d + 1);           // DocLengthCharStore::AddLength
      const int32_t id = (int32_t)d;
      memcpy(&dl[12 + 5 * d], &id, 4);
      dl[12 + 5 * d + 4] = UintToChar4((uint32_t)len);
    }
    const int32_t count = (int32_t)N;
    memcpy(&dl[0], &count, 4);
    memcpy(&dl[4], &avg, 8);
    if (!WriteFile(out + "/my.doc_length", dl.data(), dl.size())) { perror("my.doc_length"); return false; }
  }
  // stub doc store: 0 documents (ChunkedDocStoreReader::LoadFdx, doc_store.h:365-392)
  {
    const uint8_t fdx[4] = {0x00, 0x80, 0x80, 0x01};
    const uint8_t fdt[1] = {0};
    if (!WriteFile(out + "/my.fdx", fdx, 4) || !WriteFile(out + "/my.fdt", fdt, 1)) return false;
  }
  *file_size_out = file_size;
  *n_lists_out = n_lists;
  *postings_out = postings;
  return true;
}

}  // namespace wsrw
#endif
