// Test helper (CPU): loads a vacuum directory through the PRODUCT loader (host_index.cc),
// decodes the resulting HBM-layout blocks with a plain host loop and writes
// { u32 term_len; term; u32 df_shard; df x {u32 doc; u32 tf} } per term, so that the layout
// can be checked against the reference iterators' dump without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../wiser_b200/csrc/host_index.h"

using namespace wsr;

// Host restatement of the lane-major block layout documented in host_index.h (test only).
static void DecodeBlock(const HostIndex &ix, const BlockInfo &bi, uint32_t *docs, uint32_t *tfs, int *n_out) {
  const BlockShape sh = UnpackShape(bi.bits);
  const uint8_t *p = ix.payload.data() + (size_t)bi.payload_off16 * 16;
  const uint8_t *pt = p + sh.doc_bytes();
  const int R = sh.rec_words();
  const uint64_t m0 = sh.w0 == 32 ? 0xffffffffull : ((1ull << sh.w0) - 1);
  const uint64_t mb = sh.b == 32 ? 0xffffffffull : ((1ull << sh.b) - 1);
  for (int l = 0; l < sh.nl(); l++) {
    unsigned __int128 x = 0;
    for (int k = 0; k < R; k++) {
      uint32_t w;
      memcpy(&w, p + (size_t)(l * R + k) * 4, 4);
      x |= (unsigned __int128)w << (32 * k);
    }
    uint32_t d = bi.base_doc + (uint32_t)((uint64_t)x & m0);
    x >>= sh.w0;
    for (int i = 0; i < 4; i++) {
      if (i) { d += (uint32_t)((uint64_t)x & mb); x >>= sh.b; }
      uint32_t tf;
      if (sh.tcode == 0) { uint16_t v; memcpy(&v, pt + 2 * l, 2); tf = (v >> (4 * i)) & 15; }
      else if (sh.tcode == 1) { uint32_t v; memcpy(&v, pt + 4 * l, 4); tf = (v >> (8 * i)) & 255; }
      else { memcpy(&tf, pt + 16 * l + 4 * i, 4); }
      if (4 * l + i < sh.n) { docs[4 * l + i] = d; tfs[4 * l + i] = tf; }
    }
  }
  *n_out = sh.n;
}

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: host_index_dump <dir> <shard> <n_shards> <out.bin>\n"); return 2; }
  HostIndex ix;
  std::string err;
  if (!LoadVacuumDir(argv[1], atoi(argv[2]), atoi(argv[3]), 4, &ix, &err)) {
    fprintf(stderr, "load failed: %s\n", err.c_str());
    return 1;
  }
  FILE *out = fopen(argv[4], "wb");
  for (size_t t = 0; t < ix.lists.size(); t++) {
    uint32_t len = (uint32_t)(ix.term_off[t + 1] - ix.term_off[t]);
    fwrite(&len, 4, 1, out);
    fwrite(ix.term_arena.data() + ix.term_off[t], 1, len, out);
    const ListInfo &li = ix.lists[t];
    fwrite(&li.df_shard, 4, 1, out);
    uint32_t total = 0;
    for (uint32_t b = 0; b < li.n_blocks; b++) {
      const BlockInfo &bi = ix.blk_info[li.first_block + b];
      uint32_t docs[128], tfs[128];
      int n;
      DecodeBlock(ix, bi, docs, tfs, &n);
      for (int i = 0; i < n; i++) {
        uint32_t rec[2] = {docs[i], tfs[i]};
        fwrite(rec, 4, 2, out);
      }
      if (docs[n - 1] != ix.blk_last[li.first_block + b]) { fprintf(stderr, "blk_last mismatch\n"); return 3; }
      {
        const BlockShape sh = UnpackShape(bi.bits);
        for (int t = 0; t < 8; t++) {
          const uint16_t want = sh.w0 <= 16 && 16 * t < n ? (uint16_t)(docs[16 * t] - bi.base_doc) : (uint16_t)0xFFFF;
          if (ix.blk_heads[(size_t)(li.first_block + b) * 8 + t] != want) { fprintf(stderr, "blk_heads mismatch\n"); return 5; }
        }
      }
      total += n;
    }
    if (total != li.df_shard) { fprintf(stderr, "df mismatch\n"); return 4; }
  }
  fclose(out);
  printf("%d %d %lld %lld %zu %d %d\n", ix.n_docs, (int)ix.lists.size(), (long long)ix.n_postings,
         (long long)ix.n_postings_global, ix.blk_info.size(), ix.doc_lo, ix.doc_hi);
  return 0;
}
