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

static uint32_t GetBits(const uint8_t *s, uint64_t bit, int bits) {
  uint64_t w = 0;
  memcpy(&w, s + (bit >> 3), 8);   // payload has a 64-byte tail pad
  uint64_t mask = bits == 32 ? 0xffffffffull : ((1ull << bits) - 1);
  return (uint32_t)((w >> (bit & 7)) & mask);
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
      int dbits = bi.bits & 63, tbits = (bi.bits >> 6) & 63, n = ((bi.bits >> 12) & 127) + 1;
      const uint8_t *p = ix.payload.data() + (size_t)bi.payload_off16 * 16;
      const uint8_t *pt = p + StreamBytes(n, dbits);
      uint32_t doc = bi.base_doc;
      for (int i = 0; i < n; i++) {
        doc += GetBits(p, (uint64_t)i * dbits, dbits);
        uint32_t rec[2] = {doc, GetBits(pt, (uint64_t)i * tbits, tbits)};
        fwrite(rec, 4, 2, out);
      }
      if (doc != ix.blk_last[li.first_block + b]) { fprintf(stderr, "blk_last mismatch\n"); return 3; }
      total += n;
    }
    if (total != li.df_shard) { fprintf(stderr, "df mismatch\n"); return 4; }
  }
  fclose(out);
  printf("%d %d %lld %lld %zu %d %d\n", ix.n_docs, (int)ix.lists.size(), (long long)ix.n_postings,
         (long long)ix.n_postings_global, ix.blk_info.size(), ix.doc_lo, ix.doc_hi);
  return 0;
}
