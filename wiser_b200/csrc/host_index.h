// Host-side image of the HBM index layout, and the loader that builds it from a vacuum
// index directory (my.tip / my.vacuum / my.doc_length).
//
// HBM layout (all arrays flat, one cudaMalloc each):
//   payload   u8[]      per block: doc-id delta bitstream then tf bitstream, each the
//                       reference's LSB-first little-endian b-bit stream (value i at bit i*b,
//                       packed_value.h:87-128), each padded to 16 B; full blocks are 16*b B.
//   blk_info  uint4[]   {base_doc, payload_off/16, bits(dbits|tbits<<6|(n-1)<<12), max_tfn f32}
//                       base_doc = doc id the first delta is relative to (skip row
//                       previous_doc_id, flash_containers.h:22-30; shard lower bound for the
//                       first block of a shard); max_tfn = upper bound of
//                       tf*(k1+1)/(tf+cache[norm]) over the block (block-max metadata).
//   blk_last  u32[]     last doc id of each block (binary-searched when skipping)
//   lists     uint4[]   per term {first_block, n_blocks, df_shard, df_global}
//   norms     u8[]      DocLengthCharStore bytes indexed by GLOBAL doc id
//   cache     f64[256]  Bm25Similarity::cache_ (scoring.h:85-90)
#ifndef WSR_HOST_INDEX_H
#define WSR_HOST_INDEX_H
#include <cstdint>
#include <string>
#include <vector>

namespace wsr {

constexpr int kBlock = 128;

struct BlockInfo {       // 16 B, mirrors a device uint4
  uint32_t base_doc;
  uint32_t payload_off16;
  uint32_t bits;         // dbits | tbits << 6 | (n-1) << 12
  float max_tfn;
};
static_assert(sizeof(BlockInfo) == 16, "BlockInfo must be 16 bytes");

struct ListInfo {        // 16 B, mirrors a device uint4
  uint32_t first_block;
  uint32_t n_blocks;
  uint32_t df_shard;
  uint32_t df_global;
};
static_assert(sizeof(ListInfo) == 16, "ListInfo must be 16 bytes");

inline uint32_t PackBits(int dbits, int tbits, int n) {
  return (uint32_t)dbits | ((uint32_t)tbits << 6) | ((uint32_t)(n - 1) << 12);
}
// Bytes of one b-bit stream holding n values, padded to the 16 B load granule.
inline uint32_t StreamBytes(int n, int bits) {
  return (uint32_t)(((uint64_t)n * bits + 127) / 128 * 16);
}

// Open-addressing string -> term id table over one arena (replaces the reference's hat-trie,
// term_index.h:100-159, for lookups only).
class TermDict {
 public:
  void Build(const std::vector<char> *arena, const std::vector<uint64_t> *offs);
  // returns term id or 0xFFFFFFFF
  uint32_t Find(const char *s, size_t len) const;
  size_t Size() const { return offs_ ? offs_->size() - 1 : 0; }
 private:
  static uint64_t Hash(const char *s, size_t len);
  const std::vector<char> *arena_ = nullptr;
  const std::vector<uint64_t> *offs_ = nullptr;
  std::vector<uint32_t> slots_;
  uint64_t mask_ = 0;
};

struct HostIndex {
  // global statistics (identical on every shard)
  int32_t n_docs = 0;          // DocLengthCharStore::Size()
  double avg_len = 0;
  double cache[256];
  std::vector<uint8_t> norms;
  // terms, my.tip order
  std::vector<char> term_arena;
  std::vector<uint64_t> term_off;   // n_terms + 1
  TermDict dict;
  std::vector<ListInfo> lists;
  // blocks of this shard
  std::vector<BlockInfo> blk_info;
  std::vector<uint32_t> blk_last;
  std::vector<uint8_t> payload;
  std::vector<uint64_t> list_alg_bytes;  // algorithmic bytes of all blocks of each list
  int64_t n_postings = 0, n_postings_global = 0;
  int shard = 0, n_shards = 1;
  int32_t doc_lo = 0, doc_hi = 0;
};

// Parses the vacuum directory and builds the layout for one shard. Multi-threaded over terms.
// Returns false and sets *err on malformed input.
bool LoadVacuumDir(const std::string &dir, int shard, int n_shards, int threads,
                   HostIndex *out, std::string *err);

}  // namespace wsr
#endif
