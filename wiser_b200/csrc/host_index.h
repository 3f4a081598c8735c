// Host-side image of the HBM index layout, and the loader that builds it from a vacuum
// index directory (my.tip / my.vacuum / my.doc_length).
//
// HBM layout (all arrays flat, one cudaMalloc each). A block holds up to 128 postings and is
// LANE-MAJOR: lane l of the decoding warp owns postings 4l..4l+3, so a block decodes with one
// coalesced record load per lane and three lane-local adds — no shared-memory staging and no
// cross-lane prefix sum (ncu on the first, reference-layout kernel showed the path is bound
// by instruction issue, not HBM: ~70 warp instructions per block went into unpack + scan).
//   payload   u8[]      per block, 16 B aligned:
//                         doc records  nl = ceil(n/4) records of R words (R in {1,2,3,4}); record l =
//                                      [f : w0 bits][d1 : b][d2 : b][d3 : b] LSB first, where
//                                      f  = doc[4l] - base_doc, d_i = doc[4l+i] - doc[4l+i-1];
//                                      padded to 16 B
//                         tf records   nl records of 16 / 32 / 128 bits = 4 tfs of 4 / 8 / 32 bits;
//                                      padded to 16 B
//                       slots past n in the last record repeat the last posting (delta 0, same tf)
//   blk_info  uint4[]   {base_doc, payload_off/16, bits, max_tfn f32}; bits =
//                       (w0-1) | (b-1)<<5 | (n-1)<<10 | rcode<<17 | tcode<<19 |   (rcode 0/1/2/3 =
//                       records of 1/2/4/3 words; 96-bit records serve the sparse lists, whose
//                       records need 65-90 bits)
//                       (ref_dbits-1)<<21 | (ref_tbits-1)<<26
//                       base_doc = doc id the block is relative to (the reference's skip-row
//                       previous_doc_id, flash_containers.h:22-30; shard lower bound for a
//                       shard's first block); ref_dbits/ref_tbits = the widths the REFERENCE's
//                       128-value packs use for this block (packed_value.h:87-128) — they define
//                       the algorithmic bytes of the roofline; max_tfn = upper bound of
//                       tf*(k1+1)/(tf+cache[norm]) over the block (block-max metadata).
//   blk_last  u32[]     last doc id of each block (searched when skipping)
//   blk_heads u16[8][]  per block the first doc (minus base_doc) of records 0, 4, .., 28 (0xFFFF past
//                       the block's records, and for every entry when w0 > 16): an exact probe
//                       picks its 4-record group with this ONE 16-byte load instead of seven
//                       loads strided over the whole record stream (8 sectors -> 1)
//   blk_max   f32[]     copy of max_tfn as its own array: single-term queries scan it coalesced
//                       to bound the k-th score before touching any payload
//   lists     uint4[]   per term {first_block, n_blocks, df_shard, df_global}
//   list_flt  uint2[]   per term {first filter word, shift g | 0xFFFFFFFF = no filter}
//   filters   u32[]     per list (df >= kFilterMinDf) a doc-range-partitioned Bloom filter, ~8 bits
//                       per posting (false-positive rate ~3 %): doc d sets the three bits of
//                       FilterPattern(FilterIndex(d)) in word (d - doc_lo) >> g, g chosen for
//                       ~kFilterPostingsPerWord postings per word. An AND query tests the driver's
//                       candidates against the other lists' filters first, so the exact probe (block
//                       lookup + record search) runs only for the few percent that may be present.
//                       No false negatives.
//   positions u32[]     (optional, phrase queries) token positions of every posting, absolute
//                       inside the document, postings back to back in list order — the
//                       reference's position column (delta-coded "cozy box" packs addressed
//                       through the skip list, flash_iterators.h:558-661) with the deltas summed
//   blk_pos   u32[]     index into positions[] of each block's first posting; a posting's run
//                       starts at blk_pos + sum of the tfs before it in the block
//   grp_pos   u16[8][]  per block: number of positions before record 4g (postings 0..16g-1), so the
//                       run of posting 4r+i starts at blk_pos + grp_pos[r/4] + the tfs of the up to
//                       three records before r in its group + the i tfs before it in the record (the
//                       tf records of a group are adjacent: 8 or 16 bytes); all 0xFFFF (g >= 1) when
//                       the block holds >= 65535 positions
//   positions are kept as u16 when every in-document position of the shard is < 65536 (always,
//   unless a document has more than 65535 tokens), else as u32
//   norms     u8[]      DocLengthCharStore bytes indexed by GLOBAL doc id
//   cache     f64[256]  Bm25Similarity::cache_ (scoring.h:85-90)
#ifndef WSR_HOST_INDEX_H
#define WSR_HOST_INDEX_H
#include <cstdint>
#include <string>
#include <vector>

namespace wsr {

constexpr int kBlock = 128;
constexpr uint32_t kDecodeStageBytes = 8192;   // K1 streams the payload in stages of this size (kernels.cu)

struct BlockInfo {       // 16 B, mirrors a device uint4
  uint32_t base_doc;
  uint32_t payload_off16;
  uint32_t bits;         // PackShape()
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

struct BlockShape {      // decoded view of BlockInfo::bits
  int w0, b, n, rcode, tcode, ref_dbits, ref_tbits;
  int nl() const { return (n + 3) / 4; }
  int rec_words() const { return rcode == 3 ? 3 : 1 << rcode; }  // rcode 0, 1, 2, 3 -> 1, 2, 4, 3 words
  int tf_bits() const { return tcode == 0 ? 4 : tcode == 1 ? 8 : 32; }
  uint32_t doc_bytes() const { return (uint32_t)((nl() * rec_words() * 4 + 15) / 16 * 16); }
  uint32_t tf_bytes() const { return (uint32_t)((nl() * tf_bits() / 2 + 15) / 16 * 16); }
};
inline uint32_t PackShape(const BlockShape &s) {
  return (uint32_t)(s.w0 - 1) | ((uint32_t)(s.b - 1) << 5) | ((uint32_t)(s.n - 1) << 10) |
         ((uint32_t)s.rcode << 17) | ((uint32_t)s.tcode << 19) |
         ((uint32_t)(s.ref_dbits - 1) << 21) | ((uint32_t)(s.ref_tbits - 1) << 26);
}
inline BlockShape UnpackShape(uint32_t bits) {
  BlockShape s;
  s.w0 = (int)(bits & 31) + 1;
  s.b = (int)((bits >> 5) & 31) + 1;
  s.n = (int)((bits >> 10) & 127) + 1;
  s.rcode = (int)((bits >> 17) & 3);
  s.tcode = (int)((bits >> 19) & 3);
  s.ref_dbits = (int)((bits >> 21) & 31) + 1;
  s.ref_tbits = (int)((bits >> 26) & 31) + 1;
  return s;
}
// Algorithmic bytes of a block: the reference's pack sizes for its n postings
// (16*bits bytes for a full pack; a tail re-packed at its own width, padded to 16 B) + 16 B
// of block metadata.
inline uint32_t RefStreamBytes(int n, int bits) {
  return (uint32_t)(((uint64_t)n * bits + 127) / 128 * 16);
}
// Bloom filter bit pattern of a doc inside its filter word (identical on host and device): three
// distinct bits, taken from a 1024-entry table indexed by the top 10 bits of a multiplicative hash
// of the doc id. The kernels keep the table in shared memory: a tested posting costs one multiply,
// one shift and one LDS instead of eight ALU operations for three computed bit positions.
constexpr int kFilterPatterns = 1024;
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t FilterPattern(uint32_t t) {
  uint32_t x = t * 0x9E3779B1u + 0x7F4A7C15u;
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  uint32_t a = x & 31u, b = (x >> 5) & 31u, c = (x >> 10) & 31u;
  if (b == a) b = (b + 1u) & 31u;
  while (c == a || c == b) c = (c + 1u) & 31u;
  return (1u << a) | (1u << b) | (1u << c);
}
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t FilterIndex(uint32_t doc) { return (doc * 0x9E3779B1u) >> 22; }
inline uint32_t FilterBits(uint32_t doc) { return FilterPattern(FilterIndex(doc)); }
// Lists shorter than this are probed directly (WSR_FILTER_MIN_DF overrides). With 4 postings per
// filter word and this floor the filters of the C2 corpus take 0.9 GB instead of 2.0 GB (2 per word,
// floor 256) for +1.5 % on the two-term step (profiles/r2_notes.md).
constexpr uint32_t kFilterMinDf = 1024;
constexpr uint32_t kFilterPostingsPerWord = 4;

inline uint32_t AlgorithmicBytes(const BlockShape &s) {
  return RefStreamBytes(s.n, s.ref_dbits) + RefStreamBytes(s.n, s.ref_tbits) + 16;
}

// Open-addressing string -> term id table over one arena (replaces the reference's hat-trie,
// term_index.h:100-159, for lookups only).
class TermDict {
 public:
  void Build(const std::vector<char> *arena, const std::vector<uint64_t> *offs);
  // returns term id or 0xFFFFFFFF
  uint32_t Find(const char *s, size_t len) const;
  size_t Size() const { return offs_ ? offs_->size() - 1 : 0; }
  // FNV-1a 64 with a final fold; slot = Hash & Mask(), linear probing. The device copy of the
  // table (frontend.cu) uses the same function and probe order.
  static uint64_t Hash(const char *s, size_t len);
  const std::vector<uint32_t> &Slots() const { return slots_; }
  uint64_t Mask() const { return mask_; }
 private:
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
  std::vector<uint16_t> blk_heads;  // 8 per block (see layout comment)
  std::vector<uint8_t> payload;
  std::vector<uint32_t> filters;
  std::vector<uint64_t> list_flt;     // low 32: first word, high 32: shift (0xFFFFFFFF none)
  bool has_positions = false;
  std::vector<uint32_t> positions;
  std::vector<uint32_t> blk_pos;
  std::vector<uint16_t> grp_pos;    // 8 per block
  std::vector<uint64_t> list_alg_bytes;  // algorithmic bytes of all blocks of each list
  int64_t n_postings = 0, n_postings_global = 0;
  int shard = 0, n_shards = 1;
  int32_t doc_lo = 0, doc_hi = 0;
};

// Parses the vacuum directory and builds the layout for one shard. Multi-threaded over terms.
// Returns false and sets *err on malformed input.
enum LoadFlags { kLoadPositions = 1 };
bool LoadVacuumDir(const std::string &dir, int shard, int n_shards, int threads,
                   HostIndex *out, std::string *err, int flags = 0);

}  // namespace wsr
#endif
