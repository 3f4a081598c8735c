// Hand-written sm_100a kernels for the WiSER/Vacuum hot path: packed-block decode, conjunctive
// intersection with skip metadata, fused BM25 scoring and per-query top-k.
//
// Execution model: one WARP owns one work unit = (query, range of blocks of the query's
// shortest list). Units of a whole query batch sit in per-class dynamic queues (an atomic
// counter each) drained by a persistent grid sized to the SM count x resident CTAs. Blocks are
// LANE-MAJOR (host_index.h): lane l loads one 32/64/128-bit record holding postings 4l..4l+3
// and reconstructs their doc ids with three adds. Intersection is galloping PER LANE: the
// driver (shortest) list's block is decoded into registers, and every lane probes the longer
// lists for its own four candidates — per-block last-doc skip metadata picks the block, a
// 5-step search over the block's record heads picks the 4-posting record, one record decode
// settles membership — so probe lists are never decoded wholesale and the work per driver
// block does not grow with the length ratio of the lists. Everything on the path is
// integer/byte work bound by instruction issue, memory latency and HBM bandwidth; tensor cores
// are not used (nothing here is a dense contraction).
//
// Reference semantics restated here (paths relative to the reference's src/qq_mem/src/):
//   block decode      LittlePackedIntsReader / DeltaEncodedPackedIntsIterator, packed_value.h:184-235, 320-369
//   AND               TwoTermNonPhraseQueryProcessor::Process, query_processing.h:656-677;
//                     QueryProcessor::ProcessMultipleTerms, :710-728, 810-852
//   BM25              CalcDocScoreLossy + TfNormLossy, scoring.h:65-69, 124-145 (fp64, no FMA)
//   top-k             RankDoc / SortHeap, query_processing.h:551-603 (strict > replacement);
//                     device order is (score desc, doc id asc) — the reference's order among
//                     equal scores is implementation-defined (SURVEY §7 "Ties").
#include "kernels.cuh"
#include "host_index.h"

#include <algorithm>
#include <cstddef>
#include <cub/device/device_segmented_sort.cuh>

namespace wsr {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kNoDoc = 0xffffffffu;
constexpr double kK1Plus1 = 1.2 + 1;   // (k1_ + 1) evaluated in double, scoring.h:68
constexpr int kCandCap = 160;          // survivors: 31 carried over + up to 128 of one driver block
constexpr int kHitCap = 64;            // hits: 31 queued + up to 32 of one probe batch

// ---- shared memory through 32-bit shared-window addresses --------------------------------------
// The scratch of a warp and the CTA's tables are reached through generic pointers handed down the
// (inlined) call chain; ptxas converts such a pointer back to a shared-window address in front of
// every access (S2R SR_CgaCtaId + MOV + LEA: 37 of the 574 warp instructions per driver block in
// the round-2 profile of the two-term kernel). The hot accesses therefore use explicit ld/st.shared
// on an address converted once per unit.
__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t LdsU32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void StsU32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 LdsU64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void StsU64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ uint4 LdsU128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void StsU128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// Asynchronous global -> shared copies (LDGSTS): the look-ahead loads of the two-term driver loop
// land in the warp's scratch instead of in registers. With 64 registers per thread ptxas spilled
// every looked-ahead value right behind its load (STL of a register an LDG had just been issued
// into), which waits for the load on the spot: 15 % of the kernel's stall samples in the round-2
// profile sat on such stores and the software pipeline hid nothing.
__device__ __forceinline__ void CpAsync4(uint32_t sa, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void CpAsync16(uint32_t sa, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- block shape (host_index.h PackShape) --------------------------------------------------
__device__ __forceinline__ uint32_t ShW0(uint32_t b) { return (b & 31u) + 1u; }
__device__ __forceinline__ uint32_t ShB(uint32_t b) { return ((b >> 5) & 31u) + 1u; }
__device__ __forceinline__ uint32_t ShN(uint32_t b) { return ((b >> 10) & 127u) + 1u; }
__device__ __forceinline__ uint32_t ShRcode(uint32_t b) { return (b >> 17) & 3u; }
__device__ __forceinline__ uint32_t ShTcode(uint32_t b) { return (b >> 19) & 3u; }
__device__ __forceinline__ uint32_t ShRefD(uint32_t b) { return ((b >> 21) & 31u) + 1u; }
__device__ __forceinline__ uint32_t ShRefT(uint32_t b) { return ((b >> 26) & 31u) + 1u; }
// 32-bit words of a doc record: rcode 0, 1, 2, 3 -> 1, 2, 4, 3 (host_index.h BlockShape::rec_words)
__device__ __forceinline__ uint32_t RecWords(uint32_t rc) { return rc == 3u ? 3u : 1u << rc; }
// 16-byte granules of the doc-record stream of a block
__device__ __forceinline__ uint32_t DocGranules(uint32_t bits) {
  const uint32_t nl = (ShN(bits) + 3u) >> 2;
  return (nl * RecWords(ShRcode(bits)) + 3u) >> 2;
}
// Algorithmic bytes of a block: the reference's pack sizes + 16 B metadata
__device__ __forceinline__ uint32_t AlgBytes(uint32_t bits, bool with_tf) {
  const uint32_t n = ShN(bits);
  uint32_t g = (n * ShRefD(bits) + 127u) >> 7;
  if (with_tf) g += (n * ShRefT(bits) + 127u) >> 7;
  return 16u * g + 16u;
}

// ---- K1: block decode ------------------------------------------------------------------------
// Doc ids of the four postings of record `rec` = [f:w0][d1:b][d2:b][d3:b] (host_index.h).
// Padded slots of a block's last record repeat its last doc.
__device__ __forceinline__ uint4 LoadRecord(const DevIndexView &ix, const uint4 info, uint32_t rec) {
  const uint32_t rc = ShRcode(info.z);
  const uint4 *src = ix.payload + info.y;
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (rc == 0) {
    r.x = __ldg(reinterpret_cast<const uint32_t *>(src) + rec);
  } else if (rc == 1) {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src) + rec);
    r.x = v.x; r.y = v.y;
  } else if (rc == 2) {
    r = __ldg(src + rec);
  } else {   // 96-bit records: three words, 4-byte aligned
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src) + 3u * rec;
    r.x = __ldg(w); r.y = __ldg(w + 1); r.z = __ldg(w + 2);
  }
  return r;
}

__device__ __forceinline__ void DecodeRaw(const uint4 info, const uint4 raw, uint32_t d[4]) {
  const uint32_t bits = info.z;
  const uint32_t w0 = ShW0(bits), b = ShB(bits), rc = ShRcode(bits);
  const uint32_t r0 = raw.x, r1 = raw.y, r2 = raw.z, r3 = raw.w;
  const uint32_t m0 = w0 >= 32u ? 0xffffffffu : ((1u << w0) - 1u);
  const uint32_t mb = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
  uint32_t f, d1, d2, d3;
  if (rc <= 1) {
    unsigned long long x = ((unsigned long long)r1 << 32) | r0;
    f = (uint32_t)x & m0;
    x >>= w0;
    d1 = (uint32_t)x & mb;
    x >>= b;
    d2 = (uint32_t)x & mb;
    x >>= b;
    d3 = (uint32_t)x & mb;
  } else {
    // 128-bit record: fields may straddle words
    f = r0 & m0;
    uint32_t o = w0;
    uint32_t out[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const uint32_t wi = o >> 5, sh = o & 31u;
      const uint32_t lo = wi == 0 ? r0 : wi == 1 ? r1 : wi == 2 ? r2 : r3;
      const uint32_t hi = wi == 0 ? r1 : wi == 1 ? r2 : wi == 2 ? r3 : 0u;
      out[i] = __funnelshift_r(lo, hi, sh) & mb;
      o += b;
    }
    d1 = out[0]; d2 = out[1]; d3 = out[2];
  }
  d[0] = info.x + f;
  d[1] = d[0] + d1;
  d[2] = d[1] + d2;
  d[3] = d[2] + d3;
}

__device__ __forceinline__ void DecodeRecord(const DevIndexView &ix, const uint4 info, uint32_t rec,
                                             uint32_t d[4]) {
  DecodeRaw(info, LoadRecord(ix, info, rec), d);
}

// Whole block: lane l decodes record l (postings 4l..4l+3); lanes past the records get kNoDoc.
__device__ __forceinline__ void DecodeDocs(const DevIndexView &ix, const uint4 info, int lane,
                                           uint32_t d[4]) {
  const uint32_t nl = (ShN(info.z) + 3u) >> 2;
  d[0] = d[1] = d[2] = d[3] = kNoDoc;
  if ((uint32_t)lane < nl) DecodeRecord(ix, info, (uint32_t)lane, d);
}

// tfs of postings 4l..4l+3 (0 for lanes past the block's records)
__device__ __forceinline__ void DecodeTfs(const DevIndexView &ix, const uint4 info, int lane,
                                          uint32_t tf[4]) {
  const uint32_t bits = info.z;
  const uint32_t nl = (ShN(bits) + 3u) >> 2, tc = ShTcode(bits);
  const uint4 *src = ix.payload + info.y + DocGranules(bits);
  tf[0] = tf[1] = tf[2] = tf[3] = 0;
  if ((uint32_t)lane < nl) {
    if (tc == 0) {
      const uint32_t v = __ldg(reinterpret_cast<const unsigned short *>(src) + lane);
      tf[0] = v & 15u; tf[1] = (v >> 4) & 15u; tf[2] = (v >> 8) & 15u; tf[3] = v >> 12;
    } else if (tc == 1) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src) + lane);
      tf[0] = v & 255u; tf[1] = (v >> 8) & 255u; tf[2] = (v >> 16) & 255u; tf[3] = v >> 24;
    } else {
      const uint4 v = __ldg(src + lane);
      tf[0] = v.x; tf[1] = v.y; tf[2] = v.z; tf[3] = v.w;
    }
  }
}

// tf of posting `pos` = (global block index << 7) | slot — random access for intersection hits.
__device__ __forceinline__ uint32_t TfAt(const DevIndexView &ix, uint32_t pos) {
  const uint4 info = __ldg(&ix.blk_info[pos >> 7]);
  const uint32_t bits = info.z, tc = ShTcode(bits), slot = pos & 127u;
  const uint4 *src = ix.payload + info.y + DocGranules(bits);
  if (tc == 0)
    return (__ldg(reinterpret_cast<const unsigned short *>(src) + (slot >> 2)) >> (4u * (slot & 3u))) & 15u;
  if (tc == 1)
    return (__ldg(reinterpret_cast<const uint32_t *>(src) + (slot >> 2)) >> (8u * (slot & 3u))) & 255u;
  return __ldg(reinterpret_cast<const uint32_t *>(src) + slot);
}

// TfNormLossy + one term of CalcDocScoreLossy, every operation rounded separately (the
// reference build has no FMA contraction).
__device__ __forceinline__ double TermScore(double idf, uint32_t tf, double cache_norm) {
  const double f = (double)tf;
  const double tfnorm = __ddiv_rn(__dmul_rn(f, kK1Plus1), __dadd_rn(f, cache_norm));
  return __dmul_rn(idf, tfnorm);
}

// ---- per-warp top-k: lane r holds the rank-r entry, ordered (score desc, doc asc) ------------
struct TopK {
  double s;
  int d;
  int count;   // warp-uniform
};
__device__ __forceinline__ void TopKInit(TopK &t) { t.s = -1.0; t.d = 0x7fffffff; t.count = 0; }
__device__ __forceinline__ void TopKInsert(TopK &t, int k, double s, int d, int lane) {
  const bool mine_first = (t.s > s) || (t.s == s && t.d < d);
  const int pos = __popc(__ballot_sync(kFull, mine_first));
  if (pos >= k) return;
  const double us = __shfl_up_sync(kFull, t.s, 1);
  const int ud = __shfl_up_sync(kFull, t.d, 1);
  if (lane > pos) { t.s = us; t.d = ud; }
  else if (lane == pos) { t.s = s; t.d = d; }
  if (t.count < k) t.count++;
}
__device__ __forceinline__ double TopKKth(const TopK &t, int k) {
  return __shfl_sync(kFull, t.s, k - 1);   // meaningful when count == k
}


// Work counters of a warp (roofline bookkeeping, never used for results). Counting costs ~7 % of
// the two-term kernel, so ordinary runs instantiate the kernels with ON = false and only the
// profiling pass (wsr_batch_profile) counts.
template <bool ON>
struct UnitStatsT {
  static constexpr bool kOn = ON;
  unsigned long long decoded, bytes, matches, probe_blocks;
  uint32_t last_probe;   // last partner block counted by the current unit (a block counts once per unit)
};
#define WSR_STAT(...) do { if constexpr (ST::kOn) { __VA_ARGS__ } } while (0)

struct CtaShared {
  double cache[256];
  float cache32[256];
  uint32_t fpat[kFilterPatterns];   // host_index.h FilterPattern
};

struct HitRec {
  uint32_t doc;
  uint32_t pos_a;   // (global block index << 7) | slot
  uint32_t pos_b;
};

// Per-warp scratch of the intersecting kernels.
struct CandRec {
  uint32_t doc;
  uint32_t pos_a;
};
struct __align__(16) ProbeScratch {
  uint32_t win[128];             // the probe list's blk_last window, for per-lane block lookup
  CandRec cand[kCandCap];        // filter survivors awaiting the exact probe (doc ascending)
  HitRec hits[kHitCap];          // intersection hits awaiting scoring
  // look-ahead staging of the two-term driver loop (cp.async targets, see ProcessTwo)
  uint4 inf[4];                  // blk_info of driver blocks ja .. ja+3, ring by (ja - b0) & 3
  uint32_t rec[2][128];          // doc-record stream of driver blocks ja+1 / ja+2 (<= 32 granules)
  uint32_t fw[2][128];           // filter words of blocks ja / ja+1, lane-private: 4 per lane
  uint32_t docs[128];            // decoded doc ids of block ja (+1 once staged), 4 per lane
};
struct __align__(16) NoScratch { uint32_t unused; };

// Merge-mode scratch of the two-term kernel (balanced lists, ProcessTwoMerge): the driver list's
// current blocks are a SET in shared memory that the partner list's decoded postings are tested
// against — a byte map hashed by the low doc-id bits (byte stores of the same value race
// benignly, so no shared-memory atomics), backed by a ring of the set blocks' doc ids and tfs
// that settles false positives and yields the driver posting's tf.
constexpr int kMapBytes = 4096;        // doc & (kMapBytes - 1)
constexpr int kRingBlocks = 4;         // driver blocks whose docs/tfs stay addressable
constexpr int kRingMask = kRingBlocks * 128 - 1;
constexpr int kSurvCap = 160;          // 31 queued + up to 128 of one partner block
constexpr int kMergeHitCap = 64;       // 31 queued + up to 32 of one verification pass
struct __align__(16) MergeScratch {
  uint8_t map[kMapBytes];
  uint32_t rdoc[kRingBlocks * 128];    // doc ids of driver blocks ja-4 .. ja-1, slot = posting index & kRingMask
  uint8_t rtf[kRingBlocks * 128];      // min(tf, 255); 255 = read the exact tf from the payload
  uint2 surv[kSurvCap];                // partner postings whose map byte was set: {doc, tf}
  uint32_t hdoc[kMergeHitCap];         // verified matches awaiting scoring
  uint32_t htfa[kMergeHitCap];         // tf in the driver list
  uint32_t htfb[kMergeHitCap];         // tf in the partner list
};

// Emits one unit's result: straight to the caller's hit array when the query has one unit,
// else to the unit's candidate slot for the merge pass. doc ids leave as GLOBAL ids.
__device__ __forceinline__ void EmitTopK(const BatchView &bv, const DevQuery &q, uint32_t local,
                                         const TopK &t, int lane) {
  wsr_hit h;
  h.doc_id = t.d + (int)bv.doc_base; h.reserved = 0; h.score = t.s;
  if (q.n_units == 1) {
    if (lane < t.count) bv.hits[(size_t)q.out_slot * bv.k_stride + lane] = h;
    if (lane == 0) bv.n_hits[q.out_slot] = t.count;
  } else {
    const uint32_t gunit = q.cand_begin + local;
    if (lane < t.count) bv.cand[(size_t)gunit * kMaxFastK + lane] = h;
    if (lane == 0) bv.cand_n[gunit] = t.count;
  }
}

// Collect mode: appends this lane's match to the query's segment.
__device__ __forceinline__ void CollectAppend(const BatchView &bv, const DevQuery &q, uint32_t qi,
                                              bool has, int doc, double score, int lane) {
  const unsigned m = __ballot_sync(kFull, has);
  if (!m) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&bv.seg_count[qi], (uint32_t)__popc(m));
  base = __shfl_sync(kFull, base, 0);
  if (has) {
    const uint32_t at = q.seg_begin + base + __popc(m & ((1u << lane) - 1u));
    bv.seg_doc[at] = doc + (int)bv.doc_base;
    bv.seg_score[at] = score;
  }
}

// Offers scored candidates (one per lane) to the unit's top-k and publishes the new k-th score.
__device__ __forceinline__ void OfferToTopK(const BatchView &bv, uint32_t qi, bool multi, int k, bool has,
                                            double s, int doc, TopK &top, double &published, int lane) {
  double thr = 0.0;
  if (multi) thr = __longlong_as_double((long long)__ldcg(&bv.thr[qi]));
  const double kth = top.count == k ? TopKKth(top, k) : -1.0;
  // a score strictly below another unit's k-th score can never reach the final top-k; ties stay
  unsigned m = __ballot_sync(kFull, has && !(s < thr) && !(s < kth));
  if (!m) return;
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    TopKInsert(top, k, __shfl_sync(kFull, s, src), __shfl_sync(kFull, doc, src), lane);
  }
  if (multi && top.count == k) {
    const double nk = TopKKth(top, k);
    if (nk > published) {
      published = nk;
      if (lane == 0) atomicMax(&bv.thr[qi], (unsigned long long)__double_as_longlong(nk));
    }
  }
}

// ---- single-term queries: SingleTermQueryProcessor::Process, query_processing.h:632-641 ------
// Every posting is a hit, so the work is choosing which blocks can hold a top-k document.
// Fast path (k <= 32, one warp per query): the list's block-max array is scanned coalesced to
// find the k-th largest block maximum — k different blocks each hold a document scoring at
// least that, so it bounds the k-th best score from below before any payload is read. Blocks
// whose block-max score is below the bound are skipped; survivors are decoded, pre-filtered
// with an fp32 upper bound, and only candidates are re-scored in exact fp64.
// Collect path (k > 32): all postings are scored and appended, split into units.
template <bool COLLECT, class ST>
__device__ void ProcessOneTerm(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                               uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                               const CtaShared *sh, int lane, ST &st) {
  const uint32_t term = q.term[0];
  const uint4 li = __ldg(&ix.lists[term]);
  const double idf = __ldg(&ix.idf[term]);
  const float idf_up = __double2float_ru(idf);
  const int k = (int)q.k;
  TopK top;
  TopKInit(top);
  double published = 0.0;
  double lb = 0.0;   // lower bound of the k-th best score of the whole list
  const float *__restrict__ bmax = ix.blk_max + li.x;

  if (!COLLECT && b1 - b0 > (uint32_t)k) {
    // k-th largest block maximum (TopK reused with the block index as tie-breaker)
    TopK bt;
    TopKInit(bt);
    for (uint32_t base = b0; base < b1; base += 32) {
      const uint32_t j = base + lane;
      const float v = j < b1 ? __ldg(bmax + j) : -1.f;
      const double kth = bt.count == k ? TopKKth(bt, k) : -1.0;
      unsigned m = __ballot_sync(kFull, (double)v > kth);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        TopKInsert(bt, k, (double)__shfl_sync(kFull, v, src), (int)(base + src), lane);
      }
    }
    WSR_STAT(st.bytes += 4ull * (b1 - b0););
    // the best document of a block scores idf * tfn with block_max >= tfn > block_max*(1 - 2^-23)
    if (bt.count == k) lb = __dmul_rn(__dmul_rn(idf, TopKKth(bt, k)), 1.0 - 3e-7);
  }

  for (uint32_t base = b0; base < b1; base += 32) {
    unsigned need;
    if (COLLECT) {
      need = base + 32 <= b1 ? kFull : ((1u << (b1 - base)) - 1u);
    } else {
      const uint32_t j = base + lane;
      const float v = j < b1 ? __ldg(bmax + j) : -1.f;
      // upper bound of every exact score in the block (rounded up at each step)
      need = __ballot_sync(kFull, j < b1 && (double)(idf_up * v * 1.00001f) >= lb);
    }
    while (need) {
      const uint32_t j = base + (uint32_t)(__ffs(need) - 1);
      need &= need - 1;
      const uint4 cur = __ldg(&ix.blk_info[li.x + j]);
      const uint32_t n = ShN(cur.z);
      double kth = -1.0;
      if (!COLLECT && top.count == k) {
        kth = TopKKth(top, k);
        // later blocks hold larger doc ids: a tie with the k-th entry cannot displace it
        if ((double)(idf_up * __uint_as_float(cur.w) * 1.00001f) <= kth) continue;
      }
      uint32_t d[4], tf[4];
      DecodeDocs(ix, cur, lane, d);
      DecodeTfs(ix, cur, lane, tf);
      WSR_STAT(st.decoded += n;);
      WSR_STAT(st.bytes += AlgBytes(cur.z, true) + n;);   // + one norm byte per posting
      WSR_STAT(st.matches += n;);
      bool pass[4];
      double s64[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool valid = 4u * lane + i < n;
        pass[i] = valid;
        s64[i] = 0.0;
        if (valid) {
          const uint32_t nb = __ldg(ix.norms + d[i]);
          if (!COLLECT) {
            const float f = (float)tf[i];
            const float s32 = idf_up * __fdividef(f * 2.2f, f + sh->cache32[nb]) * 1.00001f;
            pass[i] = (double)s32 >= fmax(lb, kth);
          }
          if (pass[i]) {
            s64[i] = __dadd_rn(0.0, TermScore(idf, tf[i], sh->cache[nb]));
            if (!COLLECT) pass[i] = !(s64[i] < lb) && !(s64[i] < kth);
          }
        }
      }
      if (COLLECT) {
#pragma unroll
        for (int i = 0; i < 4; i++) CollectAppend(bv, q, qi, pass[i], (int)d[i], s64[i], lane);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
          OfferToTopK(bv, qi, false, k, pass[i], s64[i], (int)d[i], top, published, lane);
      }
    }
  }
  if (!COLLECT) EmitTopK(bv, q, local, top, lane);
}

// ---- probe-list machinery shared by the intersecting kernels ---------------------------------
// Walk state of one probe list. A 32-entry window of its blk_last lives in registers (lane l
// holds last[wbase + l]); it moves forward with the driver, by a 32-ary cooperative search when
// the next candidate lies beyond it.
struct ProbeList {
  const uint32_t *last;   // blk_last of this list
  uint32_t first, nb;     // first global block, block count
  uint32_t wbase, wl;
};

__device__ __forceinline__ void ProbeInit(ProbeList &p, const DevIndexView &ix, const uint4 li, int lane) {
  p.first = li.x;
  p.nb = li.y;
  p.last = ix.blk_last + li.x;
  p.wbase = 0;
  p.wl = (uint32_t)lane < p.nb ? __ldg(p.last + lane) : kNoDoc;
}

// Moves the window so that it contains the first block j with last[j] >= x; returns j, or kNoDoc
// when the list has no doc >= x.
__device__ __forceinline__ uint32_t ProbeFind(ProbeList &p, uint32_t x, int lane) {
  if (x > __shfl_sync(kFull, p.wl, 31)) {
    uint32_t lo = p.wbase + 32u, hi = p.nb;
    if (lo >= hi) return kNoDoc;
    while (hi - lo > 32u) {
      const uint32_t step = (hi - lo + 31u) >> 5;
      const uint32_t idx = min(lo + ((uint32_t)lane + 1u) * step - 1u, hi - 1u);
      const int c = __popc(__ballot_sync(kFull, __ldg(p.last + idx) < x));
      if (c == 32) return kNoDoc;
      lo += (uint32_t)c * step;
      hi = min(lo + step, hi);
    }
    p.wbase = lo;
    p.wl = lo + lane < p.nb ? __ldg(p.last + lo + lane) : kNoDoc;
  }
  const int c = __popc(__ballot_sync(kFull, p.wl < x));
  const uint32_t j = p.wbase + (uint32_t)c;
  return (c == 32 || j >= p.nb) ? kNoDoc : j;
}

// Bloom pre-test of one candidate against a probe list's filter (no false negatives).
struct ListFilter {
  const uint32_t *words;   // nullptr: the list has no filter, everything passes
  uint32_t shift;
};
__device__ __forceinline__ ListFilter FilterOf(const DevIndexView &ix, uint32_t term) {
  const uint2 f = __ldg(&ix.list_flt[term]);
  ListFilter lf;
  lf.words = f.y == 0xffffffffu ? nullptr : ix.filters + f.x;
  lf.shift = f.y & 31u;
  return lf;
}
__device__ __forceinline__ bool FilterTest(const CtaShared *sh, uint32_t w, uint32_t doc) {
  const uint32_t need = sh->fpat[FilterIndex(doc)];   // host_index.h FilterBits
  return (w & need) == need;
}
__device__ __forceinline__ bool FilterTestS(uint32_t fpat_addr, uint32_t w, uint32_t doc) {
  const uint32_t need = LdsU32(fpat_addr + 4u * FilterIndex(doc));
  return (w & need) == need;
}
// Filter word of a candidate (all ones = "may be present" when the list has no filter). Slots past
// a block's postings decode to docs inside the shard's range, so their word may be read as well.
__device__ __forceinline__ uint32_t FilterWord(const DevIndexView &ix, const ListFilter &lf, uint32_t doc) {
  if (lf.words == nullptr) return 0xffffffffu;
  return __ldg(lf.words + ((doc - ix.doc_lo) >> lf.shift));
}

// Exact probe of ONE candidate per lane (filter survivors, doc ascending across lanes):
// skip metadata -> block; blk_heads (one 16-byte load) -> 4-record group; the group's records in
// one or two 16-byte loads -> record -> membership. Three dependent round trips, three sectors.
// Returns false when the list has nothing at or after the first candidate. `win`: shared-window
// address of the warp's 128-word scratch window.
template <class ST>
__device__ __forceinline__ bool ProbeOne(const DevIndexView &ix, ProbeList &p, uint32_t win, bool has,
                                         uint32_t x, bool *hit, uint32_t *pos, int lane,
                                         ST &st) {
  *hit = false;
  const uint32_t mn = __reduce_min_sync(kFull, has ? x : kNoDoc);
  const uint32_t mx = __reduce_max_sync(kFull, has ? x : 0u);
  const uint32_t j_lo = ProbeFind(p, mn, lane);
  if (j_lo == kNoDoc) return false;
  uint32_t j;
  if (mx <= __shfl_sync(kFull, p.wl, 31)) {
    // every candidate falls inside the 32-block register window
    __syncwarp();
    StsU32(win + 4u * (uint32_t)lane, p.wl);
    __syncwarp();
    uint32_t c = 0;
#pragma unroll
    for (uint32_t s = 16; s; s >>= 1)
      if (LdsU32(win + 4u * (c + s - 1u)) < x) c += s;
    c += LdsU32(win + 4u * c) < x;
    j = p.wbase + c;
    if (c >= 32u) has = false;
  } else {
    // the batch spans more: stage the next 96 entries too (coalesced, usually L2 hits) and search
    // the 128-entry window in shared memory; only candidates beyond it walk blk_last in global
    __syncwarp();
    StsU32(win + 4u * (uint32_t)lane, p.wl);
#pragma unroll
    for (uint32_t i = 1; i < 4; i++) {
      const uint32_t at = p.wbase + 32u * i + (uint32_t)lane;
      StsU32(win + 4u * (32u * i + (uint32_t)lane), at < p.nb ? __ldg(p.last + at) : kNoDoc);
    }
    __syncwarp();
    if (!has || x <= LdsU32(win + 4u * 127u)) {
      uint32_t c = 0;
#pragma unroll
      for (uint32_t s = 64; s; s >>= 1)
        if (LdsU32(win + 4u * (c + s - 1u)) < x) c += s;
      c += LdsU32(win + 4u * c) < x;
      j = p.wbase + c;
    } else {
      uint32_t lo = p.wbase + 128u, hi = p.nb;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(p.last + mid) < x) lo = mid + 1; else hi = mid;
      }
      j = lo;
    }
  }
  if (j >= p.nb) has = false;
  uint4 info = make_uint4(0u, 0u, 0u, 0u);
  if (has) {
    info = __ldg(&ix.blk_info[p.first + j]);
    const uint4 gh = __ldg(&ix.blk_heads[p.first + j]);
    const uint32_t bits = info.z;
    const uint32_t nl = (ShN(bits) + 3u) >> 2, rcs = ShRcode(bits), w0 = ShW0(bits);
    const uint32_t m0 = w0 >= 32u ? 0xffffffffu : ((1u << w0) - 1u);
    const uint32_t rel = x - info.x;     // x > base of its block
    uint32_t rec, e[4];
    if (w0 <= 16u && rcs <= 1u) {
      // group = number of group heads (records 4, 8, .., 28) that are <= rel; heads ascend
      const uint32_t grp = (uint32_t)(4u < nl && (gh.x >> 16) <= rel) + (uint32_t)(8u < nl && (gh.y & 0xffffu) <= rel) +
                           (uint32_t)(12u < nl && (gh.y >> 16) <= rel) + (uint32_t)(16u < nl && (gh.z & 0xffffu) <= rel) +
                           (uint32_t)(20u < nl && (gh.z >> 16) <= rel) + (uint32_t)(24u < nl && (gh.w & 0xffffu) <= rel) +
                           (uint32_t)(28u < nl && (gh.w >> 16) <= rel);
      const uint4 *rp4 = ix.payload + info.y;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      uint32_t add;
      if (rcs == 0u) {
        // four one-word records in one granule
        const uint4 g = __ldg(rp4 + grp);
        add = (uint32_t)(4u * grp + 1u < nl && (g.y & m0) <= rel) + (uint32_t)(4u * grp + 2u < nl && (g.z & m0) <= rel) +
              (uint32_t)(4u * grp + 3u < nl && (g.w & m0) <= rel);
        raw.x = add == 0u ? g.x : add == 1u ? g.y : add == 2u ? g.z : g.w;
      } else {
        // four two-word records in two granules (the head is in the low word: w0 <= 16)
        const uint4 ga = __ldg(rp4 + 2u * grp);
        uint4 gb = make_uint4(0xffffffffu, 0u, 0xffffffffu, 0u);
        if (4u * grp + 2u < nl) gb = __ldg(rp4 + 2u * grp + 1u);
        add = (uint32_t)(4u * grp + 1u < nl && (ga.z & m0) <= rel) + (uint32_t)(4u * grp + 2u < nl && (gb.x & m0) <= rel) +
              (uint32_t)(4u * grp + 3u < nl && (gb.z & m0) <= rel);
        raw.x = add == 0u ? ga.x : add == 1u ? ga.z : add == 2u ? gb.x : gb.z;
        raw.y = add == 0u ? ga.y : add == 1u ? ga.w : add == 2u ? gb.y : gb.w;
      }
      rec = 4u * grp + add;
      DecodeRaw(info, raw, e);
    } else {
      // wide blocks (span >= 2^16 docs or 16-byte records): two levels of strided head loads
      const uint32_t *rp = reinterpret_cast<const uint32_t *>(ix.payload + info.y);
      const uint32_t rw = RecWords(rcs);
      uint32_t grp = 0;
#pragma unroll
      for (uint32_t t = 1; t < 8; t++) {
        const uint32_t mid = 4u * t;
        if (mid < nl && (__ldg(rp + mid * rw) & m0) <= rel) grp = t;   // heads ascend: last true wins
      }
      uint32_t add = 0;
#pragma unroll
      for (uint32_t t = 1; t < 4; t++) {
        const uint32_t mid = 4u * grp + t;
        if (mid < nl && (__ldg(rp + mid * rw) & m0) <= rel) add = t;
      }
      rec = 4u * grp + add;
      DecodeRecord(ix, info, rec, e);
    }
    const int slot = e[0] == x ? 0 : e[1] == x ? 1 : e[2] == x ? 2 : e[3] == x ? 3 : -1;
    *hit = slot >= 0;
    *pos = ((p.first + j) << 7) | (4u * rec + (uint32_t)max(slot, 0));
  }
  // accounting (B_touched): a partner block of which a record was read counts once per
  // unit with its doc-id pack + 16 B of metadata. Candidates ascend across lanes and batches, so a
  // block is new when it differs from the previous lane's and lies past the last one counted.
  if constexpr (ST::kOn) {
    const uint32_t gj = has ? p.first + j : kNoDoc;
    const uint32_t prev = __shfl_up_sync(kFull, gj, 1);
    const bool fresh = has && (lane == 0 || gj != prev) && (st.last_probe == kNoDoc || gj > st.last_probe);
    st.probe_blocks += __popc(__ballot_sync(kFull, fresh));
    st.bytes += __reduce_add_sync(kFull, fresh ? AlgBytes(info.z, false) : 0u);
    const uint32_t mxj = __reduce_max_sync(kFull, has ? gj + 1u : 0u);
    if (mxj && (st.last_probe == kNoDoc || mxj - 1u > st.last_probe)) st.last_probe = mxj - 1u;
  }
  return true;
}

// ---- phrase verification (QueryProcessor::HandleTheFoundDoc + PhraseQueryProcessor2,
// query_processing.h:282-362, 886-895): a document that holds every term is ranked only if the
// terms occur at consecutive positions, in query order. Only existence matters (the score
// ignores the phrase frequency). Runs per intersection hit, one hit per lane.
struct PosRun {
  const void *p;       // the posting's positions, ascending (u16 or u32 entries: DevIndexView::pos16)
  uint32_t n;          // = tf
};
__device__ __forceinline__ uint32_t PosAt(const PosRun &r, uint32_t i, bool p16) {
  return p16 ? (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(r.p) + i)
             : __ldg(reinterpret_cast<const uint32_t *>(r.p) + i);
}
// tf of posting `pos` and the start of its run in positions[]: the block's first position index
// plus the tfs of the postings before it in the block.
__device__ __forceinline__ void LoadTfs4(const uint4 *src, uint32_t tc, uint32_t r, uint32_t t[4]) {
  if (tc == 0) {
    const uint32_t v = __ldg(reinterpret_cast<const unsigned short *>(src) + r);
    t[0] = v & 15u; t[1] = (v >> 4) & 15u; t[2] = (v >> 8) & 15u; t[3] = v >> 12;
  } else if (tc == 1) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src) + r);
    t[0] = v & 255u; t[1] = (v >> 8) & 255u; t[2] = (v >> 16) & 255u; t[3] = v >> 24;
  } else {
    const uint4 v = __ldg(src + r);
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
}
__device__ __forceinline__ uint32_t ByteSum(uint32_t x) { return (uint32_t)__dp4a(x, 0x01010101u, 0u); }
__device__ __forceinline__ uint32_t NibbleSum(uint32_t x) {
  return ByteSum(x & 0x0f0f0f0fu) + ByteSum((x >> 4) & 0x0f0f0f0fu);
}
__device__ __forceinline__ PosRun PositionsOf(const DevIndexView &ix, uint32_t pos) {
  const uint32_t blk = pos >> 7, slot = pos & 127u, rec = slot >> 2, grp = rec >> 2;
  const uint4 info = __ldg(&ix.blk_info[blk]);
  const uint32_t first = __ldg(&ix.blk_pos[blk]);
  uint32_t before = grp ? (uint32_t)__ldg(&ix.grp_pos[(size_t)blk * 8u + grp]) : 0u;
  const uint32_t bits = info.z, tc = ShTcode(bits);
  const uint4 *src = ix.payload + info.y + DocGranules(bits);
  const uint32_t r = rec & 3u, sl = slot & 3u;
  PosRun run;
  if (before != 0xFFFFu && tc == 0u) {
    // the group's four tf records (4 x 4 nibbles) in one 8-byte load; positions before the
    // posting inside its group = sum of the nibbles below it
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src) + grp);
    uint32_t w = v.x, nb = 16u * r + 4u * sl;
    if (nb >= 32u) { before += NibbleSum(w); w = v.y; nb -= 32u; }
    before += NibbleSum(w & ((1u << nb) - 1u));
    run.n = (w >> nb) & 15u;
  } else if (before != 0xFFFFu && tc == 1u) {
    // four records of four bytes: one 16-byte load
    const uint4 v = __ldg(src + grp);
    const uint32_t w = r == 0u ? v.x : r == 1u ? v.y : r == 2u ? v.z : v.w;
    before += (r > 0u ? ByteSum(v.x) : 0u) + (r > 1u ? ByteSum(v.y) : 0u) + (r > 2u ? ByteSum(v.z) : 0u);
    before += ByteSum(w & ((1u << (8u * sl)) - 1u));
    run.n = (w >> (8u * sl)) & 255u;
  } else {
    // 32-bit tfs, or a block with >= 65535 positions (prefixes not stored): walk the tf records
    uint32_t t[4];
    uint32_t r0 = grp << 2;
    if (before == 0xFFFFu) {
      before = 0;
      r0 = 0;
    }
    for (uint32_t q = r0; q < rec; q++) {
      LoadTfs4(src, tc, q, t);
      before += t[0] + t[1] + t[2] + t[3];
    }
    LoadTfs4(src, tc, rec, t);
    before += (sl > 0 ? t[0] : 0u) + (sl > 1 ? t[1] : 0u) + (sl > 2 ? t[2] : 0u);
    run.n = sl == 0 ? t[0] : sl == 1 ? t[1] : sl == 2 ? t[2] : t[3];
  }
  const size_t at = (size_t)first + before;
  run.p = ix.pos16 ? static_cast<const void *>(reinterpret_cast<const unsigned short *>(ix.positions) + at)
                   : static_cast<const void *>(reinterpret_cast<const uint32_t *>(ix.positions) + at);
  return run;
}
// exists p in a with p + 1 in b
__device__ __forceinline__ bool PhraseTwo(const PosRun a, const PosRun b, bool p16) {
  uint32_t i = 0, j = 0;
  while (i < a.n && j < b.n) {
    const uint32_t x = PosAt(a, i, p16) + 1u, y = PosAt(b, j, p16);
    if (x == y) return true;
    if (x < y) i++; else j++;
  }
  return false;
}

// ---- two-term units (the headline path): TwoTermNonPhraseQueryProcessor::Process -------------
// Driver blocks are walked in order; for the smallest unresolved candidate the probe block is
// located, staged once, and every candidate that falls inside it is resolved in the same pass.
// Hits are queued in shared memory and scored 32 at a time (one lane per hit) so the divergent,
// latency-heavy tf / norm gathers and the fp64 divisions stay off the per-block path.
template <bool COLLECT, bool PHRASE, class ST>
__device__ void FlushHits(const DevIndexView &ix, const BatchView &bv, const DevQuery &q, uint32_t qi,
                          const CtaShared *sh, uint32_t a_ws, int nq, int drv, double idf0,
                          double idf1, TopK &top, double &published, bool multi, int lane,
                          ST &st) {
  __syncwarp();
  for (int base = 0; base < nq; base += 32) {
    bool has = base + lane < nq;
    bool keep = true;
    double s = 0.0;
    int doc = 0;
    if (has) {
      HitRec h;
      {
        const uint32_t ha = a_ws + (uint32_t)offsetof(ProbeScratch, hits) + 12u * (uint32_t)(base + lane);
        h.doc = LdsU32(ha); h.pos_a = LdsU32(ha + 4u); h.pos_b = LdsU32(ha + 8u);
      }
      doc = (int)h.doc;
      uint32_t tfa, tfb;
      uint32_t nbyte = 0;
      if constexpr (PHRASE) {   // phrase: query term 0 must be directly followed by term 1
        const PosRun ra = PositionsOf(ix, h.pos_a), rb = PositionsOf(ix, h.pos_b);
        keep = drv == 0 ? PhraseTwo(ra, rb, ix.pos16 != 0u) : PhraseTwo(rb, ra, ix.pos16 != 0u);
        tfa = ra.n;           // a posting's run length IS its tf
        tfb = rb.n;
        WSR_STAT(st.bytes += 4ull * (ra.n + rb.n););
      } else {
        // the norm byte first: its round trip overlaps the two blk_info -> tf word chains
        asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(nbyte) : "l"(ix.norms + h.doc));
        tfa = TfAt(ix, h.pos_a);
        tfb = TfAt(ix, h.pos_b);
      }
      if (keep) {
        if constexpr (PHRASE) nbyte = __ldg(ix.norms + h.doc);
        const double cn = sh->cache[nbyte & 255u];
        // query order: term 0 first (scoring.h:124-145)
        s = __dadd_rn(0.0, TermScore(idf0, drv == 0 ? tfa : tfb, cn));
        s = __dadd_rn(s, TermScore(idf1, drv == 0 ? tfb : tfa, cn));
      }
    }
    has = has && keep;
    if (COLLECT) CollectAppend(bv, q, qi, has, doc, s, lane);
    else OfferToTopK(bv, qi, multi, (int)q.k, has, s, doc, top, published, lane);
  }
  WSR_STAT(st.matches += nq;);
  WSR_STAT(st.bytes += nq;);   // one norm byte per intersection hit
  __syncwarp();
}

// Probes the first n (<= 32) queued survivors and appends the hits to the hit queue.
template <class ST>
__device__ __forceinline__ bool ProbeBatch(const DevIndexView &ix, ProbeList &pb, uint32_t wsa,
                                           int base, int n, int &nq, int lane, ST &st) {
  const bool has = lane < n;
  CandRec c;
  c.doc = 0; c.pos_a = 0;
  if (has) {
    const uint2 v = LdsU64(wsa + (uint32_t)offsetof(ProbeScratch, cand) + 8u * (uint32_t)(base + lane));
    c.doc = v.x; c.pos_a = v.y;
  }
  bool hit;
  uint32_t pos = 0;
  const bool more = ProbeOne(ix, pb, wsa + (uint32_t)offsetof(ProbeScratch, win), has, c.doc, &hit, &pos, lane, st);
  const unsigned m = __ballot_sync(kFull, hit);
  if (m) {
    if (hit) {
      const uint32_t ha = wsa + (uint32_t)offsetof(ProbeScratch, hits) + 12u * (uint32_t)(nq + __popc(m & ((1u << lane) - 1u)));
      StsU32(ha, c.doc);
      StsU32(ha + 4u, c.pos_a);
      StsU32(ha + 8u, pos);
    }
    nq += __popc(m);
  }
  return more;
}

// ---- staged look-ahead of the two-term driver loop (see ProcessTwo) -----------------------------
// the doc-record stream of a block (<= 32 granules of 16 bytes) into the warp's scratch
__device__ __forceinline__ void IssueRecords(const DevIndexView &ix, uint32_t ra, const uint4 info, int lane) {
  if ((uint32_t)lane < DocGranules(info.z)) CpAsync16(ra + 16u * (uint32_t)lane, ix.payload + info.y + lane);
}
// record `rec` of a staged stream (zero for lanes past the block's records, like LoadRecord's callers)
__device__ __forceinline__ uint4 StagedRecord(uint32_t ra, uint32_t bits, uint32_t rec) {
  const uint32_t rc = ShRcode(bits);
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (rec < ((ShN(bits) + 3u) >> 2)) {
    if (rc == 0u) {
      r.x = LdsU32(ra + 4u * rec);
    } else if (rc == 1u) {
      const uint2 v = LdsU64(ra + 8u * rec);
      r.x = v.x; r.y = v.y;
    } else if (rc == 2u) {
      r = LdsU128(ra + 16u * rec);
    } else {
      r.x = LdsU32(ra + 12u * rec); r.y = LdsU32(ra + 12u * rec + 4u); r.z = LdsU32(ra + 12u * rec + 8u);
    }
  }
  return r;
}
// the four filter words of a lane's docs into its private 16-byte slot
__device__ __forceinline__ void IssueFilterWords(const DevIndexView &ix, const ListFilter &lf, uint32_t fa, const uint32_t d[4]) {
  if (lf.words == nullptr) {
    StsU128(fa, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) CpAsync4(fa + 4u * (uint32_t)i, lf.words + ((d[i] - ix.doc_lo) >> lf.shift));
  }
}

template <bool COLLECT, bool PHRASE, class ST>
__device__ void ProcessTwo(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                           uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                           const CtaShared *sh, uint32_t a_sh, uint32_t a_ws, int lane, ST &st) {
  const int drv = (int)q.driver, oth = 1 - drv;
  const uint4 la = __ldg(&ix.lists[q.term[drv]]);
  const uint4 lb = __ldg(&ix.lists[q.term[oth]]);
  const double idf0 = __ldg(&ix.idf[q.term[0]]), idf1 = __ldg(&ix.idf[q.term[1]]);
  const ListFilter flt = FilterOf(ix, q.term[oth]);
  const uint32_t first_a = la.x;
  const bool multi = q.n_units > 1;
  TopK top;
  TopKInit(top);
  double published = 0.0;
  ProbeList pb;
  ProbeInit(pb, ix, lb, lane);
  int nq = 0, nc = 0;
  bool more = true;
  const uint32_t a_fpat = a_sh + (uint32_t)offsetof(CtaShared, fpat);
  const uint32_t a_cand = a_ws + (uint32_t)offsetof(ProbeScratch, cand);

  // Software pipeline over driver blocks, staged through the warp's scratch with cp.async: while
  // block ja is tested and its survivors are probed, the filter words of block ja+1, the doc
  // records of block ja+2 and the blk_info row of block ja+3 are in flight. Each copy group has a
  // whole iteration to land and occupies no register meanwhile.
  const uint32_t a_inf = a_ws + (uint32_t)offsetof(ProbeScratch, inf);
  const uint32_t a_rec = a_ws + (uint32_t)offsetof(ProbeScratch, rec);
  const uint32_t a_fw = a_ws + (uint32_t)offsetof(ProbeScratch, fw) + 16u * (uint32_t)lane;
  const uint32_t a_docs = a_ws + (uint32_t)offsetof(ProbeScratch, docs) + 16u * (uint32_t)lane;
  if (lane < 3 && b0 + (uint32_t)lane < b1) CpAsync16(a_inf + 16u * (uint32_t)lane, &ix.blk_info[first_a + b0 + (uint32_t)lane]);
  CpAsyncCommit();
  CpAsyncWaitAll();
  __syncwarp();
  {
    const uint4 i0 = LdsU128(a_inf);
    IssueRecords(ix, a_rec, i0, lane);
    if (b0 + 1 < b1) IssueRecords(ix, a_rec + 512u, LdsU128(a_inf + 16u), lane);
    CpAsyncCommit();
    CpAsyncWaitAll();
    __syncwarp();
    uint32_t d0[4];
    DecodeRaw(i0, StagedRecord(a_rec, i0.z, (uint32_t)lane), d0);
    StsU128(a_docs, d0[0], d0[1], d0[2], d0[3]);
    IssueFilterWords(ix, flt, a_fw, d0);
    CpAsyncCommit();
  }
  // One loop with ONE call site each for the probe and for the scoring: both are several hundred
  // instructions once inlined, and a second copy for the unit's last few survivors and hits (as in
  // the first version of this loop) doubled the kernel's instruction footprint -- the loops of
  // this kernel live or die by the 32 KB instruction cache. An iteration tests the next driver
  // block when fewer than 32 survivors are queued, probes one batch of survivors when there are
  // 32 (or the blocks are through), and scores the queued hits when there are 32 (or at the end).
  uint32_t ja = b0;
  int cbase = 0;   // survivors [0, cbase) have been probed
  for (;;) {
    if (ja < b1 && more && nc - cbase < 32) {
      if (cbase) {   // move the unprobed survivors (< 32) to the front
        uint2 c = make_uint2(0u, 0u);
        const bool mv = cbase + lane < nc;
        if (mv) c = LdsU64(a_cand + 8u * (uint32_t)(cbase + lane));
        __syncwarp();
        if (mv) StsU64(a_cand + 8u * (uint32_t)lane, c.x, c.y);
        nc -= cbase;
        cbase = 0;
        __syncwarp();
      }
      const uint32_t t = ja - b0;
      CpAsyncWaitAll();
      __syncwarp();
      const uint4 info_cur = LdsU128(a_inf + 16u * (t & 3u));
      const uint32_t na = ShN(info_cur.z);
      WSR_STAT(st.decoded += na;);
      WSR_STAT(st.bytes += AlgBytes(info_cur.z, false););
      uint32_t d[4];
      {
        const uint4 dv = LdsU128(a_docs);
        d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
      }
      // ---- stage: docs + filter words of block ja+1, records of ja+2, blk_info of ja+3
      if (ja + 1 < b1) {
        const uint4 info_nxt = LdsU128(a_inf + 16u * ((t + 1u) & 3u));
        uint32_t dn[4];
        DecodeRaw(info_nxt, StagedRecord(a_rec + 512u * ((t + 1u) & 1u), info_nxt.z, (uint32_t)lane), dn);
        StsU128(a_docs, dn[0], dn[1], dn[2], dn[3]);
        IssueFilterWords(ix, flt, a_fw + 512u * ((t + 1u) & 1u), dn);
        if (ja + 2 < b1) {
          IssueRecords(ix, a_rec + 512u * (t & 1u), LdsU128(a_inf + 16u * ((t + 2u) & 3u)), lane);
          if (ja + 3 < b1 && lane == 0) CpAsync16(a_inf + 16u * ((t + 3u) & 3u), &ix.blk_info[first_a + ja + 3]);
        }
      }
      CpAsyncCommit();
      // ---- Bloom pre-test of block ja, then compaction in (lane, slot) = doc order
      uint32_t fw[4];
      {
        const uint4 fv = LdsU128(a_fw + 512u * (t & 1u));
        fw[0] = fv.x; fw[1] = fv.y; fw[2] = fv.z; fw[3] = fv.w;
      }
      bool pass[4];
      unsigned bm[4];
#pragma unroll
      for (int i = 0; i < 4; i++) pass[i] = FilterTestS(a_fpat, fw[i], d[i]);
      if (na != 128u) {   // a list's last block: padded slots repeat the last doc
#pragma unroll
        for (int i = 0; i < 4; i++) pass[i] = pass[i] && 4u * lane + i < na;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) bm[i] = __ballot_sync(kFull, pass[i]);
      const unsigned lt = (1u << lane) - 1u;
      uint32_t at = a_cand + 8u * (uint32_t)(nc + __popc(bm[0] & lt) + __popc(bm[1] & lt) + __popc(bm[2] & lt) + __popc(bm[3] & lt));
      const uint32_t ga = ((first_a + ja) << 7) | (4u * (uint32_t)lane);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (pass[i]) {
          StsU64(at, d[i], ga | (uint32_t)i);   // CandRec {doc, pos_a}
          at += 8u;
        }
      }
      nc += __popc(bm[0]) + __popc(bm[1]) + __popc(bm[2]) + __popc(bm[3]);
      __syncwarp();
      ja++;
    }
    const bool through = !(ja < b1 && more);   // no driver block left to test
    // ---- exact probe of 32 survivors (of the last few once the blocks are through)
    if (more && (nc - cbase >= 32 || (through && nc > cbase))) {
      const int n = min(32, nc - cbase);
      more = ProbeBatch(ix, pb, a_ws, cbase, n, nq, lane, st);
      cbase += n;
    }
    const bool fin = through && (!more || nc == cbase);
    if (nq >= 32 || (fin && nq)) {
      FlushHits<COLLECT, PHRASE>(ix, bv, q, qi, sh, a_ws, nq, drv, idf0, idf1, top, published, multi, lane, st);
      nq = 0;
    }
    if (fin) break;
  }
  // nothing of this unit may still be in flight when the next unit reuses the staging slots
  CpAsyncWaitAll();
  __syncwarp();
  if (!COLLECT) EmitTopK(bv, q, local, top, lane);
}

// ---- two-term units, merge mode: lists of similar length ---------------------------------------
// TwoTermNonPhraseQueryProcessor::Process (query_processing.h:656-677) advances both iterators in
// lock step when neither list is much longer than the other. The probe path above spends a filter
// word, a skip-metadata lookup and three dependent loads per survivor on such queries, although
// the partner's blocks under a driver block are no more bytes than its filter words. Here both
// lists are streamed once, coalesced: driver blocks are decoded into a set in shared memory (byte
// map hashed by the low doc-id bits + a ring of their doc ids and tfs), every partner block that
// overlaps them is decoded and its postings tested against the map; the few that pass are queued
// with their tf and, 32 at a time, located in the ring by binary search (which settles hash
// aliases and yields the driver posting's tf), scored and offered to the top-k. No filter words,
// no record search, no tf gathers: the only random access left is the norm byte of a hit.
// 128-bit records (blocks spanning >= 2^16 docs or with wide deltas) are rare: out of line, so
// that the merge loop's instruction footprint stays small (the loops of this kernel live or die
// by the 32 KB instruction cache).
__device__ __noinline__ void DecodeDocsWide(const DevIndexView &ix, const uint4 info, int lane, uint32_t d[4]) {
  DecodeDocs(ix, info, lane, d);
}
// rcode <= 1 only: [f:w0][d1:b][d2:b][d3:b] in 64 bits
__device__ __forceinline__ void DecodeRaw64(const uint4 info, const uint2 raw, uint32_t d[4]) {
  const uint32_t bits = info.z;
  const uint32_t w0 = ShW0(bits), b = ShB(bits);
  const uint32_t m0 = w0 >= 32u ? 0xffffffffu : ((1u << w0) - 1u);
  const uint32_t mb = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
  unsigned long long x = ((unsigned long long)raw.y << 32) | raw.x;
  const uint32_t f = (uint32_t)x & m0;
  x >>= w0;
  const uint32_t d1 = (uint32_t)x & mb;
  x >>= b;
  const uint32_t d2 = (uint32_t)x & mb;
  x >>= b;
  const uint32_t d3 = (uint32_t)x & mb;
  d[0] = info.x + f;
  d[1] = d[0] + d1;
  d[2] = d[1] + d2;
  d[3] = d[2] + d3;
}
__device__ __forceinline__ uint2 LoadRec2(const DevIndexView &ix, const uint4 info, int lane) {
  // the lane's doc record in the 32/64-bit formats; zeros past the block's records and for
  // 128-bit records (those are decoded on demand)
  const uint32_t bits = info.z, rc = ShRcode(bits), nl = (ShN(bits) + 3u) >> 2;
  uint2 r = make_uint2(0u, 0u);
  if ((uint32_t)lane < nl) {
    const uint4 *src = ix.payload + info.y;
    if (rc == 0u) r.x = __ldg(reinterpret_cast<const uint32_t *>(src) + lane);
    else if (rc == 1u) r = __ldg(reinterpret_cast<const uint2 *>(src) + lane);
  }
  return r;
}
__device__ __forceinline__ uint32_t LoadTfWord(const DevIndexView &ix, const uint4 info, int lane) {
  // the lane's tf record in the 4/8-bit formats (32-bit tfs are read on demand)
  const uint32_t bits = info.z, tc = ShTcode(bits), nl = (ShN(bits) + 3u) >> 2;
  uint32_t v = 0u;
  if ((uint32_t)lane < nl) {
    const uint4 *src = ix.payload + info.y + DocGranules(bits);
    if (tc == 0u) v = __ldg(reinterpret_cast<const unsigned short *>(src) + lane);
    else if (tc == 1u) v = __ldg(reinterpret_cast<const uint32_t *>(src) + lane);
  }
  return v;
}

template <class ST>
__device__ void ProcessTwoMerge(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                                uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                                const CtaShared *sh, MergeScratch *ms, int lane, ST &st) {
  const int drv = (int)q.driver, oth = 1 - drv;
  const uint4 la = __ldg(&ix.lists[q.term[drv]]);
  const uint4 lb = __ldg(&ix.lists[q.term[oth]]);
  const double idf0 = __ldg(&ix.idf[q.term[0]]), idf1 = __ldg(&ix.idf[q.term[1]]);
  const uint32_t first_a = la.x, first_b = lb.x, nb = lb.y;
  const bool multi = q.n_units > 1;
  constexpr uint32_t kMapMask = (uint32_t)kMapBytes - 1u;
  constexpr uint32_t kSlotMask = (uint32_t)kRingBlocks - 1u;
  TopK top;
  TopKInit(top);
  double published = 0.0;
  // the map starts clean and every unit leaves it clean (see the end of this function)
  uint4 infoA = __ldg(&ix.blk_info[first_a + b0]);
  // the partner's first block that reaches the unit's doc range (docs of block b0 are >= its base)
  uint32_t jb;
  {
    ProbeList pb;
    ProbeInit(pb, ix, lb, lane);
    jb = ProbeFind(pb, infoA.x, lane);
  }
  if (jb == kNoDoc) {
    EmitTopK(bv, q, local, top, lane);
    return;
  }
  // software pipelines over both lists: a block's info arrives two blocks ahead of its use, the
  // lane's records one block ahead
  uint4 infoA_nxt = b0 + 1u < b1 ? __ldg(&ix.blk_info[first_a + b0 + 1u]) : infoA;
  uint2 rawA = LoadRec2(ix, infoA, lane);
  uint32_t tfwA = LoadTfWord(ix, infoA, lane);
  uint4 infoB = __ldg(&ix.blk_info[first_b + jb]);
  uint4 infoB_nxt = jb + 1u < nb ? __ldg(&ix.blk_info[first_b + jb + 1u]) : infoB;
  uint2 rawB = LoadRec2(ix, infoB, lane);
  uint32_t tfwB = LoadTfWord(ix, infoB, lane);
  uint32_t ja = b0, jr = b0;   // driver blocks [jr, ja) are in the map; [max(b0, ja-4), ja) in the ring
  uint32_t q_lo = b0;          // oldest driver block a queued partner posting may match
  int nq = 0, nh = 0;          // queued partner postings / verified matches
  const unsigned lt = (1u << lane) - 1u;
  __syncwarp();
  bool redo = false;           // the partner block in hand reaches past a full ring: go round again
  uint2 rawB_nxt = make_uint2(0u, 0u);
  uint32_t tfwB_nxt = 0u;
  uint4 infoB_nxt2 = infoB_nxt;
  for (;;) {
    uint32_t dB[4];
    const uint32_t nB = ShN(infoB.z);
    if (!redo) {
      // request block jb+1's records and block jb+2's info before working on block jb
      if (jb + 1u < nb) {
        rawB_nxt = LoadRec2(ix, infoB_nxt, lane);
        tfwB_nxt = LoadTfWord(ix, infoB_nxt, lane);
        if (jb + 2u < nb) infoB_nxt2 = __ldg(&ix.blk_info[first_b + jb + 2u]);
      }
      WSR_STAT(st.decoded += nB; st.bytes += AlgBytes(infoB.z, true););
    }
    if (ShRcode(infoB.z) >= 2u) DecodeDocsWide(ix, infoB, lane, dB);
    else DecodeRaw64(infoB, rawB, dB);
    const uint32_t lastB = __shfl_sync(kFull, dB[3], (int)((nB + 3u) >> 2) - 1);
    // ---- driver blocks that reach into this partner block join the set. Ring slots reused here
    // held blocks below jr - (those were retired) and no queued posting can match them: the
    // queue is settled at the end of every round in which a block it may match retired.
#pragma unroll 1
    while (ja < b1 && (infoA.x < lastB || ja == 0u) && ja - jr < (uint32_t)kRingBlocks) {
      const uint32_t nA = ShN(infoA.z), nlA = (nA + 3u) >> 2, tcA = ShTcode(infoA.z);
      uint32_t dA[4];
      if (ShRcode(infoA.z) >= 2u) DecodeDocsWide(ix, infoA, lane, dA);
      else DecodeRaw64(infoA, rawA, dA);
      const uint32_t lastA = __shfl_sync(kFull, dA[3], (int)nlA - 1);
      if ((uint32_t)lane >= nlA) dA[0] = dA[1] = dA[2] = dA[3] = lastA;   // keeps the ring sorted
      uint32_t tfp;   // the record's four tfs, one byte each, 255 = look the exact value up
      if (tcA == 0u) tfp = (tfwA & 0xfu) | ((tfwA & 0xf0u) << 4) | ((tfwA & 0xf00u) << 8) | ((tfwA & 0xf000u) << 12);
      else if (tcA == 1u) tfp = tfwA;
      else tfp = 0xffffffffu;
      WSR_STAT(st.decoded += nA; st.bytes += AlgBytes(infoA.z, true););
      const uint32_t slot = (ja & kSlotMask) << 5;
      reinterpret_cast<uint4 *>(ms->rdoc)[slot + lane] = make_uint4(dA[0], dA[1], dA[2], dA[3]);
      reinterpret_cast<uint32_t *>(ms->rtf)[slot + lane] = tfp;
#pragma unroll
      for (int i = 0; i < 4; i++) ms->map[dA[i] & kMapMask] = 1;
      ja++;
      infoA = infoA_nxt;
      if (ja < b1) {
        rawA = LoadRec2(ix, infoA, lane);
        tfwA = LoadTfWord(ix, infoA, lane);
        if (ja + 1u < b1) infoA_nxt = __ldg(&ix.blk_info[first_a + ja + 1u]);
      }
    }
    __syncwarp();
    // ---- the partner block's postings against the map
    {
      bool sv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) sv[i] = 4u * lane + i < nB && ms->map[dB[i] & kMapMask] != 0;
      if (__any_sync(kFull, sv[0] || sv[1] || sv[2] || sv[3])) {
        if (nq == 0) q_lo = jr;
        const uint32_t tcB = ShTcode(infoB.z);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const unsigned bm = __ballot_sync(kFull, sv[i]);
          if (sv[i]) {
            const uint32_t tf = tcB == 0u ? (tfwB >> (4 * i)) & 15u
                                : tcB == 1u ? (tfwB >> (8 * i)) & 255u
                                            : TfAt(ix, ((first_b + jb) << 7) | (4u * lane + i));
            ms->surv[nq + __popc(bm & lt)] = make_uint2(dB[i], tf);
          }
          nq += __popc(bm);
        }
      }
    }
    // a full ring with more driver blocks reaching into this partner block: everything in the
    // set lies below the next driver block, it all retires and the block goes round again
    redo = ja < b1 && infoA.x < lastB;
    // ---- driver blocks that end at or before this partner block's last doc leave the set
    uint32_t nr = jr;
#pragma unroll 1
    while (nr < ja && ms->rdoc[((nr & kSlotMask) << 7) + 127u] <= lastB) nr++;
    const bool last = !redo && ((ja == b1 && nr == ja) || jb + 1u >= nb);   // the unit's last round
    if (nr != jr || last) {
      // bytes may be shared with postings that stay (hash aliases): clear the leavers', then set
      // the stayers' again. The last round clears everything: the next unit finds a clean map.
      const uint32_t clr_end = last ? ja : nr;
#pragma unroll 1
      for (uint32_t t = jr; t < clr_end; t++) {
        const uint4 d4 = reinterpret_cast<const uint4 *>(ms->rdoc)[((t & kSlotMask) << 5) + lane];
        ms->map[d4.x & kMapMask] = 0; ms->map[d4.y & kMapMask] = 0;
        ms->map[d4.z & kMapMask] = 0; ms->map[d4.w & kMapMask] = 0;
      }
      __syncwarp();
#pragma unroll 1
      for (uint32_t t = clr_end; t < ja; t++) {
        const uint4 d4 = reinterpret_cast<const uint4 *>(ms->rdoc)[((t & kSlotMask) << 5) + lane];
        ms->map[d4.x & kMapMask] = 1; ms->map[d4.y & kMapMask] = 1;
        ms->map[d4.z & kMapMask] = 1; ms->map[d4.w & kMapMask] = 1;
      }
      jr = nr;
    }
    __syncwarp();
    // ---- settle the queue against the ring when it is long enough, or when a block it may match
    // has left the set (its ring slot may be reused from the next round on): binary search for the
    // driver posting, which also drops hash aliases and yields the driver-side tf
    if (nq >= 32 || (nq && q_lo < jr) || last) {
      const uint32_t lo_blk = ja >= b0 + (uint32_t)kRingBlocks ? ja - (uint32_t)kRingBlocks : b0;
      const uint32_t glo = lo_blk << 7, ghi = ja << 7;   // driver posting indices the ring holds
      int base = 0;
#pragma unroll 1
      do {
        bool has = base + lane < nq;
        uint32_t x = 0u, tfb = 0u, g = glo;
        if (has) {
          const uint2 sv = ms->surv[base + lane];
          x = sv.x;
          tfb = sv.y;
        }
#pragma unroll
        for (uint32_t s = (uint32_t)kRingBlocks * 64u; s; s >>= 1) {   // lower bound of x in the ring
          const uint32_t t = g + s;
          if (t <= ghi && ms->rdoc[(t - 1u) & (uint32_t)kRingMask] < x) g = t;
        }
        has = has && g < ghi && ms->rdoc[g & (uint32_t)kRingMask] == x;
        const unsigned hm = __ballot_sync(kFull, has);
        if (has) {
          uint32_t tfa = ms->rtf[g & (uint32_t)kRingMask];
          if (tfa == 255u) tfa = TfAt(ix, ((first_a + (g >> 7)) << 7) | (g & 127u));
          const int at = nh + __popc(hm & lt);
          ms->hdoc[at] = x;
          ms->htfa[at] = tfa;
          ms->htfb[at] = tfb;
        }
        nh += __popc(hm);
        __syncwarp();
        // ---- score 32 verified matches at a time (the only random access: the norm byte); the
        // unit's last pass scores whatever is left
        const bool final_pass = last && base + 32 >= nq;
#pragma unroll 1
        while (nh >= 32 || (final_pass && nh > 0)) {
          const int n_now = min(nh, 32);
          const bool on = lane < n_now;
          double sc = 0.0;
          int doc = 0;
          if (on) {
            doc = (int)ms->hdoc[lane];
            const uint32_t ta = ms->htfa[lane], tb = ms->htfb[lane];
            const double cn = sh->cache[__ldg(ix.norms + doc)];
            // query order: term 0 first (scoring.h:124-145)
            sc = __dadd_rn(0.0, TermScore(idf0, drv == 0 ? ta : tb, cn));
            sc = __dadd_rn(sc, TermScore(idf1, drv == 0 ? tb : ta, cn));
          }
          WSR_STAT(st.matches += n_now; st.bytes += n_now;);   // one norm byte per hit
          OfferToTopK(bv, qi, multi, (int)q.k, on, sc, doc, top, published, lane);
          // the (at most 31) matches behind the scored ones move to the front
          const int rest = nh - n_now;
          uint32_t m0 = 0, m1 = 0, m2 = 0;
          if (lane < rest) { m0 = ms->hdoc[32 + lane]; m1 = ms->htfa[32 + lane]; m2 = ms->htfb[32 + lane]; }
          __syncwarp();
          if (lane < rest) { ms->hdoc[lane] = m0; ms->htfa[lane] = m1; ms->htfb[lane] = m2; }
          nh = rest;
          __syncwarp();
        }
        base += 32;
      } while (base < nq);
      nq = 0;
    }
    if (last) break;
    if (redo) continue;
    // ---- next partner block: its records were requested at the start of this round
    jb++;
    infoB = infoB_nxt;
    infoB_nxt = infoB_nxt2;
    rawB = rawB_nxt;
    tfwB = tfwB_nxt;
  }
  EmitTopK(bv, q, local, top, lane);
}

// ---- 3..8-term units: QueryProcessor::ProcessMultipleTerms, query_processing.h:710-728 -------
// The shortest list drives. A driver posting survives when it passes the Bloom filters of ALL
// other lists; survivors are compacted in doc order and, 32 at a time (one per lane), probed
// exactly list by list in query order — a candidate that misses a list is dropped before the
// next one, which is the reference's "break on first mismatch". Matches are phrase-checked
// (when asked) and scored in query order. The walk state of the other lists lives in shared
// memory between batches, so the kernel fits 64 registers.
struct ProbeState {          // a ProbeList parked in shared memory
  uint32_t first, nb, wbase, flt_word, flt_shift, pad[3];
  uint32_t wl[32];
};
struct __align__(16) MultiScratch {
  ProbeScratch ps;                       // win + cand (hits[] unused)
  ProbeState list[WSR_MAX_TERMS];
};

__device__ __forceinline__ ProbeList Unpark(const DevIndexView &ix, const ProbeState &st, int lane) {
  ProbeList p;
  p.first = st.first; p.nb = st.nb; p.wbase = st.wbase;
  p.last = ix.blk_last + st.first;
  p.wl = st.wl[lane];
  return p;
}
__device__ __forceinline__ void Park(ProbeState &st, const ProbeList &p, int lane) {
  st.wbase = p.wbase;
  st.wl[lane] = p.wl;
}

// Probes up to 32 queued survivors (one per lane) against every non-driver list, verifies the
// phrase and scores the matches. Returns false when some list is exhausted (the unit can stop).
template <bool COLLECT, class ST>
__device__ bool MultiBatch(const DevIndexView &ix, const BatchView &bv, const DevQuery &q, uint32_t qi,
                           const CtaShared *sh, MultiScratch *ws, int base, int n, TopK &top,
                           double &published, bool multi, int lane, ST &st) {
  const int m = (int)q.n_terms, drv = (int)q.driver;
  bool has = lane < n;
  CandRec c;
  c.doc = 0; c.pos_a = 0;
  if (has) c = ws->ps.cand[base + lane];
  uint32_t pos[WSR_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < WSR_MAX_TERMS; t++) pos[t] = c.pos_a;
  bool more = true;
#pragma unroll
  for (int t = 0; t < WSR_MAX_TERMS; t++) {
    if (t >= m || t == drv) continue;
    if (!__any_sync(kFull, has)) break;          // every candidate already missed a list
    ProbeList pl = Unpark(ix, ws->list[t], lane);
    bool hit = false;
    uint32_t pb = 0;
    const bool some = ProbeOne(ix, pl, SmemAddr(ws->ps.win), has, c.doc, &hit, &pb, lane, st);
    __syncwarp();
    Park(ws->list[t], pl, lane);
    __syncwarp();
    if (!some) { more = false; has = false; break; }
    has = has && hit;
    pos[t] = pb;
  }
  double s = 0.0;
  if (has) {
    uint32_t tf[WSR_MAX_TERMS];
    bool keep = true;
    if (q.flags & 1u) {
      // positions of every term's posting in this doc: p in term 0, p + t in term t for all t
      PosRun run[WSR_MAX_TERMS];
#pragma unroll
      for (int t = 0; t < WSR_MAX_TERMS; t++) {
        run[t].p = nullptr; run[t].n = 0;
        if (t < m) {
          run[t] = PositionsOf(ix, pos[t]);
          tf[t] = run[t].n;
          WSR_STAT(st.bytes += 4ull * run[t].n;);
        }
      }
      uint32_t ptr[WSR_MAX_TERMS];
#pragma unroll
      for (int t = 0; t < WSR_MAX_TERMS; t++) ptr[t] = 0;
      bool found = false;
      const bool p16 = ix.pos16 != 0u;
      for (uint32_t a0 = 0; a0 < run[0].n && !found; a0++) {
        const uint32_t p0 = PosAt(run[0], a0, p16);
        bool ok = true;
#pragma unroll
        for (int t = 1; t < WSR_MAX_TERMS; t++) {
          if (t < m && ok) {
            const uint32_t want = p0 + (uint32_t)t;
            while (ptr[t] < run[t].n && PosAt(run[t], ptr[t], p16) < want) ptr[t]++;
            ok = ptr[t] < run[t].n && PosAt(run[t], ptr[t], p16) == want;
          }
        }
        found = ok;
      }
      keep = found;
    } else {
#pragma unroll
      for (int t = 0; t < WSR_MAX_TERMS; t++)
        if (t < m) tf[t] = TfAt(ix, pos[t]);
    }
    has = keep;
    if (keep) {
      // query order, fp64, one rounding per operation (scoring.h:124-145)
      const double cn = sh->cache[__ldg(ix.norms + c.doc)];
#pragma unroll
      for (int t = 0; t < WSR_MAX_TERMS; t++)
        if (t < m) s = __dadd_rn(s, TermScore(__ldg(&ix.idf[q.term[t]]), tf[t], cn));
    }
  }
  const unsigned hm = __ballot_sync(kFull, has);
  if (hm) {
    WSR_STAT(st.matches += __popc(hm););
    WSR_STAT(st.bytes += __popc(hm););   // one norm byte per hit
    if (COLLECT) CollectAppend(bv, q, qi, has, (int)c.doc, s, lane);
    else OfferToTopK(bv, qi, multi, (int)q.k, has, s, (int)c.doc, top, published, lane);
  }
  return more;
}

template <bool COLLECT, class ST>
__device__ void ProcessMulti(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                             uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                             const CtaShared *sh, MultiScratch *ws, int lane, ST &st) {
  const int m = (int)q.n_terms;
  const int drv = (int)q.driver;
  const bool multi = q.n_units > 1;
  uint32_t first_a = 0;
  __syncwarp();
  {
    // Unit prologue: the list records, filter descriptors and blk_last windows of all terms are
    // requested in two batches of independent loads (a short driver list makes the whole unit a
    // few dependent round trips long, so every one saved counts).
    uint4 li[WSR_MAX_TERMS];
    uint2 lf[WSR_MAX_TERMS];
#pragma unroll
    for (int t = 0; t < WSR_MAX_TERMS; t++) {
      li[t] = make_uint4(0u, 0u, 0u, 0u);
      lf[t] = make_uint2(0u, 0xffffffffu);
      if (t < m) {
        li[t] = __ldg(&ix.lists[q.term[t]]);
        lf[t] = __ldg(&ix.list_flt[q.term[t]]);
      }
    }
    uint32_t wl[WSR_MAX_TERMS];
#pragma unroll
    for (int t = 0; t < WSR_MAX_TERMS; t++)
      wl[t] = (t < m && t != drv && (uint32_t)lane < li[t].y) ? __ldg(ix.blk_last + li[t].x + lane) : kNoDoc;
#pragma unroll
    for (int t = 0; t < WSR_MAX_TERMS; t++) {
      if (t >= m) continue;
      if (t == drv) { first_a = li[t].x; continue; }
      ProbeState &ps = ws->list[t];
      if (lane == 0) {
        ps.first = li[t].x; ps.nb = li[t].y; ps.wbase = 0;
        ps.flt_word = lf[t].x; ps.flt_shift = lf[t].y;     // shift 0xFFFFFFFF: no filter
      }
      ps.wl[lane] = wl[t];
    }
  }
  __syncwarp();
  TopK top;
  TopKInit(top);
  double published = 0.0;
  int nc = 0;
  bool more = true;

  uint4 cur = __ldg(&ix.blk_info[first_a + b0]);
  for (uint32_t ja = b0; ja < b1 && more; ja++) {
    const uint4 nxt = ja + 1 < b1 ? __ldg(&ix.blk_info[first_a + ja + 1]) : cur;
    uint32_t d[4];
    DecodeDocs(ix, cur, lane, d);
    const uint32_t na = ShN(cur.z);
    WSR_STAT(st.decoded += na;);
    WSR_STAT(st.bytes += AlgBytes(cur.z, false););
    bool pass[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pass[i] = 4u * lane + i < na;   // padded slots repeat the last doc
    // Bloom filters of every other list; a term is skipped once nothing in the warp is alive
    for (int t = 0; t < m; t++) {
      if (t == drv) continue;
      if (!__any_sync(kFull, pass[0] || pass[1] || pass[2] || pass[3])) break;
      const uint32_t shift = ws->list[t].flt_shift;
      if (shift == 0xffffffffu) continue;
      const uint32_t *words = ix.filters + ws->list[t].flt_word;
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = pass[i] ? __ldg(words + ((d[i] - ix.doc_lo) >> (shift & 31u))) : 0u;
#pragma unroll
      for (int i = 0; i < 4; i++) pass[i] = pass[i] && FilterTest(sh, w[i], d[i]);
    }
    unsigned bm[4];
#pragma unroll
    for (int i = 0; i < 4; i++) bm[i] = __ballot_sync(kFull, pass[i]);
    if (bm[0] | bm[1] | bm[2] | bm[3]) {
      const unsigned lt = (1u << lane) - 1u;
      int at = nc + __popc(bm[0] & lt) + __popc(bm[1] & lt) + __popc(bm[2] & lt) + __popc(bm[3] & lt);
      const uint32_t ga = (first_a + ja) << 7;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (pass[i]) {
          CandRec c;
          c.doc = d[i];
          c.pos_a = ga | (4u * lane + i);
          ws->ps.cand[at++] = c;
        }
      }
      nc += __popc(bm[0]) + __popc(bm[1]) + __popc(bm[2]) + __popc(bm[3]);
      __syncwarp();
      int base = 0;
      for (; nc - base >= 32 && more; base += 32)
        more = MultiBatch<COLLECT>(ix, bv, q, qi, sh, ws, base, 32, top, published, multi, lane, st);
      if (base) {   // move the leftover (< 32) to the front
        CandRec c;
        c.doc = 0; c.pos_a = 0;
        const bool mv = base + lane < nc;
        if (mv) c = ws->ps.cand[base + lane];
        __syncwarp();
        if (mv) ws->ps.cand[lane] = c;
        nc -= base;
        __syncwarp();
      }
    }
    cur = nxt;
  }
  if (nc && more) MultiBatch<COLLECT>(ix, bv, q, qi, sh, ws, 0, nc, top, published, multi, lane, st);
  if (!COLLECT) EmitTopK(bv, q, local, top, lane);
}

template <int CLASS> struct ScratchOf { typedef MultiScratch type; };
template <> struct ScratchOf<kClassOne> { typedef NoScratch type; };
template <> struct ScratchOf<kClassTwo> { typedef ProbeScratch type; };

// Persistent search kernel of one query class: warps drain the class's unit queue.
// The per-warp scratch is dynamic shared memory (the two-term class needs more than the 48 KB a
// kernel may declare statically).
extern __shared__ __align__(16) unsigned char g_dyn_smem[];

// MERGE: the two-term class runs as two kernels over the same unit queue, one per path, so that
// each kernel's loops fit the instruction cache (one kernel holding both paths ran at 20 % issue
// utilisation, stalled on instruction fetch). The unit -> query map carries the unit's path in its
// top bit; a kernel skips the other path's units.
constexpr uint32_t kUnitMergeBit = 0x80000000u;
template <int CLASS, bool MERGE> struct ScratchOfKernel { typedef typename ScratchOf<CLASS>::type type; };
template <> struct ScratchOfKernel<kClassTwo, false> { typedef ProbeScratch type; };
template <> struct ScratchOfKernel<kClassTwo, true> { typedef MergeScratch type; };

template <int CLASS, bool STATS, bool MERGE>
#ifndef WSR_TWO_CTAS
#define WSR_TWO_CTAS 4   // resident CTAs per SM the two-term probe kernel is compiled for (64 registers)
#endif
__global__ void __launch_bounds__(kThreadsPerCta, CLASS == kClassTwo ? (MERGE ? 3 : WSR_TWO_CTAS) : CLASS == kClassMany ? 3 : CLASS == kClassOne ? 8 : 1)
SearchKernel(const DevIndexView ix, const BatchView bv) {
  __shared__ CtaShared sh;
  typedef typename ScratchOfKernel<CLASS, MERGE>::type Scratch;
  // static shared memory where it fits (ptxas then addresses it without generic-pointer
  // conversions); only the merge kernel's scratch needs the dynamic window
  constexpr bool kDynamic = sizeof(Scratch) * kWarpsPerCta > 40 * 1024;
  __shared__ Scratch static_scratch[kDynamic ? 1 : kWarpsPerCta];
  Scratch *scratch = kDynamic ? reinterpret_cast<Scratch *>(g_dyn_smem) : static_scratch;
  // thread id through a volatile asm: ptxas otherwise re-reads %tid and re-derives the warp's
  // scratch address in front of every shared-memory access (8-18 % of all issued instructions
  // in the round-2 profiles of the two-term kernels)
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  for (uint32_t i = tid; i < 256u; i += kThreadsPerCta) {
    const double c = ix.cache[i];
    sh.cache[i] = c;
    sh.cache32[i] = __double2float_rd(c);   // smaller denominator => larger (safe) bound
  }
  if constexpr (CLASS != kClassOne) {
    for (uint32_t i = tid; i < (uint32_t)kFilterPatterns; i += kThreadsPerCta) sh.fpat[i] = FilterPattern(i);
  }
  if constexpr (MERGE) {   // every merge unit starts from, and leaves behind, an all-zero map
    uint4 *m4 = reinterpret_cast<uint4 *>(g_dyn_smem);
    for (uint32_t i = tid; i < sizeof(Scratch) * kWarpsPerCta / 16; i += kThreadsPerCta)
      m4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int lane = (int)(tid & 31u);
  auto *ws = &scratch[tid >> 5];
  (void)ws;
  // shared-window addresses of the CTA's tables and the warp's scratch, made opaque so that they
  // live in two registers instead of being re-derived (S2R + MOV + LEA) at every use
  uint32_t a_sh = SmemAddr(&sh), a_ws = SmemAddr(ws);
  asm volatile("" : "+r"(a_sh), "+r"(a_ws));
  const uint32_t n_units = bv.class_units[CLASS];
  UnitStatsT<STATS> st = {0ull, 0ull, 0ull, 0ull, kNoDoc};
  unsigned long long units = 0;
  for (;;) {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(&bv.counters->next_unit[MERGE ? 4 : CLASS], 1u);
    u = __shfl_sync(kFull, u, 0);
    if (u >= n_units) break;
    uint32_t qi = __ldg(&bv.unit_query[bv.class_unit_base[CLASS] + u]);
    if constexpr (CLASS == kClassTwo) {
      if (((qi & kUnitMergeBit) != 0u) != MERGE) continue;   // the other kernel's unit
      qi &= ~kUnitMergeBit;
    }
    const DevQuery q = bv.queries[qi];
    st.last_probe = kNoDoc;
    const uint32_t local = u - q.unit_begin;
    // driver list block range of this unit
    const uint4 li = __ldg(&ix.lists[q.term[q.driver]]);
    // the single-term fast path takes the whole list in one unit
    const uint32_t b0 = CLASS == kClassOne ? 0u : local * q.unit_blocks;
    const uint32_t b1 = CLASS == kClassOne ? li.y : min(b0 + q.unit_blocks, li.y);
    if constexpr (CLASS == kClassOne) {
      ProcessOneTerm<false>(ix, bv, q, qi, local, b0, b1, &sh, lane, st);
    } else if constexpr (CLASS == kClassTwo) {
      if constexpr (MERGE) ProcessTwoMerge(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
      else if (q.flags & kQueryPhrase) ProcessTwo<false, true>(ix, bv, q, qi, local, b0, b1, &sh, a_sh, a_ws, lane, st);
      else ProcessTwo<false, false>(ix, bv, q, qi, local, b0, b1, &sh, a_sh, a_ws, lane, st);
    } else if constexpr (CLASS == kClassMany) {
      ProcessMulti<false>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    } else {
      if (q.n_terms == 1) ProcessOneTerm<true>(ix, bv, q, qi, local, b0, b1, &sh, lane, st);
      else if (q.n_terms == 2 && (q.flags & kQueryPhrase)) ProcessTwo<true, true>(ix, bv, q, qi, local, b0, b1, &sh, a_sh, a_ws, lane, st);
      else if (q.n_terms == 2) ProcessTwo<true, false>(ix, bv, q, qi, local, b0, b1, &sh, a_sh, a_ws, lane, st);
      else ProcessMulti<true>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    }
    units++;
  }
  if (lane == 0 && units) {
    if constexpr (STATS) {
      atomicAdd(&bv.counters->decoded_postings, st.decoded);
      atomicAdd(&bv.counters->touched_bytes, st.bytes);
      atomicAdd(&bv.counters->matches, st.matches);
      atomicAdd(&bv.counters->probe_blocks, st.probe_blocks);
    }
    atomicAdd(&bv.counters->units, units);
  }
}

// unit -> query map of a planned batch: the persistent kernels take a unit number from their
// class queue and need its query; one load here replaces a 4-level search over unit_begin.
__global__ void UnitMapKernel(const BatchView bv, uint32_t *__restrict__ unit_query, uint32_t n_planned) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_planned) return;
  const DevQuery q = bv.queries[i];
  uint32_t c = 0;
  while (c < 3 && i >= bv.class_begin[c + 1]) c++;
  uint32_t *dst = unit_query + bv.class_unit_base[c] + q.unit_begin;
  const uint32_t v = i | ((c == kClassTwo && (q.flags & kQueryMerge)) ? kUnitMergeBit : 0u);
  for (uint32_t u = 0; u < q.n_units; u++) dst[u] = v;
}

// One warp per multi-unit query: folds the units' candidate lists into the final top-k.
__global__ void __launch_bounds__(kThreadsPerCta)
MergeUnitsKernel(const BatchView bv, const uint32_t *__restrict__ multi, uint32_t n_multi) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (w >= n_multi) return;
  const uint32_t qi = multi[w];
  const DevQuery q = bv.queries[qi];
  const uint32_t g0 = q.cand_begin;
  const int k = (int)q.k;
  TopK top;
  TopKInit(top);
  // four units at a time: all of their counts and candidate rows are requested before any is
  // folded in (the candidate slots of a unit always exist, 32 per unit, so the row load does not
  // have to wait for the count) — the merge is a chain of small dependent loads otherwise
  for (uint32_t u0 = 0; u0 < q.n_units; u0 += 4) {
    int cnt[4];
    wsr_hit row[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool on = u0 + i < q.n_units;
      cnt[i] = on ? bv.cand_n[g0 + u0 + i] : 0;
      row[i].doc_id = 0x7fffffff; row[i].reserved = 0; row[i].score = -1.0;
      if (on) row[i] = bv.cand[(size_t)(g0 + u0 + i) * kMaxFastK + lane];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int n = cnt[i];
      const wsr_hit h = row[i];
      const double kth = top.count == k ? TopKKth(top, k) : -1.0;
      unsigned m = __ballot_sync(kFull, lane < n && !(h.score < kth));
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        TopKInsert(top, k, __shfl_sync(kFull, h.score, src), __shfl_sync(kFull, h.doc_id, src), lane);
      }
    }
  }
  if (lane < top.count) {
    wsr_hit h;
    h.doc_id = top.d; h.reserved = 0; h.score = top.s;
    bv.hits[(size_t)q.out_slot * bv.k_stride + lane] = h;
  }
  if (lane == 0) bv.n_hits[q.out_slot] = top.count;
}

// ---- K1 entry points -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerCta)
DecodeListKernel(const DevIndexView ix, uint32_t first_block, uint32_t n_blocks,
                 uint32_t *__restrict__ docs, uint32_t *__restrict__ tfs) {
  const int lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (b >= n_blocks) return;
  const uint4 info = __ldg(&ix.blk_info[first_block + b]);
  const uint32_t n = ShN(info.z);
  uint32_t d[4], tf[4];
  DecodeDocs(ix, info, lane, d);
  DecodeTfs(ix, info, lane, tf);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t e = 4u * lane + i;
    if (e < n) {
      docs[(size_t)b * 128 + e] = d[i];
      tfs[(size_t)b * 128 + e] = tf[i];
    }
  }
}

// ---- K1, whole-index decode: a pure stream over blk_info and the payload -------------------------
// Round 1 walked the blocks with per-lane LDGs behind a dependent blk_info load (70 % issue-bound at
// 130 warp instructions per block, two exposed round trips per step). This version is a bulk-copy
// pipeline: the payload is cut into fixed 8 KB stages (+ 1 KB of overlap, the largest block), a
// persistent CTA streams its stages — payload bytes and the blk_info rows of the blocks that START
// in the stage — into a 4-deep shared-memory ring with cp.async.bulk, completion signalled on an
// mbarrier per slot, and its warps decode the staged blocks from shared memory. No global load, no
// address arithmetic on 64-bit pointers and no load latency is left in the decode loop.
constexpr uint32_t kK1StageBytes = 8192;                 // host_index.h kDecodeStageBytes
constexpr uint32_t kK1Overlap = 1024;                    // 32 x 16 B doc records + 32 x 16 B tf records
constexpr uint32_t kK1MaxBlocks = kK1StageBytes / 32u;   // a block has at least 16 + 16 payload bytes
constexpr int kK1Depth = 4;
struct __align__(128) K1Slot {
  uint4 pay[(kK1StageBytes + kK1Overlap) / 16];
  uint4 info[kK1MaxBlocks];
};
struct __align__(128) K1Shared {
  K1Slot slot[kK1Depth];
  unsigned long long full[kK1Depth];   // mbarriers: the slot's bytes have landed
  uint32_t done[kK1Depth];             // warps of the CTA that finished reading the slot
};

__device__ __forceinline__ void MbarInit(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"   // %2: suspend-time hint, the warp
      "@p bra WAIT_DONE;\n"                                            // sleeps in hardware instead of
      "bra WAIT_LOOP;\n"                                               // spinning on issue slots
      "WAIT_DONE:\n"
      "}\n" ::"r"(SmemAddr(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP): 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void BulkLoad(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(SmemAddr(dst)), "l"(src), "r"(bytes), "r"(SmemAddr(bar)) : "memory");
}

// One staged record: doc ids and tfs of the four postings of record `rec` of the block described
// by `info`, whose payload starts at `pay` in shared memory; slots past the block's n come back as
// zeros. Written for instruction count: sequential funnel shifts walk the fields of the 32/64/96/
// 128-bit record formats (w0, b <= 31: doc ids are < 2^31), the four tfs come out of one word.
__device__ __forceinline__ void K1DecodeRecord(const uint4 info, const uint4 *pay, uint32_t rec,
                                               uint32_t d[4], uint32_t t[4]) {
  const uint32_t bits = info.z, n = ShN(bits), rc = ShRcode(bits), tc = ShTcode(bits);
  const uint32_t w0 = ShW0(bits), bw = ShB(bits);
  const uint32_t m0 = 0xffffffffu >> (32u - w0), mb = 0xffffffffu >> (32u - bw);
  uint32_t f, d1, d2, d3;
  if (rc <= 1u) {
    uint32_t lo, hi = 0u;
    if (rc == 1u) {
      const uint2 v = reinterpret_cast<const uint2 *>(pay)[rec];
      lo = v.x; hi = v.y;
    } else {
      lo = reinterpret_cast<const uint32_t *>(pay)[rec];
    }
    f = lo & m0;
    lo = __funnelshift_r(lo, hi, w0); hi >>= w0;
    d1 = lo & mb;
    lo = __funnelshift_r(lo, hi, bw); hi >>= bw;
    d2 = lo & mb;
    lo = __funnelshift_r(lo, hi, bw);
    d3 = lo & mb;
  } else {
    uint32_t x0, x1, x2, x3 = 0u;
    if (rc == 2u) {
      const uint4 v = pay[rec];
      x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
    } else {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(pay) + 3u * rec;
      x0 = w[0]; x1 = w[1]; x2 = w[2];
    }
    f = x0 & m0;
    x0 = __funnelshift_r(x0, x1, w0); x1 = __funnelshift_r(x1, x2, w0); x2 = __funnelshift_r(x2, x3, w0); x3 >>= w0;
    d1 = x0 & mb;
    x0 = __funnelshift_r(x0, x1, bw); x1 = __funnelshift_r(x1, x2, bw); x2 = __funnelshift_r(x2, x3, bw);
    d2 = x0 & mb;
    x0 = __funnelshift_r(x0, x1, bw);
    d3 = x0 & mb;
  }
  d[0] = info.x + f;
  d[1] = d[0] + d1;
  d[2] = d[1] + d2;
  d[3] = d[2] + d3;
  const uint4 *ts = pay + DocGranules(bits);
  if (tc == 0u) {
    const uint32_t v = reinterpret_cast<const unsigned short *>(ts)[rec];
    t[0] = v & 15u; t[1] = (v >> 4) & 15u; t[2] = (v >> 8) & 15u; t[3] = v >> 12;
  } else if (tc == 1u) {
    const uint32_t v = reinterpret_cast<const uint32_t *>(ts)[rec];
    t[0] = v & 255u; t[1] = (v >> 8) & 255u; t[2] = (v >> 16) & 255u; t[3] = v >> 24;
  } else {
    const uint4 v = ts[rec];
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  }
  if (4u * rec + 4u > n) {   // the block's last, partial record: slots past n repeat the last posting
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (4u * rec + i >= n) { d[i] = 0u; t[i] = 0u; }
  }
}

__global__ void __launch_bounds__(kThreadsPerCta)
DecodeAllKernel(const DevIndexView ix, uint32_t n_blocks, unsigned long long *checksum) {
  extern __shared__ __align__(128) unsigned char k1_smem[];
  K1Shared *sh = reinterpret_cast<K1Shared *>(k1_smem);
  const uint32_t tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t n_stages = ix.k1_stages;
  const unsigned long long pay_bytes = (unsigned long long)ix.payload_granules * 16ull;
  if (tid == 0) {
    for (int i = 0; i < kK1Depth; i++) {
      MbarInit(&sh->full[i], 1u);
      sh->done[i] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // stage `it` of this CTA = global stage blockIdx.x + it * gridDim.x
  auto issue = [&](uint32_t it) {
    const unsigned long long stage = (unsigned long long)blockIdx.x + (unsigned long long)it * gridDim.x;
    if (stage >= n_stages) return;
    const uint32_t f0 = __ldg(&ix.k1_stage_first[stage]), f1 = __ldg(&ix.k1_stage_first[stage + 1]);
    if (f1 == f0) return;   // no block starts in this stage (only the tail padding can do that)
    const unsigned long long off = stage * kK1StageBytes;
    const uint32_t pb = (uint32_t)min((unsigned long long)(kK1StageBytes + kK1Overlap), pay_bytes - off);
    const uint32_t ib = (f1 - f0) * 16u;
    K1Slot &sl = sh->slot[it % kK1Depth];
    MbarExpectTx(&sh->full[it % kK1Depth], pb + ib);
    BulkLoad(sl.pay, reinterpret_cast<const unsigned char *>(ix.payload) + off, pb, &sh->full[it % kK1Depth]);
    BulkLoad(sl.info, ix.blk_info + f0, ib, &sh->full[it % kK1Depth]);
  };
  if (tid == 0)
    for (uint32_t it = 0; it < (uint32_t)kK1Depth; it++) issue(it);
  unsigned long long sum = 0;
  uint32_t phase_bits = 0;   // per slot: parity of the phase to wait for next
  for (uint32_t it = 0;; it++) {
    const unsigned long long stage = (unsigned long long)blockIdx.x + (unsigned long long)it * gridDim.x;
    if (stage >= n_stages) break;
    const uint32_t f0 = __ldg(&ix.k1_stage_first[stage]), f1 = __ldg(&ix.k1_stage_first[stage + 1]);
    const int slot = it % kK1Depth;
    const uint32_t nblk = f1 - f0;   // 0 only for a stage made of tail padding: nothing was copied
    if (nblk) {
      MbarWait(&sh->full[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
    }
    const K1Slot &sl = sh->slot[slot];
    const uint32_t base16 = (uint32_t)(stage * (kK1StageBytes / 16u));
    // A warp takes 4 consecutive blocks per step. Short blocks (at most 8 records: the single-block
    // lists that make up 40 % of all blocks) are decoded four at a time, one per group of 8 lanes;
    // a step that holds a longer block decodes its blocks one after the other with all 32 lanes.
    for (uint32_t j0 = 4u * (uint32_t)warp; j0 < nblk; j0 += 4u * kWarpsPerCta) {
      const uint32_t jg = min(j0 + ((uint32_t)lane >> 3), nblk - 1u);
      const uint4 ig = sl.info[jg];
      const uint32_t nlg = (ShN(ig.z) + 3u) >> 2;
      const bool in_range = j0 + ((uint32_t)lane >> 3) < nblk;
      if (__all_sync(kFull, nlg <= 8u || !in_range)) {
        const uint32_t rec = (uint32_t)lane & 7u;
        if (in_range && rec < nlg) {
          uint32_t d[4], t[4];
          K1DecodeRecord(ig, sl.pay + (ig.y - base16), rec, d, t);
          sum += (unsigned long long)(d[0] + d[1]);
          sum += (unsigned long long)(d[2] + d[3]);
          sum += (unsigned long long)(t[0] + t[1] + t[2] + t[3]);
        }
      } else {
        const uint32_t jend = min(j0 + 4u, nblk);
        for (uint32_t j = j0; j < jend; j++) {
          const uint4 info = sl.info[j];
          if ((uint32_t)lane < ((ShN(info.z) + 3u) >> 2)) {
            uint32_t d[4], t[4];
            K1DecodeRecord(info, sl.pay + (info.y - base16), (uint32_t)lane, d, t);
            sum += (unsigned long long)(d[0] + d[1]);
            sum += (unsigned long long)(d[2] + d[3]);
            sum += (unsigned long long)(t[0] + t[1] + t[2] + t[3]);
          }
        }
      }
    }
    // the warp is done with the slot; the last of the CTA's warps to get here refills it (no
    // CTA-wide barrier: warps drift apart by up to the depth of the ring)
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&sh->done[slot], 1u) == (uint32_t)kWarpsPerCta - 1u) {
        sh->done[slot] = 0u;
        __threadfence_block();
        issue(it + kK1Depth);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  if (lane == 0 && sum) atomicAdd(checksum, sum);
}

// Block-max refresh after the global statistics changed (document-partitioned load): recomputes
// every block's upper bound of tf*(k1+1)/(tf+cache[norm]) with the new cache.
__global__ void __launch_bounds__(kThreadsPerCta)
RefreshBlockMaxKernel(const DevIndexView ix, uint32_t n_blocks, uint4 *blk_info_rw, float *blk_max_rw) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * kWarpsPerCta;
  for (uint32_t b = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); b < n_blocks; b += warps) {
    const uint4 info = blk_info_rw[b];
    const uint32_t n = ShN(info.z);
    uint32_t d[4], tf[4];
    DecodeDocs(ix, info, lane, d);
    DecodeTfs(ix, info, lane, tf);
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (4u * lane + i < n) {
        const double f = (double)tf[i];
        const double x = (f * kK1Plus1) / (f + ix.cache[__ldg(ix.norms + d[i])]);
        mx = fmaxf(mx, __double2float_ru(x));
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    if (lane == 0) {
      blk_info_rw[b].w = __float_as_uint(mx);
      blk_max_rw[b] = mx;
    }
  }
}

// ---- cross-shard merge: rank of every gathered entry among all shards' entries ---------------
__device__ __forceinline__ bool HitBefore(double s1, int d1, double s2, int d2) {
  return s1 > s2 || (s1 == s2 && d1 < d2);
}
__global__ void MergeShardsKernel(const wsr_hit *__restrict__ g, const int32_t *__restrict__ gn,
                                  int n_shards, int n_queries, int k_stride, wsr_hit *out,
                                  int32_t *out_n) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_q = (long long)n_shards * k_stride;
  if (tid >= (long long)n_queries * per_q) return;
  const int q = (int)(tid / per_q);
  const int r = (int)(tid % per_q);
  const int s = r / k_stride, i = r % k_stride;
  const int my_n = gn[(size_t)s * n_queries + q];
  if (s == 0 && i == 0) {
    int tot = 0;
    for (int t = 0; t < n_shards; t++) tot += gn[(size_t)t * n_queries + q];
    out_n[q] = min(tot, k_stride);
  }
  if (i >= my_n) return;
  const wsr_hit me = g[((size_t)s * n_queries + q) * k_stride + i];
  int rank = i;
  for (int t = 0; t < n_shards; t++) {
    if (t == s) continue;
    const wsr_hit *o = g + ((size_t)t * n_queries + q) * k_stride;
    int lo = 0, hi = gn[(size_t)t * n_queries + q];
    while (lo < hi) {   // entries of shard t that come before me
      const int mid = (lo + hi) >> 1;
      if (HitBefore(o[mid].score, o[mid].doc_id, me.score, me.doc_id)) lo = mid + 1; else hi = mid;
    }
    rank += lo;
  }
  if (rank < k_stride) out[(size_t)q * k_stride + rank] = me;
}

// doc_freqs across document partitions: every partition reports the collection-wide df of the
// terms ITS dictionary holds (0 terms when one is missing there), so the merged row is the
// element-wise maximum, and the term count the maximum count.
__global__ void MergeDocFreqsKernel(const uint32_t *__restrict__ g_df, const int32_t *__restrict__ g_ndf,
                                    int n_shards, int n, uint32_t *__restrict__ out_df,
                                    int32_t *__restrict__ out_ndf) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * WSR_MAX_TERMS) return;
  const int q = (int)(t / WSR_MAX_TERMS), j = (int)(t % WSR_MAX_TERMS);
  uint32_t m = 0;
  int32_t c = 0;
  for (int s = 0; s < n_shards; s++) {
    const int32_t cs = g_ndf[(size_t)s * n + q];
    if (j < cs) m = max(m, g_df[((size_t)s * n + q) * WSR_MAX_TERMS + j]);
    c = max(c, cs);
  }
  out_df[t] = j < c ? m : 0u;
  if (j == 0) out_ndf[q] = c;
}

// ---- collect mode epilogue ---------------------------------------------------------------------
__global__ void CollectCopyKernel(const BatchView bv, uint32_t q_begin, uint32_t n_collect,
                                  const int32_t *__restrict__ doc, const double *__restrict__ score) {
  const uint32_t w = blockIdx.x;
  if (w >= n_collect) return;
  const DevQuery q = bv.queries[q_begin + w];
  const uint32_t cnt = bv.seg_count[q_begin + w];
  const uint32_t n = min(cnt, q.k);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    wsr_hit h;
    h.doc_id = doc[q.seg_begin + i]; h.reserved = 0; h.score = score[q.seg_begin + i];
    bv.hits[(size_t)q.out_slot * bv.k_stride + i] = h;
  }
  if (threadIdx.x == 0) bv.n_hits[q.out_slot] = (int32_t)n;
}

__global__ void SegEndKernel(const BatchView bv, uint32_t q_begin, uint32_t n_collect,
                             uint32_t *seg_begin, uint32_t *seg_end) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_collect) return;
  const uint32_t b = bv.queries[q_begin + i].seg_begin;
  seg_begin[i] = b;
  seg_end[i] = b + bv.seg_count[q_begin + i];
}

template <int CLASS, bool STATS, bool MERGE>
void LaunchClassKernel(const DevIndexView &ix, const BatchView &b, uint32_t nu, int sm_count, cudaStream_t s) {
  if (!nu) return;
  // persistent grid = resident CTAs per SM (occupancy query, cached) x SM count
  constexpr size_t kScratch = sizeof(typename ScratchOfKernel<CLASS, MERGE>::type) * kWarpsPerCta;
  constexpr size_t kDyn = kScratch > 40 * 1024 ? kScratch : 0;   // see SearchKernel: small scratch is static
  static int occ = 0;
  if (!occ) {
    int o = 0;
    if (kDyn) cudaFuncSetAttribute(SearchKernel<CLASS, STATS, MERGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDyn);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, SearchKernel<CLASS, STATS, MERGE>, kThreadsPerCta, kDyn);
    occ = o > 0 ? o : 1;
  }
  const uint32_t want = (nu + kWarpsPerCta - 1) / kWarpsPerCta;
  const uint32_t grid = std::min<uint32_t>(want, (uint32_t)(sm_count * occ));
  SearchKernel<CLASS, STATS, MERGE><<<grid, kThreadsPerCta, kDyn, s>>>(ix, b);
}

template <int CLASS, bool STATS>
void LaunchClass(const DevIndexView &ix, const BatchView &b, int sm_count, cudaStream_t s) {
  if constexpr (CLASS == kClassTwo) {
    // both kernels walk the class's whole unit queue and take their own path's units
    const uint32_t nu = b.class_units[CLASS];
    if (b.merge_units < nu) LaunchClassKernel<CLASS, STATS, false>(ix, b, nu, sm_count, s);
    if (b.merge_units) LaunchClassKernel<CLASS, STATS, true>(ix, b, nu, sm_count, s);
  } else {
    LaunchClassKernel<CLASS, STATS, false>(ix, b, b.class_units[CLASS], sm_count, s);
  }
}

}  // namespace

template <bool STATS>
static void LaunchSearchClassT(const DevIndexView &ix, const BatchView &b, int c, int sm_count,
                               cudaStream_t s) {
  switch (c) {
    case kClassOne: LaunchClass<kClassOne, STATS>(ix, b, sm_count, s); break;
    case kClassTwo: LaunchClass<kClassTwo, STATS>(ix, b, sm_count, s); break;
    case kClassMany: LaunchClass<kClassMany, STATS>(ix, b, sm_count, s); break;
    default: LaunchClass<kClassCollect, STATS>(ix, b, sm_count, s); break;
  }
}

void LaunchSearchClass(const DevIndexView &ix, const BatchView &b, int c, int sm_count,
                       cudaStream_t s, bool count_work) {
  if (count_work) LaunchSearchClassT<true>(ix, b, c, sm_count, s);
  else LaunchSearchClassT<false>(ix, b, c, sm_count, s);
}

void LaunchUnitMap(const BatchView &b, uint32_t *unit_query, uint32_t n_planned, cudaStream_t s) {
  if (!n_planned) return;
  UnitMapKernel<<<(n_planned + 255) / 256, 256, 0, s>>>(b, unit_query, n_planned);
}

void LaunchMerge(const BatchView &b, const uint32_t *multi_queries, uint32_t n_multi,
                 cudaStream_t s) {
  if (!n_multi) return;
  const uint32_t grid = (n_multi + kWarpsPerCta - 1) / kWarpsPerCta;
  MergeUnitsKernel<<<grid, kThreadsPerCta, 0, s>>>(b, multi_queries, n_multi);
}

void LaunchDecodeList(const DevIndexView &ix, uint32_t first_block, uint32_t n_blocks,
                      uint32_t *docs, uint32_t *tfs, cudaStream_t s) {
  if (!n_blocks) return;
  const uint32_t grid = (n_blocks + kWarpsPerCta - 1) / kWarpsPerCta;
  DecodeListKernel<<<grid, kThreadsPerCta, 0, s>>>(ix, first_block, n_blocks, docs, tfs);
}

void LaunchDecodeAll(const DevIndexView &ix, uint32_t n_blocks, unsigned long long *checksum,
                     int sm_count, cudaStream_t s) {
  if (!n_blocks || !ix.k1_stages) return;
  static int occ = 0;
  if (!occ) {
    int o = 0;
    cudaFuncSetAttribute(DecodeAllKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Shared));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, DecodeAllKernel, kThreadsPerCta, sizeof(K1Shared));
    occ = o > 0 ? o : 1;
  }
  const uint32_t grid = std::min<uint32_t>(ix.k1_stages, (uint32_t)(sm_count * occ));
  DecodeAllKernel<<<grid, kThreadsPerCta, sizeof(K1Shared), s>>>(ix, n_blocks, checksum);
}

void LaunchRefreshBlockMax(const DevIndexView &ix, uint32_t n_blocks, uint4 *blk_info_rw,
                           float *blk_max_rw, int sm_count, cudaStream_t s) {
  if (!n_blocks) return;
  const uint32_t grid = std::min<uint32_t>((n_blocks + kWarpsPerCta - 1) / kWarpsPerCta,
                                           (uint32_t)(sm_count * 8));
  RefreshBlockMaxKernel<<<grid, kThreadsPerCta, 0, s>>>(ix, n_blocks, blk_info_rw, blk_max_rw);
}

void LaunchMergeShards(const wsr_hit *gathered, const int32_t *gathered_n, int n_shards,
                       int n_queries, int k_stride, wsr_hit *out, int32_t *out_n,
                       cudaStream_t s) {
  const long long total = (long long)n_queries * n_shards * k_stride;
  if (total <= 0) return;
  const int threads = 256;
  const unsigned grid = (unsigned)((total + threads - 1) / threads);
  MergeShardsKernel<<<grid, threads, 0, s>>>(gathered, gathered_n, n_shards, n_queries, k_stride,
                                             out, out_n);
}

void LaunchMergeDocFreqs(const uint32_t *g_df, const int32_t *g_ndf, int n_shards, int n, uint32_t *out_df,
                         int32_t *out_ndf, cudaStream_t s) {
  const long long total = (long long)n * WSR_MAX_TERMS;
  if (total <= 0) return;
  MergeDocFreqsKernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(g_df, g_ndf, n_shards, n, out_df, out_ndf);
}

size_t CollectSortTempBytes(uint32_t n_entries, uint32_t n_collect) {
  size_t a = 0, b = 0;
  cub::DeviceSegmentedSort::StableSortPairs(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr,
                                            (const double *)nullptr, (double *)nullptr,
                                            (int)n_entries, (int)n_collect,
                                            (const uint32_t *)nullptr, (const uint32_t *)nullptr);
  cub::DeviceSegmentedSort::StableSortPairsDescending(nullptr, b, (const double *)nullptr,
                                                      (double *)nullptr, (const int32_t *)nullptr,
                                                      (int32_t *)nullptr, (int)n_entries,
                                                      (int)n_collect, (const uint32_t *)nullptr,
                                                      (const uint32_t *)nullptr);
  return a > b ? a : b;
}

void LaunchCollectFinish(const BatchView &b, uint32_t n_collect, uint32_t n_entries,
                         uint32_t *seg_begin, uint32_t *seg_end, int32_t *tmp_doc,
                         double *tmp_score, void *cub_tmp, size_t cub_tmp_bytes, cudaStream_t s) {
  if (!n_collect) return;
  const uint32_t q_begin = b.class_begin[kClassCollect];
  SegEndKernel<<<(n_collect + 255) / 256, 256, 0, s>>>(b, q_begin, n_collect, seg_begin, seg_end);
  if (n_entries) {
    // (1) doc id ascending, (2) STABLE score descending => (score desc, doc asc)
    size_t bytes = cub_tmp_bytes;
    cub::DeviceSegmentedSort::StableSortPairs(cub_tmp, bytes, b.seg_doc, tmp_doc, b.seg_score,
                                              tmp_score, (int)n_entries, (int)n_collect, seg_begin,
                                              seg_end, s);
    bytes = cub_tmp_bytes;
    cub::DeviceSegmentedSort::StableSortPairsDescending(cub_tmp, bytes, tmp_score, b.seg_score,
                                                        tmp_doc, b.seg_doc, (int)n_entries,
                                                        (int)n_collect, seg_begin, seg_end, s);
  }
  CollectCopyKernel<<<n_collect, 128, 0, s>>>(b, q_begin, n_collect, b.seg_doc, b.seg_score);
}

}  // namespace wsr
