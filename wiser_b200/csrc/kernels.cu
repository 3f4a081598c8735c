// Hand-written sm_100a kernels for the WiSER/Vacuum hot path: packed-block decode, conjunctive
// intersection with skip metadata, fused BM25 scoring and per-query top-k.
//
// Execution model: one WARP owns one work unit = (query, range of <=kUnitBlocks blocks of the
// query's shortest list). Units of a whole query batch sit in per-class dynamic queues
// (atomic counter) drained by a persistent grid sized to the SM count. A block is staged
// with one coalesced 128-bit load per lane, unpacked from shared memory with funnel shifts,
// and prefix-summed with warp shuffles. Everything on the path is integer/byte work bound by
// HBM bandwidth and the integer issue rate; tensor cores are not used.
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

#include <algorithm>
#include <cub/device/device_segmented_sort.cuh>

namespace wsr {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kK1Plus1 = 1.2 + 1;   // (k1_ + 1) evaluated in double, scoring.h:68

struct __align__(16) WarpScratch {
  uint32_t stage[136];   // packed words of one stream (<=128) + zero pad for the hi word
  uint32_t docs[128];    // doc ids of the most recently decoded probe-side block
};

struct CtaShared {
  double cache[256];
  float cache32[256];
  WarpScratch warp[kWarpsPerCta];
};

__device__ __forceinline__ uint32_t BlkN(uint32_t bits) { return ((bits >> 12) & 127u) + 1u; }
__device__ __forceinline__ uint32_t BlkDBits(uint32_t bits) { return bits & 63u; }
__device__ __forceinline__ uint32_t BlkTBits(uint32_t bits) { return (bits >> 6) & 63u; }
__device__ __forceinline__ uint32_t Granules(uint32_t n, uint32_t bits) { return (n * bits + 127u) >> 7; }

// Copies `nvec` 16-byte granules of one packed stream into the warp's staging buffer with one
// coalesced 128-bit load per lane; lanes past the stream store zeros (zero pad).
__device__ __forceinline__ void StageStream(const uint4 *__restrict__ src, uint32_t nvec,
                                            WarpScratch *ws, int lane) {
  __syncwarp();   // previous readers of the staging buffer are done
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if ((uint32_t)lane < nvec) v = __ldg(src + lane);
  reinterpret_cast<uint4 *>(ws->stage)[lane] = v;
  if (lane == 0) ws->stage[128] = 0u;
  __syncwarp();
}

// Lane l extracts elements 4l..4l+3 of the staged b-bit LSB-first stream.
__device__ __forceinline__ void Unpack4(const WarpScratch *ws, uint32_t bits, int lane,
                                        uint32_t v[4]) {
  const uint32_t mask = bits >= 32u ? 0xffffffffu : ((1u << bits) - 1u);
  uint32_t bit = 4u * (uint32_t)lane * bits;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t w = bit >> 5;
    const uint32_t lo = ws->stage[w], hi = ws->stage[w + 1];
    v[i] = __funnelshift_r(lo, hi, bit & 31u) & mask;
    bit += bits;
  }
}

__device__ __forceinline__ uint32_t WarpInclusiveScan(uint32_t x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += t;
  }
  return x;
}

// Decodes the doc-id stream of one block: lane l gets doc ids of elements 4l..4l+3
// (0xFFFFFFFF past the block's n postings). Doc ids are base + running sum of the deltas.
__device__ __forceinline__ void DecodeDocs4(const DevIndexView &ix, const uint4 info,
                                            WarpScratch *ws, int lane, uint32_t d[4]) {
  const uint32_t n = BlkN(info.z), dbits = BlkDBits(info.z);
  StageStream(ix.payload + info.y, Granules(n, dbits), ws, lane);
  Unpack4(ws, dbits, lane, d);
  d[1] += d[0];
  d[2] += d[1];
  d[3] += d[2];
  const uint32_t incl = WarpInclusiveScan(d[3], lane);
  const uint32_t off = info.x + incl - d[3];
#pragma unroll
  for (int i = 0; i < 4; i++) d[i] = (4u * lane + i < n) ? d[i] + off : 0xffffffffu;
}

// Decodes the tf stream of one block (elements 4l..4l+3 per lane).
__device__ __forceinline__ void DecodeTfs4(const DevIndexView &ix, const uint4 info,
                                           WarpScratch *ws, int lane, uint32_t tf[4]) {
  const uint32_t n = BlkN(info.z), dbits = BlkDBits(info.z), tbits = BlkTBits(info.z);
  StageStream(ix.payload + info.y + Granules(n, dbits), Granules(n, tbits), ws, lane);
  Unpack4(ws, tbits, lane, tf);
}

// Random access to one tf of a block straight from HBM/L2 (used for intersection hits only).
__device__ __forceinline__ uint32_t ExtractTf(const DevIndexView &ix, const uint4 info, uint32_t pos) {
  const uint32_t n = BlkN(info.z), dbits = BlkDBits(info.z), tbits = BlkTBits(info.z);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.payload + info.y + Granules(n, dbits));
  const uint32_t bit = pos * tbits;
  const uint32_t lo = __ldg(w + (bit >> 5)), hi = __ldg(w + (bit >> 5) + 1);
  const uint32_t mask = tbits >= 32u ? 0xffffffffu : ((1u << tbits) - 1u);
  return __funnelshift_r(lo, hi, bit & 31u) & mask;
}

// First block index in [lo, hi) whose last doc id is >= x, or hi.
__device__ __forceinline__ uint32_t LowerBoundBlock(const uint32_t *__restrict__ last, uint32_t lo,
                                                    uint32_t hi, uint32_t x) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(last + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Position of the first doc id >= x among the 128 staged doc ids (sentinel-padded).
__device__ __forceinline__ uint32_t LowerBound128(const uint32_t *docs, uint32_t x) {
  uint32_t pos = 0;
#pragma unroll
  for (uint32_t step = 64; step; step >>= 1)
    if (docs[pos + step - 1] < x) pos += step;
  return pos;
}

// TfNormLossy + one term of CalcDocScoreLossy with every operation rounded separately
// (the reference build has no FMA contraction).
__device__ __forceinline__ double TermScore(double idf, uint32_t tf, double cache_norm) {
  const double f = (double)tf;
  const double tfnorm = __ddiv_rn(__dmul_rn(f, kK1Plus1), __dadd_rn(f, cache_norm));
  return __dmul_rn(idf, tfnorm);
}

// ---- per-warp top-k: lane r holds the rank-r entry, ordered (score desc, doc asc) --------
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

// 32-ary cooperative search: last planned query in [lo, hi) whose unit_begin <= u.
__device__ __forceinline__ uint32_t FindQuery(const DevQuery *__restrict__ q, uint32_t lo,
                                              uint32_t hi, uint32_t u, int lane) {
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + 31u) >> 5;
    const uint32_t idx = lo + (uint32_t)lane * step;
    const bool le = idx < hi && __ldg(&q[idx].unit_begin) <= u;
    const int cnt = __popc(__ballot_sync(kFull, le));
    lo += (uint32_t)(cnt - 1) * step;
    hi = min(lo + step, hi);
  }
  return lo;
}

struct UnitStats {
  unsigned long long decoded, bytes, matches;
};

// Emits one unit's result: straight to the caller's hit array when the query has one unit,
// else to the unit's candidate slot for the merge pass.
__device__ __forceinline__ void EmitTopK(const BatchView &bv, const DevQuery &q, uint32_t local,
                                         const TopK &t, int lane) {
  const uint32_t gunit = q.cand_begin + local;
  if (q.n_units == 1) {
    if (lane < t.count) {
      wsr_hit h;
      h.doc_id = t.d; h.reserved = 0; h.score = t.s;
      bv.hits[(size_t)q.out_slot * bv.k_stride + lane] = h;
    }
    if (lane == 0) bv.n_hits[q.out_slot] = t.count;
  } else {
    if (lane < t.count) {
      wsr_hit h;
      h.doc_id = t.d; h.reserved = 0; h.score = t.s;
      bv.cand[(size_t)gunit * kMaxFastK + lane] = h;
    }
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
    bv.seg_doc[at] = doc;
    bv.seg_score[at] = score;
  }
}

// ---- single-term units: SingleTermQueryProcessor::Process, query_processing.h:632-641 ----
// Every posting is a hit. Blocks whose block-max score cannot reach the running k-th score
// are skipped without touching their payload; surviving postings are pre-filtered with an
// fp32 upper bound and only candidates are re-scored in exact fp64.
template <bool COLLECT>
__device__ void ProcessOneTerm(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                               uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                               CtaShared *sh, WarpScratch *ws, int lane, UnitStats &st) {
  const uint32_t term = q.term[0];
  const uint4 li = __ldg(&ix.lists[term]);
  const double idf = __ldg(&ix.idf[term]);
  const float idf_up = __double2float_ru(idf);
  const int k = (int)q.k;
  const bool multi = q.n_units > 1;
  TopK top;
  TopKInit(top);
  double thr_shared = 0.0;
  double published = 0.0;

  for (uint32_t j = b0; j < b1; j++) {
    const uint4 info = __ldg(&ix.blk_info[li.x + j]);
    const uint32_t n = BlkN(info.z);
    double kth = -1.0;
    if (!COLLECT) {
      if (multi) {
        const unsigned long long tb = __ldcg(&bv.thr[qi]);
        thr_shared = fmax(thr_shared, __longlong_as_double((long long)tb));
      }
      const bool full = top.count == k;
      kth = full ? TopKKth(top, k) : -1.0;
      // upper bound of every exact score in the block (rounded up at each step)
      const double ub = (double)(idf_up * __uint_as_float(info.w) * 1.00001f);
      if (ub < thr_shared || (full && ub <= kth)) continue;
    }
    uint32_t d[4], tf[4];
    DecodeDocs4(ix, info, ws, lane, d);
    DecodeTfs4(ix, info, ws, lane, tf);
    st.decoded += n;
    st.bytes += 16ull * (Granules(n, BlkDBits(info.z)) + Granules(n, BlkTBits(info.z))) + 16ull;
    st.matches += n;

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
          const double thr_eff = fmax(thr_shared, kth);
          pass[i] = (double)s32 >= thr_eff;
        }
        if (pass[i]) {
          s64[i] = TermScore(idf, tf[i], sh->cache[nb]);
          s64[i] = __dadd_rn(0.0, s64[i]);
          if (!COLLECT) pass[i] = !(s64[i] < thr_shared) && !(s64[i] < kth);
        }
      }
    }
    if (COLLECT) {
#pragma unroll
      for (int i = 0; i < 4; i++) CollectAppend(bv, q, qi, pass[i], (int)d[i], s64[i], lane);
    } else {
      bool inserted = false;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        unsigned m = __ballot_sync(kFull, pass[i]);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const double s = __shfl_sync(kFull, s64[i], src);
          const int dd = (int)__shfl_sync(kFull, d[i], src);
          TopKInsert(top, k, s, dd, lane);
          inserted = true;
        }
      }
      if (multi && inserted && top.count == k) {
        const double nk = TopKKth(top, k);
        if (nk > published) {
          published = nk;
          if (lane == 0) atomicMax(&bv.thr[qi], (unsigned long long)__double_as_longlong(nk));
        }
      }
    }
  }
  if (!COLLECT) EmitTopK(bv, q, local, top, lane);
}

// ---- multi-term units: shortest list drives, the other lists are probed through their
// per-block last-doc skip metadata; only blocks that can contain a candidate are decoded.
template <int M, bool COLLECT>
__device__ void ProcessMulti(const DevIndexView &ix, const BatchView &bv, const DevQuery &q,
                             uint32_t qi, uint32_t local, uint32_t b0, uint32_t b1,
                             CtaShared *sh, WarpScratch *ws, int lane, UnitStats &st) {
  const int m = (int)q.n_terms;
  const int drv = (int)q.driver;
  const int k = (int)q.k;
  const bool multi = q.n_units > 1;
  uint32_t first[M], nblk[M], cur[M];
  double idf[M];
  uint32_t first_a = 0;
#pragma unroll
  for (int t = 0; t < M; t++) {
    first[t] = nblk[t] = cur[t] = 0;
    idf[t] = 0.0;
    if (t < m) {
      const uint4 li = __ldg(&ix.lists[q.term[t]]);
      first[t] = li.x;
      nblk[t] = li.y;
      idf[t] = __ldg(&ix.idf[q.term[t]]);
      if (t == drv) first_a = li.x;
    }
  }
  TopK top;
  TopKInit(top);
  double published = 0.0;
  uint32_t cached_blk = 0xffffffffu;
  uint4 cached_info = make_uint4(0u, 0u, 0u, 0u);

  for (uint32_t ja = b0; ja < b1; ja++) {
    const uint4 info_a = __ldg(&ix.blk_info[first_a + ja]);
    const uint32_t na = BlkN(info_a.z);
    uint32_t d[4];
    DecodeDocs4(ix, info_a, ws, lane, d);
    st.decoded += na;
    st.bytes += 16ull * Granules(na, BlkDBits(info_a.z)) + 16ull;
    bool al[4];
#pragma unroll
    for (int i = 0; i < 4; i++) al[i] = 4u * lane + i < na;
    uint32_t tfv[M][4];

#pragma unroll
    for (int t = 0; t < M; t++) {
      if (t >= m || t == drv) continue;
      // warp-wide doc-id range still alive
      uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (al[i]) { mn = min(mn, d[i]); mx = max(mx, d[i]); }
      mn = __reduce_min_sync(kFull, mn);
      mx = __reduce_max_sync(kFull, mx);
      if (mn == 0xffffffffu) continue;   // nothing alive
      const uint32_t *last = ix.blk_last + first[t];
      const uint32_t jlo = LowerBoundBlock(last, cur[t], nblk[t], mn);
      cur[t] = jlo;
      uint32_t jhi = jlo < nblk[t] ? LowerBoundBlock(last, jlo, nblk[t], mx) : jlo;
      jhi = min(jhi + 1u, nblk[t]);
      uint32_t jb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        jb[i] = 0xffffffffu;
        if (al[i]) {
          const uint32_t j = LowerBoundBlock(last, jlo, jhi, d[i]);
          if (j >= nblk[t]) al[i] = false; else jb[i] = j;
        }
      }
      for (;;) {
        uint32_t jm = min(min(jb[0], jb[1]), min(jb[2], jb[3]));
        jm = __reduce_min_sync(kFull, jm);
        if (jm == 0xffffffffu) break;
        const uint32_t gblk = first[t] + jm;
        if (cached_blk != gblk) {
          cached_info = __ldg(&ix.blk_info[gblk]);
          uint32_t e[4];
          DecodeDocs4(ix, cached_info, ws, lane, e);
          __syncwarp();
          reinterpret_cast<uint4 *>(ws->docs)[lane] = make_uint4(e[0], e[1], e[2], e[3]);
          __syncwarp();
          cached_blk = gblk;
          const uint32_t nbk = BlkN(cached_info.z);
          st.decoded += nbk;
          st.bytes += 16ull * Granules(nbk, BlkDBits(cached_info.z)) + 16ull;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (jb[i] == jm) {
            const uint32_t pos = LowerBound128(ws->docs, d[i]);
            if (ws->docs[pos] == d[i]) {
              tfv[t][i] = ExtractTf(ix, cached_info, pos);
              st.bytes += 8ull;
            } else {
              al[i] = false;
            }
            jb[i] = 0xffffffffu;
          }
        }
        __syncwarp();
      }
    }

    // survivors matched every list: score in QUERY order, fp64, one rounding per operation
    bool hit[4];
    double s64[4];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      hit[i] = al[i];
      s64[i] = 0.0;
      if (al[i]) {
        const uint32_t tfa = ExtractTf(ix, info_a, 4u * lane + i);
        const double cn = sh->cache[__ldg(ix.norms + d[i])];
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < M; t++) {
          if (t < m) s = __dadd_rn(s, TermScore(idf[t], t == drv ? tfa : tfv[t][i], cn));
        }
        s64[i] = s;
        any = true;
      }
    }
    const unsigned anym = __ballot_sync(kFull, any);
    if (!anym) continue;
    st.matches += __popc(__ballot_sync(kFull, hit[0])) + __popc(__ballot_sync(kFull, hit[1])) +
                  __popc(__ballot_sync(kFull, hit[2])) + __popc(__ballot_sync(kFull, hit[3]));
    if (COLLECT) {
#pragma unroll
      for (int i = 0; i < 4; i++) CollectAppend(bv, q, qi, hit[i], (int)d[i], s64[i], lane);
    } else {
      double thr_shared = 0.0;
      if (multi) thr_shared = __longlong_as_double((long long)__ldcg(&bv.thr[qi]));
      bool inserted = false;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        unsigned mm = __ballot_sync(kFull, hit[i] && !(s64[i] < thr_shared));
        while (mm) {
          const int src = __ffs(mm) - 1;
          mm &= mm - 1;
          const double s = __shfl_sync(kFull, s64[i], src);
          const int dd = (int)__shfl_sync(kFull, d[i], src);
          TopKInsert(top, k, s, dd, lane);
          inserted = true;
        }
      }
      if (multi && inserted && top.count == k) {
        const double nk = TopKKth(top, k);
        if (nk > published) {
          published = nk;
          if (lane == 0) atomicMax(&bv.thr[qi], (unsigned long long)__double_as_longlong(nk));
        }
      }
    }
  }
  if (!COLLECT) EmitTopK(bv, q, local, top, lane);
}

// Persistent search kernel of one query class: warps drain the class's unit queue.
template <int CLASS>
__global__ void __launch_bounds__(kThreadsPerCta)
SearchKernel(const DevIndexView ix, const BatchView bv) {
  __shared__ CtaShared sh;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const double c = ix.cache[i];
    sh.cache[i] = c;
    sh.cache32[i] = __double2float_rd(c);   // smaller denominator => larger (safe) bound
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WarpScratch *ws = &sh.warp[threadIdx.x >> 5];
  const uint32_t q_lo = bv.class_begin[CLASS], q_hi = bv.class_begin[CLASS + 1];
  const uint32_t n_units = bv.class_units[CLASS];
  UnitStats st = {0ull, 0ull, 0ull};
  unsigned long long units = 0;
  for (;;) {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(&bv.counters->next_unit[CLASS], 1u);
    u = __shfl_sync(kFull, u, 0);
    if (u >= n_units) break;
    const uint32_t qi = FindQuery(bv.queries, q_lo, q_hi, u, lane);
    const DevQuery q = bv.queries[qi];
    const uint32_t local = u - q.unit_begin;
    // driver list block range of this unit
    const uint4 li = __ldg(&ix.lists[q.term[q.driver]]);
    const uint32_t b0 = local * kUnitBlocks;
    const uint32_t b1 = min(b0 + kUnitBlocks, li.y);
    if (CLASS == kClassOne) {
      ProcessOneTerm<false>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    } else if (CLASS == kClassTwo) {
      ProcessMulti<2, false>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    } else if (CLASS == kClassMany) {
      ProcessMulti<WSR_MAX_TERMS, false>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    } else {
      if (q.n_terms == 1) ProcessOneTerm<true>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
      else ProcessMulti<WSR_MAX_TERMS, true>(ix, bv, q, qi, local, b0, b1, &sh, ws, lane, st);
    }
    units++;
  }
  if (lane == 0 && units) {
    atomicAdd(&bv.counters->decoded_postings, st.decoded);
    atomicAdd(&bv.counters->touched_bytes, st.bytes);
    atomicAdd(&bv.counters->matches, st.matches);
    atomicAdd(&bv.counters->units, units);
  }
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
  for (uint32_t u = 0; u < q.n_units; u++) {
    const int n = bv.cand_n[g0 + u];
    wsr_hit h;
    h.doc_id = 0x7fffffff; h.score = -1.0;
    if (lane < n) h = bv.cand[(size_t)(g0 + u) * kMaxFastK + lane];
    const double kth = top.count == k ? TopKKth(top, k) : -1.0;
    unsigned m = __ballot_sync(kFull, lane < n && !(h.score < kth));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      TopKInsert(top, k, __shfl_sync(kFull, h.score, src), __shfl_sync(kFull, h.doc_id, src), lane);
    }
  }
  if (lane < top.count) {
    wsr_hit h;
    h.doc_id = top.d; h.reserved = 0; h.score = top.s;
    bv.hits[(size_t)q.out_slot * bv.k_stride + lane] = h;
  }
  if (lane == 0) bv.n_hits[q.out_slot] = top.count;
}

// ---- K1: block decode ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerCta)
DecodeListKernel(const DevIndexView ix, uint32_t first_block, uint32_t n_blocks,
                 uint32_t *__restrict__ docs, uint32_t *__restrict__ tfs) {
  __shared__ WarpScratch wsh[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  WarpScratch *ws = &wsh[threadIdx.x >> 5];
  const uint32_t b = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (b >= n_blocks) return;
  const uint4 info = __ldg(&ix.blk_info[first_block + b]);
  const uint32_t n = BlkN(info.z);
  uint32_t d[4], tf[4];
  DecodeDocs4(ix, info, ws, lane, d);
  DecodeTfs4(ix, info, ws, lane, tf);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t e = 4u * lane + i;
    if (e < n) {
      docs[(size_t)b * 128 + e] = d[i];
      tfs[(size_t)b * 128 + e] = tf[i];
    }
  }
}

__global__ void __launch_bounds__(kThreadsPerCta)
DecodeAllKernel(const DevIndexView ix, uint32_t n_blocks, unsigned long long *checksum) {
  __shared__ WarpScratch wsh[kWarpsPerCta];
  const int lane = threadIdx.x & 31;
  WarpScratch *ws = &wsh[threadIdx.x >> 5];
  const uint32_t warps = gridDim.x * kWarpsPerCta;
  unsigned long long sum = 0;
  for (uint32_t b = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); b < n_blocks; b += warps) {
    const uint4 info = __ldg(&ix.blk_info[b]);
    const uint32_t n = BlkN(info.z);
    uint32_t d[4], tf[4];
    DecodeDocs4(ix, info, ws, lane, d);
    DecodeTfs4(ix, info, ws, lane, tf);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (4u * lane + i < n) sum += (unsigned long long)d[i] + tf[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  if (lane == 0 && sum) atomicAdd(checksum, sum);
}

// ---- cross-shard merge: rank of every gathered entry among all shards' entries -----------
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

// ---- collect mode epilogue -----------------------------------------------------------------
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

}  // namespace

void LaunchSearchClass(const DevIndexView &ix, const BatchView &b, int c, int sm_count,
                       cudaStream_t s) {
  const int ctas_per_sm = 4;
  {
    const uint32_t nu = b.class_units[c];
    if (nu) {
      const uint32_t want = (nu + kWarpsPerCta - 1) / kWarpsPerCta;
      const uint32_t grid = std::min<uint32_t>(want, (uint32_t)(sm_count * ctas_per_sm));
      switch (c) {
        case kClassOne: SearchKernel<kClassOne><<<grid, kThreadsPerCta, 0, s>>>(ix, b); break;
        case kClassTwo: SearchKernel<kClassTwo><<<grid, kThreadsPerCta, 0, s>>>(ix, b); break;
        case kClassMany: SearchKernel<kClassMany><<<grid, kThreadsPerCta, 0, s>>>(ix, b); break;
        default: SearchKernel<kClassCollect><<<grid, kThreadsPerCta, 0, s>>>(ix, b); break;
      }
    }
  }
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
  if (!n_blocks) return;
  const uint32_t grid = std::min<uint32_t>((n_blocks + kWarpsPerCta - 1) / kWarpsPerCta,
                                      (uint32_t)(sm_count * 8));
  DecodeAllKernel<<<grid, kThreadsPerCta, 0, s>>>(ix, n_blocks, checksum);
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
