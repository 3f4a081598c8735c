// Device front end of the query-log path: the text of a query log goes to the GPU as it is and
// comes out as the planned, class-sorted DevQuery array the search kernels consume.
//
//   reference                                                   here
//   QueryProducerByLog: trim, phrase quotes, explode(' ')       ParsePlanKernel (one thread per line)
//     (query_pool.h:251-352)
//   TermTrieIndex::Find (term_index.h:136-144)                  DictFind: the host TermDict's table in HBM
//   VacuumEngine::Search early outs (vacuum_engine.h:206-215)   ParsePlanKernel validity rules
//   — (host PlanBatch in wsr_capi.cu)                           ParsePlanKernel + one scan + PlaceKernel
//
// The host planner (PlanBatch) stays the path for wsr_search_batch and for k > kMaxFastK; both
// planners produce the same plan (tests/test_gpu_parity.py compares them).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "kernels.cuh"

namespace wsr {
namespace {

constexpr uint32_t kEmptySlot = 0xffffffffu;

__device__ __forceinline__ bool IsSpace(char c) {   // isspace() in the "C" locale
  return c == ' ' || (c >= '\t' && c <= '\r');
}

// TermDict::Find (host_index.cc) on the device copy of the table.
__device__ uint32_t DictFind(const DevDict &d, const char *s, uint32_t len) {
  unsigned long long h = 0xcbf29ce484222325ull;
  for (uint32_t i = 0; i < len; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ull; }
  h ^= h >> 29;
  const uint32_t tag = (uint32_t)(h >> 32);
  uint32_t at = (uint32_t)h & d.mask;
  for (;;) {
    const uint2 slot = __ldg(&d.slots[at]);
    if (slot.x == kEmptySlot) return WSR_TERM_ABSENT;
    if (slot.y == tag) {
      const uint32_t a = __ldg(&d.term_off[slot.x]), b = __ldg(&d.term_off[slot.x + 1]);
      if (b - a == len) {
        uint32_t i = 0;
        while (i < len && __ldg(&d.arena[a + i]) == s[i]) i++;
        if (i == len) return slot.x;
      }
    }
    at = (at + 1) & d.mask;
  }
}

// Guided scheduling: a class's unit queue is drained in log order, so the queries near the end of
// a batch get smaller units and the persistent grid's tail is made of short units (a full
// 128-block unit is ~0.35 ms of one warp's time). Same rule in PlanBatch (host) and PlanOne (device).
__device__ __forceinline__ unsigned long long UnitCapAt(unsigned long long i, unsigned long long n, unsigned long long cap) {
  // last eighth of the batch: a quarter of the size (last quarter, two levels, an eighth of the
  // size and 192-block units with two levels all measured 0.5 - 0.8 % slower)
  const unsigned long long c = 8 * i >= 7 * n ? cap / 4 : cap;
  return c < 1 ? 1 : c;
}

// Validity, driver list, unit size and class of one query whose term ids are in q.term[] — the
// device statement of the host planner's per-query work (PlanBatch in wsr_capi.cu). tmp[i] is the
// query before placement (cand_begin holds its class, 255 = produces no work).
__device__ __forceinline__ void PlanOne(const DevIndexView &ix, DevQuery &q, uint32_t n_terms, uint32_t flags,
                                        uint32_t k, uint32_t i, uint32_t n_all, bool resolvable, DevQuery *__restrict__ tmp,
                                        PlanItem *__restrict__ item, uint32_t *__restrict__ err) {
  PlanItem it;
#pragma unroll
  for (int j = 0; j < kPlanWords; j++) it.v[j] = 0;
  uint32_t cls = 255;
  bool ok = resolvable && k > 0 && n_terms > 0;             // vacuum_engine.h:206-215
  uint32_t best = 0, best_df = 0xffffffffu;
  unsigned long long all_blocks = 0;
  if (ok) {
    for (uint32_t t = 0; t < n_terms; t++) {
      const uint4 li = __ldg(&ix.lists[q.term[t]]);
      if (li.z == 0) ok = false;                           // nothing of this list on this shard
      if (li.z < best_df) { best_df = li.z; best = t; }
      all_blocks += li.y;
    }
  }
  if (ok) {
    const uint32_t drv_blocks = __ldg(&ix.lists[q.term[best]]).y;
    const unsigned long long probe_blocks = all_blocks - drv_blocks;
    // unit size: same rule as the host planner (PlanBatch)
    const unsigned long long ratio = drv_blocks ? (probe_blocks + drv_blocks - 1) / drv_blocks : 0;
    const unsigned long long cap = UnitCapAt(i, n_all, kUnitBlocks);
    unsigned long long ub = (unsigned long long)kUnitBudget / (1ull + ratio);
    ub = ub < 1 ? 1 : ub > cap ? cap : ub;
    q.n_terms = (uint8_t)n_terms;
    q.flags = (flags && n_terms > 1) ? 1 : 0;              // a one-term "phrase" is a plain query
    if (q.flags && ix.positions == nullptr) atomicOr(err, 2u);
    q.unit_blocks = (uint16_t)ub;
    q.k = k;
    q.driver = best;
    q.out_slot = i;
    cls = n_terms == 1 ? kClassOne : n_terms == 2 ? kClassTwo : kClassMany;
    if (UseMergePath(n_terms, q.flags & kQueryPhrase, drv_blocks, probe_blocks, ix.merge_ratio_x4))
      q.flags |= kQueryMerge;
    q.n_units = cls == kClassOne ? 1u : (drv_blocks + (uint32_t)ub - 1u) / (uint32_t)ub;
    it.v[cls] = 1;
    it.v[3 + cls] = q.n_units;
    if (q.n_units > 1) { it.v[6] = q.n_units; it.v[7] = 1; }
    if (q.flags & kQueryMerge) it.v[8] = q.n_units;
  }
  q.cand_begin = cls;
  // VacuumEngine::Search fills doc_freqs unless it returned early (n_results == 0, no terms, a
  // term missing from the dictionary: vacuum_engine.h:206-219); kept for DocFreqsKernel
  q.seg_begin = (resolvable && k > 0 && n_terms > 0) ? n_terms : 0u;
  tmp[i] = q;
  item[i] = it;
}

// doc_freqs of a planned log (SearchResult::doc_freqs, vacuum_engine.h:217-219): per input query
// the collection-wide df of its terms in query order, and how many there are (0 on the early-out
// paths). Reads the unplaced queries the planner left in tmp[].
__global__ void DocFreqsKernel(const DevQuery *__restrict__ tmp, uint32_t n, const DevIndexView ix,
                               uint32_t *__restrict__ doc_freqs, int32_t *__restrict__ n_doc_freqs) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t / WSR_MAX_TERMS, j = t % WSR_MAX_TERMS;
  if (i >= n) return;
  const uint32_t m = tmp[i].seg_begin;
  if (j == 0) n_doc_freqs[i] = (int32_t)m;
  doc_freqs[t] = j < m ? __ldg(&ix.lists[tmp[i].term[j]]).w : 0u;
}

// Same planning from term ids the host already resolved (wsr_search_batch): one thread per
// wsr_query. err bit 2: k > k_stride or a term id outside the index.
__global__ void PlanQueriesKernel(const wsr_query *__restrict__ in, uint32_t n, uint32_t k_stride,
                                  const DevIndexView ix, DevQuery *__restrict__ tmp,
                                  PlanItem *__restrict__ item, uint32_t *__restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const wsr_query w = in[i];
  DevQuery q;
  memset(&q, 0, sizeof(q));
  bool resolvable = true;
  uint32_t n_terms = w.n_terms;
  if (n_terms > WSR_MAX_TERMS) { atomicOr(err, 1u); n_terms = 0; resolvable = false; }
  if (w.k > k_stride) { atomicOr(err, 4u); resolvable = false; }
  for (uint32_t t = 0; t < n_terms; t++) {
    const uint32_t id = w.term_ids[t];
    q.term[t] = id;
    if (id == WSR_TERM_ABSENT) resolvable = false;                      // vacuum_engine.h:213-215
    else if (id >= ix.n_terms) { atomicOr(err, 4u); resolvable = false; }
  }
  PlanOne(ix, q, n_terms, w.flags & 1u, w.k, i, n, resolvable, tmp, item, err);
}

// One thread per log line: parse -> term ids -> validity -> driver list, unit size, class.
// tmp[i] is the query before placement (cand_begin holds its class, 255 = produces no work).
__global__ void ParsePlanKernel(const char *__restrict__ text, uint32_t len,
                                const uint32_t *__restrict__ nl, uint32_t n_nl, uint32_t n_lines,
                                uint32_t k, const DevDict dict, const DevIndexView ix,
                                DevQuery *__restrict__ tmp, PlanItem *__restrict__ item,
                                uint32_t *__restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lines) return;
  uint32_t a = i ? __ldg(&nl[i - 1]) + 1u : 0u;
  uint32_t b = i < n_nl ? __ldg(&nl[i]) : len;
  while (a < b && IsSpace(text[a])) a++;                 // utils::trim
  while (b > a && IsSpace(text[b - 1])) b--;
  uint32_t flags = 0;
  if (b - a >= 1 && text[a] == '"' && text[b - 1] == '"') {   // QueryProducerByLog::IsPhrase
    flags = 1;
    a++;
    if (b > a) b--;
  }
  DevQuery q;
  memset(&q, 0, sizeof(q));
  uint32_t n_terms = 0;
  bool present = true, too_many = false;
  for (uint32_t t = a; t < b;) {                          // utils::explode(line, ' ')
    while (t < b && text[t] == ' ') t++;
    uint32_t u = t;
    while (u < b && text[u] != ' ') u++;
    if (u > t) {
      if (n_terms >= WSR_MAX_TERMS) { too_many = true; break; }
      const uint32_t id = DictFind(dict, text + t, u - t);
      present = present && id != WSR_TERM_ABSENT;
      q.term[n_terms++] = id;
    }
    t = u;
  }
  if (too_many) atomicOr(err, 1u);
  PlanOne(ix, q, n_terms, flags, k, i, n_lines, !too_many && present, tmp, item, err);
}

struct PlanAdd {
  __device__ __forceinline__ PlanItem operator()(const PlanItem &x, const PlanItem &y) const {
    PlanItem r;
#pragma unroll
    for (int j = 0; j < kPlanWords; j++) r.v[j] = x.v[j] + y.v[j];
    return r;
  }
};

// Scatters tmp[] into the class-sorted plan (batch order kept within a class) and writes the
// totals the host needs to size buffers and launch the search kernels.
__global__ void PlaceKernel(const DevQuery *__restrict__ tmp, const PlanItem *__restrict__ item,
                            const PlanItem *__restrict__ excl, uint32_t n_lines,
                            DevQuery *__restrict__ planned, uint32_t *__restrict__ multi,
                            PlanItem *__restrict__ totals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lines) return;
  const PlanItem tot = PlanAdd()(excl[n_lines - 1], item[n_lines - 1]);
  if (i == 0) *totals = tot;
  DevQuery q = tmp[i];
  const uint32_t cls = q.cand_begin;
  if (cls == 255) return;
  const PlanItem ex = excl[i];
  uint32_t pos = ex.v[cls];
  for (uint32_t c = 0; c < cls; c++) pos += tot.v[c];
  q.unit_begin = ex.v[3 + cls];
  q.cand_begin = 0;
  if (q.n_units > 1) {
    q.cand_begin = ex.v[6];
    multi[ex.v[7]] = pos;
  }
  planned[pos] = q;
}

// Back end of the log path: results of a batch packed densely for the trip to the host. Most
// queries of a real log return fewer than k hits (the two-term benchmark log: 2.9 of 10 on
// average), so copying the [n, k] array moves mostly unused slots.
__global__ void PackResultsKernel(const wsr_hit *__restrict__ hits, const int32_t *__restrict__ n_hits,
                                  const int32_t *__restrict__ off, uint32_t n, uint32_t k,
                                  wsr_hit *__restrict__ packed) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t q = t / k, j = t - q * k;
  if (q >= n || (int32_t)j >= n_hits[q]) return;
  const int4 v = *reinterpret_cast<const int4 *>(&hits[(size_t)q * k + j]);
  *reinterpret_cast<int4 *>(&packed[(size_t)off[q] + j]) = v;
}

struct IsNewline {
  const char *text;
  __device__ __forceinline__ bool operator()(uint32_t i) const { return text[i] == '\n'; }
};

}  // namespace

// Number of '\n' bytes of the log text: 16 bytes per thread, one atomic per warp.
__global__ void CountNewlinesKernel(const char *__restrict__ text, uint32_t len, uint32_t *__restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t at = (size_t)i * 16u;
  uint32_t c = 0;
  if (at + 16u <= len) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text) + i);   // cudaMalloc'd: 16-byte aligned
    c = (uint32_t)(__popc(__vcmpeq4(v.x, 0x0a0a0a0au)) + __popc(__vcmpeq4(v.y, 0x0a0a0a0au)) +
                   __popc(__vcmpeq4(v.z, 0x0a0a0a0au)) + __popc(__vcmpeq4(v.w, 0x0a0a0a0au))) >> 3;
  } else {
    for (size_t j = at; j < len; j++) c += text[j] == '\n';
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
void LaunchCountNewlines(const char *d_text, uint32_t len, uint32_t *d_count, cudaStream_t s) {
  if (!len) return;
  const uint32_t threads = (len + 15u) / 16u;
  CountNewlinesKernel<<<(threads + 255u) / 256u, 256, 0, s>>>(d_text, len, d_count);
}

size_t FrontEndTempBytes(uint32_t len, uint32_t n_lines) {
  size_t a = 0, b = 0;
  thrust::counting_iterator<uint32_t> idx(0);
  cub::DeviceSelect::If(nullptr, a, idx, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)len,
                        IsNewline{nullptr});
  cub::DeviceScan::ExclusiveScan(nullptr, b, (const PlanItem *)nullptr, (PlanItem *)nullptr,
                                 PlanAdd(), PlanItem(), (int)n_lines);
  return (a > b ? a : b) + 256;
}

void LaunchPlanQueries(const wsr_query *d_in, uint32_t n, uint32_t k_stride, const DevIndexView &ix,
                       DevQuery *d_tmp, PlanItem *d_item, PlanItem *d_excl, DevQuery *d_planned,
                       uint32_t *d_multi, PlanItem *d_totals, uint32_t *d_err, void *d_cub, size_t cub_bytes,
                       cudaStream_t s) {
  if (!n) return;
  const uint32_t grid = (n + 127) / 128;
  PlanQueriesKernel<<<grid, 128, 0, s>>>(d_in, n, k_stride, ix, d_tmp, d_item, d_err);
  size_t bytes = cub_bytes;
  cub::DeviceScan::ExclusiveScan(d_cub, bytes, d_item, d_excl, PlanAdd(), PlanItem(), (int)n, s);
  PlaceKernel<<<grid, 128, 0, s>>>(d_tmp, d_item, d_excl, n, d_planned, d_multi, d_totals);
}

void LaunchDocFreqs(const DevQuery *d_tmp, uint32_t n, const DevIndexView &ix, uint32_t *d_doc_freqs,
                    int32_t *d_n_doc_freqs, cudaStream_t s) {
  if (!n) return;
  const unsigned long long threads = (unsigned long long)n * WSR_MAX_TERMS;
  DocFreqsKernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_tmp, n, ix, d_doc_freqs, d_n_doc_freqs);
}

size_t PackTempBytes(uint32_t n) {
  size_t a = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n + 1);
  return a + 256;
}

void LaunchResultOffsets(const int32_t *d_n_hits, int32_t *d_off, uint32_t n, void *d_cub, size_t cub_bytes,
                         cudaStream_t s) {
  // n + 1 inputs (the slot after the last count is kept zero): d_off[n] = total number of hits
  size_t bytes = cub_bytes;
  cub::DeviceScan::ExclusiveSum(d_cub, bytes, d_n_hits, d_off, (int)n + 1, s);
}

void LaunchPackResults(const wsr_hit *d_hits, const int32_t *d_n_hits, const int32_t *d_off, uint32_t n,
                       uint32_t k, wsr_hit *d_packed, cudaStream_t s) {
  const unsigned long long threads = (unsigned long long)n * k;
  if (!threads) return;
  PackResultsKernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_hits, d_n_hits, d_off, n, k, d_packed);
}

void LaunchFrontEnd(const char *d_text, uint32_t len, uint32_t n_nl, uint32_t n_lines, uint32_t k,
                    const DevDict &dict, const DevIndexView &ix, uint32_t *d_nl, uint32_t *d_n_nl,
                    DevQuery *d_tmp, PlanItem *d_item, PlanItem *d_excl, DevQuery *d_planned,
                    uint32_t *d_multi, PlanItem *d_totals, uint32_t *d_err, void *d_cub,
                    size_t cub_bytes, cudaStream_t s) {
  if (!n_lines) return;
  // newline positions, ascending. Their number n_nl is known (LaunchCountNewlines; n_lines = n_nl,
  // + 1 when the log does not end in '\n'), so the select's own count is not read back.
  thrust::counting_iterator<uint32_t> idx(0);
  size_t bytes = cub_bytes;
  cub::DeviceSelect::If(d_cub, bytes, idx, d_nl, d_n_nl, (int)len, IsNewline{d_text}, s);
  const uint32_t grid = (n_lines + 127) / 128;
  ParsePlanKernel<<<grid, 128, 0, s>>>(d_text, len, d_nl, n_nl, n_lines, k, dict, ix, d_tmp, d_item,
                                       d_err);
  bytes = cub_bytes;
  cub::DeviceScan::ExclusiveScan(d_cub, bytes, d_item, d_excl, PlanAdd(), PlanItem(), (int)n_lines, s);
  PlaceKernel<<<grid, 128, 0, s>>>(d_tmp, d_item, d_excl, n_lines, d_planned, d_multi, d_totals);
}

}  // namespace wsr
