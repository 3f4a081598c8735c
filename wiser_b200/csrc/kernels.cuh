// Device-side structures shared by kernels.cu and wsr_capi.cu.
#ifndef WSR_KERNELS_CUH
#define WSR_KERNELS_CUH
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wsr.h"

namespace wsr {

#ifndef WSR_WARPS_PER_CTA
#define WSR_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = WSR_WARPS_PER_CTA;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kMaxFastK = 32;      // top-k held one entry per lane
#ifndef WSR_UNIT_BLOCKS
#define WSR_UNIT_BLOCKS 128
#endif
constexpr int kUnitBlocks = WSR_UNIT_BLOCKS;      // max driver-list blocks per warp work unit
// a unit's budget in blocks, driver + the probe-list blocks they span: unit_blocks =
// clamp(kUnitBudget / (1 + ratio), 1, kUnitBlocks), ratio = probe blocks per driver block
#ifndef WSR_UNIT_BUDGET_MULT
#define WSR_UNIT_BUDGET_MULT 4
#endif
constexpr int kUnitBudget = WSR_UNIT_BUDGET_MULT * WSR_UNIT_BLOCKS;

// Read-only view of the HBM-resident index (layout: host_index.h).
struct DevIndexView {
  const uint4 *payload;       // 16 B granules
  const uint4 *blk_info;      // {base_doc, payload_off16, bits, max_tfn}
  const uint32_t *blk_last;
  const uint4 *blk_heads;     // 8 x u16 per block: first doc (minus base) of records 0,4,..,28
  const uint4 *lists;         // {first_block, n_blocks, df_shard, df_global}
  const uint8_t *norms;
  const double *cache;        // 256 entries
  const double *idf;          // per term, calc_es_idf(N_global, df_global)
  const float *blk_max;       // SoA copy of blk_info.w
  const uint32_t *filters;    // per-list doc-range-partitioned Bloom filters
  const uint2 *list_flt;      // per term {first filter word, shift g (0xFFFFFFFF: no filter)}
  const void *positions;      // optional: in-document token positions, postings back to back; u16 entries
                              // when pos16 (every position of the shard < 65536), else u32
  const uint32_t *blk_pos;    // optional: index into positions[] of each block's first posting
  const uint16_t *grp_pos;    // optional: 8 per block, positions before record 4g of the block
  uint32_t pos16;
  uint32_t n_terms;
  uint32_t n_docs;
  uint32_t doc_lo;            // first doc id held by this shard (filter origin)
  uint32_t n_filter_words;    // length of filters[] (bound for look-ahead prefetches)
  uint32_t merge_ratio_x4;    // planner rule for the two-term merge path (UseMergePath)
  // K1 (whole-index decode): the payload cut into kDecodeStageBytes stages; k1_stage_first[s] = first
  // block whose payload starts in stage s (k1_stages + 1 entries)
  const uint32_t *k1_stage_first;
  uint32_t k1_stages;
  uint32_t payload_granules;  // 16-byte granules of payload[] (tail padding included)
};

// One planned query. unit_begin = index of its first work unit in its class queue.
struct DevQuery {
  uint32_t term[WSR_MAX_TERMS];  // query order
  uint8_t n_terms;
  uint8_t flags;                 // kQueryPhrase | kQueryMerge
  uint16_t unit_blocks;          // driver-list blocks per work unit of this query
  uint32_t k;
  uint32_t unit_begin;
  uint32_t n_units;
  uint32_t driver;               // position (query order) of the shortest list
  uint32_t out_slot;             // index of the query in the caller's batch
  uint32_t seg_begin;            // collect mode: first entry of its match segment
  uint32_t cand_begin;           // multi-unit queries: first candidate slot (in units)
};
static_assert(sizeof(DevQuery) == 64, "DevQuery is 64 bytes");
constexpr uint8_t kQueryPhrase = 1;   // terms must occur at consecutive positions, in query order
constexpr uint8_t kQueryMerge = 2;    // two-term query whose lists are of similar length: merge path
// Merge path when partner blocks <= merge_ratio_x4/4 x driver blocks (DevIndexView::merge_ratio_x4,
// set from WSR_MERGE_RATIO_X4 at wsr_index_open). OFF by default: measured on B200 (C2, two-term
// log) the merge kernel needs ~460 warp instructions per partner block against ~590 per DRIVER
// block on the probe path, so it loses even on exactly balanced lists (profiles/r2_notes.md).
constexpr uint32_t kMergeRatioX4 = 0;
__host__ __device__ inline bool UseMergePath(uint32_t n_terms, bool phrase, unsigned long long drv_blocks,
                                             unsigned long long probe_blocks, uint32_t ratio_x4) {
  return n_terms == 2 && !phrase && ratio_x4 != 0 && probe_blocks * 4ull <= drv_blocks * (unsigned long long)ratio_x4;
}

struct DevCounters {
  unsigned long long decoded_postings;
  unsigned long long touched_bytes;
  unsigned long long matches;
  unsigned long long units;
  unsigned long long probe_blocks;
  unsigned int next_unit[5];      // one dynamic queue per query class; [4]: the two-term merge kernel
  unsigned int pad_;
};

struct BatchView {
  const DevQuery *queries;     // class-sorted; class c occupies [class_begin[c], class_begin[c+1])
  const uint32_t *unit_query;  // work unit -> planned query; class c's units start at class_unit_base[c]
  uint32_t class_unit_base[4];
  uint32_t class_begin[5];
  uint32_t class_units[4];
  uint32_t merge_units;        // units of the two-term class that take the merge path
  wsr_hit *hits;               // n * k_stride
  int32_t *n_hits;             // n
  wsr_hit *cand;               // kMaxFastK entries per unit of every multi-unit query
  int32_t *cand_n;             // one per such unit
  unsigned long long *thr;     // per planned query: best known k-th score (double bits)
  DevCounters *counters;
  uint32_t k_stride;
  uint32_t doc_base;           // added to every emitted doc id (document-partitioned shards)
  // collect mode (k > kMaxFastK): every match is appended to the query's segment
  int32_t *seg_doc;
  double *seg_score;
  uint32_t *seg_count;         // per planned query
};

// class ids
enum { kClassOne = 0, kClassTwo = 1, kClassMany = 2, kClassCollect = 3 };

// count_work: instantiate the kernels with the work counters (decoded postings, touched bytes,
// matches) switched on — the profiling pass; ordinary runs do not pay for the bookkeeping.
void LaunchSearchClass(const DevIndexView &ix, const BatchView &b, int cls, int sm_count,
                       cudaStream_t s, bool count_work);
// Fills unit_query[] from the planned queries (one thread per query writes its n_units entries).
void LaunchUnitMap(const BatchView &b, uint32_t *unit_query, uint32_t n_planned, cudaStream_t s);
void LaunchMerge(const BatchView &b, const uint32_t *multi_queries, uint32_t n_multi,
                 cudaStream_t s);
void LaunchDecodeList(const DevIndexView &ix, uint32_t first_block, uint32_t n_blocks,
                      uint32_t *docs, uint32_t *tfs, cudaStream_t s);
void LaunchDecodeAll(const DevIndexView &ix, uint32_t n_blocks, unsigned long long *checksum,
                     int sm_count, cudaStream_t s);
void LaunchRefreshBlockMax(const DevIndexView &ix, uint32_t n_blocks, uint4 *blk_info_rw,
                           float *blk_max_rw, int sm_count, cudaStream_t s);
void LaunchMergeShards(const wsr_hit *gathered, const int32_t *gathered_n, int n_shards,
                       int n_queries, int k_stride, wsr_hit *out, int32_t *out_n,
                       cudaStream_t s);
// doc_freqs of n queries reported by n_shards partitions (g_df[shard][n][WSR_MAX_TERMS],
// g_ndf[shard][n]) -> element-wise maximum.
void LaunchMergeDocFreqs(const uint32_t *g_df, const int32_t *g_ndf, int n_shards, int n, uint32_t *out_df,
                         int32_t *out_ndf, cudaStream_t s);
// Collect mode epilogue: per query segment sort by (score desc, doc asc) and copy the first k.
// seg_begin/seg_end: device arrays [n_collect]; tmp buffers sized like the segment arrays.
size_t CollectSortTempBytes(uint32_t n_entries, uint32_t n_collect);
void LaunchCollectFinish(const BatchView &b, uint32_t n_collect, uint32_t n_entries,
                         uint32_t *seg_begin, uint32_t *seg_end, int32_t *tmp_doc,
                         double *tmp_score, void *cub_tmp, size_t cub_tmp_bytes, cudaStream_t s);

// ---- device front end of the query-log path (frontend.cu) -------------------------------------
// The host TermDict's open-addressing table in HBM: slot = {term id (0xFFFFFFFF empty), upper 32
// bits of the term's hash}; term bytes in arena[term_off[id] .. term_off[id + 1]).
struct DevDict {
  const uint2 *slots;
  uint32_t mask;
  const uint32_t *term_off;
  const char *arena;
};
// Per-query plan contribution / running sums: v[0..2] queries per class (one, two, many),
// v[3..5] work units per class, v[6] candidate-list units of multi-unit queries, v[7] their number,
// v[8] units of two-term queries that take the merge path.
constexpr int kPlanWords = 9;
struct PlanItem {
  uint32_t v[kPlanWords];
};
size_t FrontEndTempBytes(uint32_t len, uint32_t n_lines);
// *d_count += number of '\n' bytes in d_text[0, len)
void LaunchCountNewlines(const char *d_text, uint32_t len, uint32_t *d_count, cudaStream_t s);
// Enqueues: newline positions -> parse + term lookup + per-query planning -> exclusive scan ->
// class-sorted placement. d_totals receives the batch totals, d_err bit 0: a line with more than
// WSR_MAX_TERMS terms, bit 1: a phrase query on an index without positions.
void LaunchFrontEnd(const char *d_text, uint32_t len, uint32_t n_nl, uint32_t n_lines, uint32_t k,
                    const DevDict &dict, const DevIndexView &ix, uint32_t *d_nl, uint32_t *d_n_nl,
                    DevQuery *d_tmp, PlanItem *d_item, PlanItem *d_excl, DevQuery *d_planned,
                    uint32_t *d_multi, PlanItem *d_totals, uint32_t *d_err, void *d_cub,
                    size_t cub_bytes, cudaStream_t s);

// Planning of a batch whose term ids the host already resolved (wsr_search_batch): per-query
// planning kernel -> exclusive scan -> placement, as in LaunchFrontEnd. d_err bit 0: more than
// WSR_MAX_TERMS terms, bit 1: phrase query without positions, bit 2: k > k_stride or bad term id.
void LaunchPlanQueries(const wsr_query *d_in, uint32_t n, uint32_t k_stride, const DevIndexView &ix,
                       DevQuery *d_tmp, PlanItem *d_item, PlanItem *d_excl, DevQuery *d_planned,
                       uint32_t *d_multi, PlanItem *d_totals, uint32_t *d_err, void *d_cub, size_t cub_bytes,
                       cudaStream_t s);

// doc_freqs of a batch planned by LaunchFrontEnd / LaunchPlanQueries (reads its d_tmp):
// d_doc_freqs[n * WSR_MAX_TERMS], d_n_doc_freqs[n].
void LaunchDocFreqs(const DevQuery *d_tmp, uint32_t n, const DevIndexView &ix, uint32_t *d_doc_freqs,
                    int32_t *d_n_doc_freqs, cudaStream_t s);

// Dense packing of a batch's results for the D2H copy: offsets = exclusive sum of n_hits over
// n + 1 slots (d_off[n] = total), then packed[off[q] + j] = hits[q*k + j] for j < n_hits[q].
size_t PackTempBytes(uint32_t n);
void LaunchResultOffsets(const int32_t *d_n_hits, int32_t *d_off, uint32_t n, void *d_cub, size_t cub_bytes,
                         cudaStream_t s);
void LaunchPackResults(const wsr_hit *d_hits, const int32_t *d_n_hits, const int32_t *d_off, uint32_t n,
                       uint32_t k, wsr_hit *d_packed, cudaStream_t s);

}  // namespace wsr
#endif
