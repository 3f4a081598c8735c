// C ABI of libwsr.so (include/wsr.h): index residency in HBM, the batched query scheduler
// (host-side planning of warp work units + device queues), and result marshalling.
// There is no CPU execution path: every search runs the CUDA kernels in kernels.cu, and every
// entry point fails with WSR_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/wsr.h"
#include "host_index.h"
#include "kernels.cuh"

using namespace wsr;

namespace {

thread_local std::string g_err;

int Fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return Fail(WSR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t Ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
};

template <typename T>
struct PinnedBuf {
  T *p = nullptr;
  size_t cap = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  cudaError_t Ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
};

// Persistent host worker pool: ParallelFor(T, fn) runs fn(t, T) for t in [0, T) with the caller
// as thread 0. Planning and log parsing call it several times per batch, so threads are kept.
class WorkerPool {
 public:
  static WorkerPool &Get() {
    static WorkerPool *p = new WorkerPool();   // intentionally leaked: no teardown-order issues
    return *p;
  }
  int Size() const { return (int)workers_.size() + 1; }
  void Run(int T, const std::function<void(int, int)> &fn) {
    T = std::max(1, std::min(T, Size()));
    if (T == 1) { fn(0, 1); return; }
    std::lock_guard<std::mutex> serial(run_mu_);      // one parallel region at a time
    {
      std::lock_guard<std::mutex> g(mu_);
      fn_ = &fn;
      T_ = T;
      pending_ = T - 1;
      epoch_++;
    }
    cv_.notify_all();
    fn(0, T);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&]() { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  WorkerPool() {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n = (int)std::min(16u, hw) - 1;
    for (int i = 0; i < n; i++) workers_.emplace_back([this, i]() { Loop(i + 1); });
    for (auto &w : workers_) w.detach();
  }
  void Loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int, int)> *fn;
      int T;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return epoch_ != seen; });
        seen = epoch_;
        fn = fn_;
        T = T_;
      }
      if (id < T && fn) {
        (*fn)(id, T);
        std::lock_guard<std::mutex> g(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)> *fn_ = nullptr;
  int T_ = 1, pending_ = 0;
  uint64_t epoch_ = 0;
};

template <typename F>
void ParallelFor(int T, F fn) {
  WorkerPool::Get().Run(T, std::function<void(int, int)>(fn));
}
int HostThreads(size_t items, size_t per_thread) {
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)WorkerPool::Get().Size(), items / per_thread));
}

bool IsPinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

struct wsr_index {
  HostIndex host;                 // dictionary + per-list metadata stay on the host
  std::vector<double> idf;        // calc_es_idf per term (scoring.h:21-25), libm log on the host
  int device = 0;
  int sm_count = 148;
  DevBuf<uint4> d_payload;
  DevBuf<uint4> d_blk_info;
  DevBuf<uint32_t> d_blk_last;
  DevBuf<uint4> d_blk_heads;
  DevBuf<uint4> d_lists;
  DevBuf<uint8_t> d_norms;
  DevBuf<double> d_cache;
  DevBuf<double> d_idf;
  DevBuf<float> d_blk_max;
  DevBuf<uint32_t> d_filters;
  DevBuf<uint2> d_list_flt;
  DevBuf<uint32_t> d_k1_first;    // K1 stage table (DevIndexView::k1_stage_first)
  DevBuf<uint32_t> d_positions, d_blk_pos;   // d_positions holds u16 entries when view.pos16
  DevBuf<uint16_t> d_grp_pos;
  // device copy of the term dictionary for the query-log front end (frontend.cu)
  DevBuf<uint2> d_dict_slots;
  DevBuf<uint32_t> d_term_off;
  DevBuf<char> d_arena;
  DevDict dict = {nullptr, 0, nullptr, nullptr};
  bool dict_on_device = false;
  std::atomic<int> result_fill_ppm{1000000};   // hits / (n*k) of the last wsr_search_log, parts per million
  DevIndexView view;
  int64_t n_blocks = 0, payload_bytes = 0, hbm_bytes = 0;
  uint32_t doc_base = 0;          // global id of this partition's doc 0
  std::mutex pool_mu;
  std::vector<wsr_batch *> pool;  // reusable batches for wsr_search / wsr_search_batch
};

struct wsr_batch {
  wsr_index *idx = nullptr;
  int n = 0;
  int k_stride = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;       // doc_freqs of a log travel here, under the search kernels
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host plan
  std::vector<DevQuery> planned;
  std::vector<DevQuery> tmp;         // per input query, before class placement
  std::vector<uint8_t> tmp_cls;
  std::vector<uint32_t> multi;       // planned indices of multi-unit queries
  uint32_t class_begin[5] = {0, 0, 0, 0, 0};
  uint32_t class_units[4] = {0, 0, 0, 0};
  uint32_t np = 0, n_multi = 0;      // planned queries / multi-unit queries (either planner)
  uint32_t n_cand_units = 0, n_seg_entries = 0, n_collect = 0;
  uint32_t merge_units = 0;          // units of the two-term class on the merge path
  uint64_t listed_postings = 0, listed_bytes = 0;
  uint32_t launches = 0;
  // device
  DevBuf<DevQuery> d_queries;
  DevBuf<uint32_t> d_unit_query;
  // results: ONE allocation [hits n*k_stride | n_hits n | DevCounters] so that one memset clears
  // counts + counters and one D2H brings hits + counts back
  DevBuf<uint8_t> d_out;
  wsr_hit *out_hits = nullptr;
  int32_t *out_n = nullptr;
  DevCounters *out_cnt = nullptr;
  size_t out_hits_bytes = 0, out_n_bytes = 0;
  DevBuf<wsr_hit> d_cand;
  DevBuf<int32_t> d_cand_n;
  DevBuf<unsigned long long> d_thr;
  DevBuf<uint32_t> d_multi;
  DevBuf<int32_t> d_seg_doc, d_seg_doc_tmp;
  DevBuf<double> d_seg_score, d_seg_score_tmp;
  DevBuf<uint32_t> d_seg_count, d_seg_begin, d_seg_end;
  DevBuf<uint8_t> d_cub_tmp;
  size_t cub_tmp_bytes = 0;
  // device front end (wsr_search_log): log text, newline positions, unplaced queries, scan
  DevBuf<char> d_text;
  DevBuf<uint32_t> d_nl, d_fe_small;   // d_fe_small: [0] newline count (unused), [1] error bits
  DevBuf<DevQuery> d_tmp;
  DevBuf<wsr_query> d_wq;              // wsr_search_batch: the caller's queries, planned on the GPU
  PinnedBuf<wsr_query> h_wq;
  DevBuf<PlanItem> d_item, d_excl, d_totals;
  DevBuf<uint8_t> d_fe_cub;
  PinnedBuf<char> h_text;
  // packed results of the log path (frontend.cu PackResultsKernel)
  DevBuf<int32_t> d_off;
  DevBuf<wsr_hit> d_packed;
  DevBuf<uint8_t> d_pack_cub;
  PinnedBuf<wsr_hit> h_packed;
  PinnedBuf<int32_t> h_n;
  PinnedBuf<uint32_t> h_totals;        // PlanItem (8 words) + error bits
  DevBuf<uint32_t> d_df;               // doc_freqs of a log planned on the device
  DevBuf<int32_t> d_ndf;
  PinnedBuf<uint32_t> h_df;
  PinnedBuf<int32_t> h_ndf;
  // pinned staging for the host-buffer API
  PinnedBuf<DevQuery> h_queries;
  PinnedBuf<uint8_t> h_out;            // same layout as d_out (hits | n_hits)
  PinnedBuf<uint32_t> h_multi;
  BatchView view;
};

namespace {

// Guided scheduling: a class's unit queue is drained in log order, so the queries near the end of
// a batch get smaller units and the persistent grid's tail is made of short units (a full
// 128-block unit is ~0.35 ms of one warp's time). Same rule in PlanBatch (host) and PlanOne (device).
static inline unsigned long long UnitCapAt(unsigned long long i, unsigned long long n, unsigned long long cap) {
  // last eighth of the batch: a quarter of the size (last quarter, two levels, an eighth of the
  // size and 192-block units with two levels all measured 0.5 - 0.8 % slower)
  const unsigned long long c = 8 * i >= 7 * n ? cap / 4 : cap;
  return c < 1 ? 1 : c;
}

// Host-side half of the batch scheduler: validates queries, picks each query's driver list,
// cuts it into warp work units and groups queries into kernel classes. Two parallel passes over
// contiguous query ranges (classify + count, then place) keep the planned order deterministic:
// within a class, queries stay in batch order.
int PlanBatch(wsr_batch *b, const wsr_query *queries, int n, int k_stride) {
  wsr_index *ix = b->idx;
  const uint32_t n_terms_index = (uint32_t)ix->host.lists.size();
  b->n = n;
  b->k_stride = k_stride;
  b->multi.clear();
  const int T = HostThreads((size_t)n, 1024);
  struct Part {
    uint32_t count[4] = {0, 0, 0, 0}, units[4] = {0, 0, 0, 0};
    uint32_t cand = 0, multi = 0, merge_units = 0;
    uint64_t seg = 0, listed = 0, listed_bytes = 0;
    int err = 0;
  };
  std::vector<Part> part(T);
  b->tmp.resize((size_t)n);
  b->tmp_cls.resize((size_t)n);
  // Small batches (the Search() serving path) cannot fill the GPU with full-size units: one warp
  // would walk a long list alone while thousands idle, and the query's latency is that walk.
  // Cap the unit size so that the batch yields about one unit per resident warp.
  uint64_t unit_cap = kUnitBlocks;
  if (n <= 1024) {
    uint64_t drv_blocks = 0;
    for (int i = 0; i < n; i++) {
      const wsr_query &q = queries[i];
      if (q.n_terms < 2 || q.n_terms > WSR_MAX_TERMS) continue;
      uint32_t best = 0xffffffffu;
      for (uint32_t t2 = 0; t2 < q.n_terms; t2++)
        if (q.term_ids[t2] < n_terms_index) best = std::min(best, ix->host.lists[q.term_ids[t2]].n_blocks);
      if (best != 0xffffffffu) drv_blocks += best;
    }
    const uint64_t warps = (uint64_t)ix->sm_count * 32;
    unit_cap = std::max<uint64_t>(4, std::min<uint64_t>(kUnitBlocks, drv_blocks / warps + 1));
  }
  auto classify = [&](int t, int TT) {
    Part &p = part[t];
    const int lo = (int)((int64_t)n * t / TT), hi = (int)((int64_t)n * (t + 1) / TT);
    for (int i = lo; i < hi; i++) {
      const wsr_query &q = queries[i];
      b->tmp_cls[i] = 255;
      if (q.n_terms > WSR_MAX_TERMS) { p.err = WSR_ERR_UNSUPPORTED; return; }
      if ((int)q.k > k_stride) { p.err = WSR_ERR_ARG; return; }
      if (q.k == 0 || q.n_terms == 0) continue;            // vacuum_engine.h:206-208
      bool ok = true;
      uint32_t best = 0, best_df = 0xffffffffu;
      for (uint32_t t2 = 0; t2 < q.n_terms; t2++) {
        const uint32_t id = q.term_ids[t2];
        if (id == WSR_TERM_ABSENT) { ok = false; break; }   // vacuum_engine.h:213-215
        if (id >= n_terms_index) { p.err = WSR_ERR_ARG; return; }
        const ListInfo &li = ix->host.lists[id];
        if (li.df_shard == 0) ok = false;                   // nothing of this list on this shard
        if (li.df_shard < best_df) { best_df = li.df_shard; best = t2; }
      }
      if (!ok) continue;
      DevQuery dq;
      memset(&dq, 0, sizeof(dq));
      uint64_t probe_blocks = 0;
      for (uint32_t t2 = 0; t2 < q.n_terms; t2++) {
        const ListInfo &li = ix->host.lists[q.term_ids[t2]];
        dq.term[t2] = q.term_ids[t2];
        p.listed += li.df_shard;
        p.listed_bytes += ix->host.list_alg_bytes[q.term_ids[t2]];
        if (t2 != best) probe_blocks += li.n_blocks;
      }
      dq.n_terms = (uint8_t)q.n_terms;
      dq.flags = (q.flags & 1u) && q.n_terms > 1 ? 1 : 0;   // a one-term "phrase" is a plain query
      if (dq.flags && !ix->host.has_positions) { p.err = WSR_ERR_IO; return; }
      dq.k = q.k;
      dq.driver = best;
      dq.out_slot = (uint32_t)i;
      const ListInfo &drv = ix->host.lists[q.term_ids[best]];
      // Unit size: a unit's work is its driver blocks plus the probe-list blocks they can
      // reach, so skewed queries (long probe lists) get fewer driver blocks per unit.
      const uint64_t ratio = drv.n_blocks ? (probe_blocks + drv.n_blocks - 1) / drv.n_blocks : 0;
      const uint64_t cap_i = UnitCapAt((uint64_t)i, (uint64_t)n, unit_cap);
      const uint32_t ub = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(cap_i, (uint64_t)kUnitBudget / (1 + ratio)));
      dq.unit_blocks = (uint16_t)ub;
      dq.n_units = (drv.n_blocks + ub - 1) / ub;
      const int c = q.k > (uint32_t)kMaxFastK ? kClassCollect
                    : q.n_terms == 1 ? kClassOne : q.n_terms == 2 ? kClassTwo : kClassMany;
      if (c == kClassTwo && UseMergePath(q.n_terms, dq.flags & kQueryPhrase, drv.n_blocks, probe_blocks,
                                         ix->view.merge_ratio_x4))
        dq.flags |= kQueryMerge;
      if (dq.flags & kQueryMerge) p.merge_units += dq.n_units;
      if (c == kClassOne) dq.n_units = 1;   // block-max prepass + selective decode, one warp
      b->tmp[i] = dq;
      b->tmp_cls[i] = (uint8_t)c;
      p.count[c]++;
      p.units[c] += dq.n_units;
      if (c == kClassCollect) p.seg += drv.df_shard;
      else if (dq.n_units > 1) { p.cand += dq.n_units; p.multi++; }
    }
  };
  ParallelFor(T, classify);
  for (const Part &p : part) {
    if (p.err == WSR_ERR_UNSUPPORTED) return Fail(p.err, "query has more than WSR_MAX_TERMS terms");
    if (p.err == WSR_ERR_IO)
      return Fail(WSR_ERR_UNSUPPORTED, "phrase query on an index opened without WSR_OPEN_POSITIONS");
    if (p.err) return Fail(p.err, "query k exceeds k_stride, or term id out of range");
  }
  // exclusive prefixes: class-major, thread-minor
  struct Base { uint32_t pos[4], unit[4], cand, multi, seg; };
  std::vector<Base> base(T);
  uint32_t pos = 0, cand = 0, multi = 0;
  uint64_t seg = 0;
  b->listed_postings = b->listed_bytes = 0;
  b->merge_units = 0;
  for (const Part &p : part) b->merge_units += p.merge_units;
  for (int c = 0; c < 4; c++) {
    b->class_begin[c] = pos;
    uint32_t units = 0;
    for (int t = 0; t < T; t++) {
      base[t].pos[c] = pos;
      base[t].unit[c] = units;
      pos += part[t].count[c];
      units += part[t].units[c];
    }
    b->class_units[c] = units;
  }
  b->class_begin[4] = pos;
  for (int t = 0; t < T; t++) {
    base[t].cand = cand;
    base[t].multi = multi;
    base[t].seg = (uint32_t)seg;
    cand += part[t].cand;
    multi += part[t].multi;
    seg += part[t].seg;
    b->listed_postings += part[t].listed;
    b->listed_bytes += part[t].listed_bytes;
  }
  if (seg > 0x7fffffffull) return Fail(WSR_ERR_UNSUPPORTED, "collect-mode batch too large (cub counts in int)");
  b->planned.resize(pos);
  b->multi.resize(multi);
  auto place = [&](int t, int TT) {
    Base bs = base[t];
    const int lo = (int)((int64_t)n * t / TT), hi = (int)((int64_t)n * (t + 1) / TT);
    for (int i = lo; i < hi; i++) {
      const int c = b->tmp_cls[i];
      if (c == 255) continue;
      DevQuery dq = b->tmp[i];
      dq.unit_begin = bs.unit[c];
      bs.unit[c] += dq.n_units;
      if (c == kClassCollect) {
        dq.seg_begin = bs.seg;
        bs.seg += ix->host.lists[dq.term[dq.driver]].df_shard;
      } else if (dq.n_units > 1) {
        dq.cand_begin = bs.cand;
        bs.cand += dq.n_units;
        b->multi[bs.multi++] = bs.pos[c];
      }
      b->planned[bs.pos[c]++] = dq;
    }
  };
  ParallelFor(T, place);
  b->np = pos;
  b->n_multi = multi;
  b->n_cand_units = cand;
  b->n_seg_entries = (uint32_t)seg;
  b->n_collect = b->class_begin[4] - b->class_begin[kClassCollect];
  return WSR_OK;
}

// Sizes the device buffers of a planned batch (b->n, np, n_multi, n_cand_units, ... are set) and
// fills the kernel view. d_queries / d_multi are (re)allocated only when `plan_on_host`: the device
// planner has already written them.
int PrepareBatch(wsr_batch *b, bool plan_on_host) {
  const size_t np = b->np;
  if (plan_on_host) {
    CU(b->d_queries.Ensure(np + 1));
    CU(b->d_multi.Ensure((size_t)b->n_multi + 1));
  }
  b->out_hits_bytes = ((size_t)b->n * b->k_stride * sizeof(wsr_hit) + 15) / 16 * 16;
  b->out_n_bytes = ((size_t)b->n * 4 + 4 + 15) / 16 * 16;
  CU(b->d_out.Ensure(b->out_hits_bytes + b->out_n_bytes + sizeof(DevCounters) + 16));
  b->out_hits = reinterpret_cast<wsr_hit *>(b->d_out.p);
  b->out_n = reinterpret_cast<int32_t *>(b->d_out.p + b->out_hits_bytes);
  b->out_cnt = reinterpret_cast<DevCounters *>(b->d_out.p + b->out_hits_bytes + b->out_n_bytes);
  CU(b->d_cand.Ensure((size_t)b->n_cand_units * kMaxFastK + 1));
  CU(b->d_cand_n.Ensure((size_t)b->n_cand_units + 1));
  CU(b->d_thr.Ensure(np + 1));
  if (b->n_collect) {
    CU(b->d_seg_doc.Ensure(b->n_seg_entries + 1));
    CU(b->d_seg_doc_tmp.Ensure(b->n_seg_entries + 1));
    CU(b->d_seg_score.Ensure(b->n_seg_entries + 1));
    CU(b->d_seg_score_tmp.Ensure(b->n_seg_entries + 1));
    CU(b->d_seg_begin.Ensure(b->n_collect + 1));
    CU(b->d_seg_end.Ensure(b->n_collect + 1));
    b->cub_tmp_bytes = CollectSortTempBytes(b->n_seg_entries, b->n_collect);
    CU(b->d_cub_tmp.Ensure(b->cub_tmp_bytes + 16));
  }
  CU(b->d_seg_count.Ensure(np + 1));
  BatchView &v = b->view;
  v.queries = b->d_queries.p;
  {
    uint32_t base = 0;
    for (int c = 0; c < 4; c++) { v.class_unit_base[c] = base; base += b->class_units[c]; }
    CU(b->d_unit_query.Ensure((size_t)base + 1));
    v.unit_query = b->d_unit_query.p;
  }
  for (int c = 0; c < 5; c++) v.class_begin[c] = b->class_begin[c];
  for (int c = 0; c < 4; c++) v.class_units[c] = b->class_units[c];
  v.merge_units = b->merge_units;
  v.hits = b->out_hits;
  v.n_hits = b->out_n;
  v.cand = b->d_cand.p;
  v.cand_n = b->d_cand_n.p;
  v.thr = b->d_thr.p;
  v.counters = b->out_cnt;
  v.k_stride = (uint32_t)b->k_stride;
  v.doc_base = b->idx->doc_base;
  v.seg_doc = b->d_seg_doc.p;
  v.seg_score = b->d_seg_score.p;
  v.seg_count = b->d_seg_count.p;
  return WSR_OK;
}

// Host-planned batch: buffers + H2D of the plan.
int UploadBatch(wsr_batch *b) {
  const size_t np = b->np;
  int rc = PrepareBatch(b, /*plan_on_host=*/true);
  if (rc) return rc;
  CU(b->h_queries.Ensure(np + 1));
  CU(b->h_multi.Ensure((size_t)b->n_multi + 1));
  if (np) memcpy(b->h_queries.p, b->planned.data(), np * sizeof(DevQuery));
  if (b->n_multi) memcpy(b->h_multi.p, b->multi.data(), (size_t)b->n_multi * 4);
  if (np) CU(cudaMemcpyAsync(b->d_queries.p, b->h_queries.p, np * sizeof(DevQuery),
                             cudaMemcpyHostToDevice, b->stream));
  if (b->n_multi)
    CU(cudaMemcpyAsync(b->d_multi.p, b->h_multi.p, (size_t)b->n_multi * 4, cudaMemcpyHostToDevice,
                       b->stream));
  LaunchUnitMap(b->view, b->d_unit_query.p, (uint32_t)np, b->stream);
  CU(cudaGetLastError());
  return WSR_OK;
}

// Enqueues one pass of the batch on its stream: counters reset, search kernels per class,
// unit merge, collect-mode epilogue. No host<->device copies.
int EnqueueRun(wsr_batch *b, cudaEvent_t *ev = nullptr, bool count_work = false) {
  // ev (optional, 6 events): [0] start, [1] after class one, [2] two, [3] many, [4] collect,
  // [5] after merge + collect epilogue
  const size_t np = b->np;
  uint32_t launches = 0;
  CU(cudaMemsetAsync(b->out_n, 0, b->out_n_bytes + sizeof(DevCounters), b->stream));   // counts + counters
  if (b->n_multi) CU(cudaMemsetAsync(b->d_thr.p, 0, np * 8, b->stream));
  if (b->n_collect) CU(cudaMemsetAsync(b->d_seg_count.p, 0, np * 4, b->stream));
  if (ev) CU(cudaEventRecord(ev[0], b->stream));
  for (int c = 0; c < 4; c++) {
    LaunchSearchClass(b->idx->view, b->view, c, b->idx->sm_count, b->stream, count_work);
    if (ev) CU(cudaEventRecord(ev[1 + c], b->stream));
    launches += b->class_units[c] ? 1 : 0;
    if (c == kClassTwo && b->merge_units && b->merge_units < b->class_units[c]) launches++;   // both paths' kernels
  }
  if (b->n_multi) {
    LaunchMerge(b->view, b->d_multi.p, b->n_multi, b->stream);
    launches++;
  }
  if (b->n_collect) {
    LaunchCollectFinish(b->view, b->n_collect, b->n_seg_entries, b->d_seg_begin.p, b->d_seg_end.p,
                        b->d_seg_doc_tmp.p, b->d_seg_score_tmp.p, b->d_cub_tmp.p,
                        b->cub_tmp_bytes, b->stream);
    launches += 2;
  }
  if (ev) CU(cudaEventRecord(ev[5], b->stream));
  b->launches = launches;
  CU(cudaGetLastError());
  return WSR_OK;
}

// A batch whose (re)planning failed must not be runnable: its device plan may be half written.
void InvalidateBatch(wsr_batch *b) {
  b->n = 0;
  b->np = b->n_multi = b->n_cand_units = b->n_seg_entries = b->n_collect = 0;
  b->merge_units = b->view.merge_units = 0;
  for (int c = 0; c < 4; c++) b->class_units[c] = b->view.class_units[c] = 0;
  for (int c = 0; c < 5; c++) b->class_begin[c] = b->view.class_begin[c] = 0;
}

wsr_batch *NewBatch(wsr_index *idx) {
  if (cudaSetDevice(idx->device) != cudaSuccess) return nullptr;
  std::unique_ptr<wsr_batch> b(new wsr_batch);
  b->idx = idx;
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess) return nullptr;
  return b.release();
}

void FreeBatch(wsr_batch *b) {
  if (!b) return;
  cudaSetDevice(b->idx->device);
  if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
  if (b->side) { cudaStreamSynchronize(b->side); cudaStreamDestroy(b->side); }
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  delete b;
}

wsr_batch *AcquirePooled(wsr_index *idx) {
  {
    std::lock_guard<std::mutex> g(idx->pool_mu);
    if (!idx->pool.empty()) {
      wsr_batch *b = idx->pool.back();
      idx->pool.pop_back();
      return b;
    }
  }
  return NewBatch(idx);
}
void ReleasePooled(wsr_index *idx, wsr_batch *b) {
  std::lock_guard<std::mutex> g(idx->pool_mu);
  idx->pool.push_back(b);
}

// Returns a pooled batch when it goes out of scope, after its stream has drained (no copy may
// still target the caller's buffers, and the next user must not race the previous one's work).
struct PooledBatch {
  wsr_index *idx;
  wsr_batch *b;
  PooledBatch(wsr_index *i) : idx(i), b(AcquirePooled(i)) {}
  ~PooledBatch() {
    if (!b) return;
    cudaStreamSynchronize(b->stream);
    if (b->side) cudaStreamSynchronize(b->side);
    // a zero-copy run pointed the kernels at the caller's buffers: never leave those behind
    b->view.hits = b->out_hits;
    b->view.n_hits = b->out_n;
    ReleasePooled(idx, b);
  }
  PooledBatch(const PooledBatch &) = delete;
  PooledBatch &operator=(const PooledBatch &) = delete;
};


}  // namespace

extern "C" {

const char *wsr_last_error(void) { return g_err.c_str(); }

int wsr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void *wsr_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    Fail(WSR_ERR_CUDA, "cudaMallocHost failed");
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void wsr_host_free(void *p) { if (p) cudaFreeHost(p); }

wsr_index *wsr_index_open(const char *vacuum_dir, int device, int shard, int n_shards,
                          int loader_threads, char *err, size_t errlen) {
  return wsr_index_open_ex(vacuum_dir, device, shard, n_shards, loader_threads, 0, err, errlen);
}

wsr_index *wsr_index_open_ex(const char *vacuum_dir, int device, int shard, int n_shards,
                             int loader_threads, unsigned flags, char *err, size_t errlen) {
  auto fail = [&](const std::string &m) -> wsr_index * {
    g_err = m;
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!vacuum_dir) return fail("vacuum_dir is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail("no CUDA device: libwsr has no CPU path");
  }
  if (device < 0 || device >= ndev) return fail("bad device ordinal");
  std::unique_ptr<wsr_index> ix(new wsr_index);
  std::string e;
  if (!LoadVacuumDir(vacuum_dir, shard, n_shards, loader_threads, &ix->host, &e,
                     (flags & WSR_OPEN_POSITIONS) ? kLoadPositions : 0))
    return fail(std::string(vacuum_dir) + ": " + e);
  HostIndex &h = ix->host;
  // idf per term: calc_es_idf(doc_count, doc_freq), scoring.h:21-25 — GLOBAL N and df
  ix->idf.resize(h.lists.size());
  for (size_t t = 0; t < h.lists.size(); t++) {
    const int doc_count = h.n_docs, doc_freq = (int)h.lists[t].df_global;
    ix->idf[t] = log(1 + (doc_count - doc_freq + 0.5) / (doc_freq + 0.5));
  }
  ix->device = device;
  auto cu = [&](cudaError_t c, const char *what) -> bool {
    if (c == cudaSuccess) return true;
    e = std::string(what) + ": " + cudaGetErrorString(c);
    return false;
  };
  cudaDeviceProp prop;
  if (!cu(cudaSetDevice(device), "cudaSetDevice") ||
      !cu(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties"))
    return fail(e);
  ix->sm_count = prop.multiProcessorCount;
  const size_t n_gran = h.payload.size() / 16;
  if (!cu(ix->d_payload.Ensure(n_gran), "cudaMalloc payload") ||
      !cu(ix->d_blk_info.Ensure(h.blk_info.size() + 1), "cudaMalloc blk_info") ||
      !cu(ix->d_blk_last.Ensure(h.blk_last.size() + 1), "cudaMalloc blk_last") ||
      !cu(ix->d_blk_heads.Ensure(h.blk_heads.size() / 8 + 1), "cudaMalloc blk_heads") ||
      !cu(ix->d_lists.Ensure(h.lists.size() + 1), "cudaMalloc lists") ||
      !cu(ix->d_norms.Ensure(h.norms.size() + 1), "cudaMalloc norms") ||
      !cu(ix->d_cache.Ensure(256), "cudaMalloc cache") ||
      !cu(ix->d_idf.Ensure(ix->idf.size() + 1), "cudaMalloc idf") ||
      !cu(ix->d_blk_max.Ensure(h.blk_info.size() + 1), "cudaMalloc blk_max") ||
      !cu(ix->d_filters.Ensure(h.filters.size() + 1), "cudaMalloc filters") ||
      !cu(ix->d_list_flt.Ensure(h.list_flt.size() + 1), "cudaMalloc list_flt"))
    return fail(e);
  std::vector<float> blk_max(h.blk_info.size());
  for (size_t b = 0; b < blk_max.size(); b++) blk_max[b] = h.blk_info[b].max_tfn;
  std::vector<uint2> list_flt(h.list_flt.size());
  for (size_t t = 0; t < list_flt.size(); t++)
    list_flt[t] = make_uint2((uint32_t)h.list_flt[t], (uint32_t)(h.list_flt[t] >> 32));
  if (!cu(cudaMemcpy(ix->d_payload.p, h.payload.data(), n_gran * 16, cudaMemcpyHostToDevice), "H2D payload") ||
      !cu(cudaMemcpy(ix->d_blk_info.p, h.blk_info.data(), h.blk_info.size() * 16, cudaMemcpyHostToDevice), "H2D blk_info") ||
      !cu(cudaMemcpy(ix->d_blk_last.p, h.blk_last.data(), h.blk_last.size() * 4, cudaMemcpyHostToDevice), "H2D blk_last") ||
      !cu(cudaMemcpy(ix->d_blk_heads.p, h.blk_heads.data(), h.blk_heads.size() * 2, cudaMemcpyHostToDevice), "H2D blk_heads") ||
      !cu(cudaMemcpy(ix->d_lists.p, h.lists.data(), h.lists.size() * 16, cudaMemcpyHostToDevice), "H2D lists") ||
      !cu(cudaMemcpy(ix->d_norms.p, h.norms.data(), h.norms.size(), cudaMemcpyHostToDevice), "H2D norms") ||
      !cu(cudaMemcpy(ix->d_cache.p, h.cache, 256 * 8, cudaMemcpyHostToDevice), "H2D cache") ||
      !cu(cudaMemcpy(ix->d_idf.p, ix->idf.data(), ix->idf.size() * 8, cudaMemcpyHostToDevice), "H2D idf") ||
      !cu(cudaMemcpy(ix->d_blk_max.p, blk_max.data(), blk_max.size() * 4, cudaMemcpyHostToDevice), "H2D blk_max") ||
      !cu(cudaMemcpy(ix->d_filters.p, h.filters.data(), h.filters.size() * 4, cudaMemcpyHostToDevice), "H2D filters") ||
      !cu(cudaMemcpy(ix->d_list_flt.p, list_flt.data(), list_flt.size() * 8, cudaMemcpyHostToDevice), "H2D list_flt"))
    return fail(e);
  ix->n_blocks = (int64_t)h.blk_info.size();
  ix->payload_bytes = (int64_t)n_gran * 16;
  ix->hbm_bytes = ix->payload_bytes + ix->n_blocks * 40 + (int64_t)h.lists.size() * 32 +
                  (int64_t)h.filters.size() * 4 + (int64_t)h.norms.size() + 2048;
  DevIndexView &v = ix->view;
  v.payload = ix->d_payload.p;
  v.blk_info = ix->d_blk_info.p;
  v.blk_last = ix->d_blk_last.p;
  v.blk_heads = ix->d_blk_heads.p;
  v.lists = ix->d_lists.p;
  v.norms = ix->d_norms.p;
  v.cache = ix->d_cache.p;
  v.idf = ix->d_idf.p;
  v.blk_max = ix->d_blk_max.p;
  v.filters = ix->d_filters.p;
  v.list_flt = ix->d_list_flt.p;
  v.n_terms = (uint32_t)h.lists.size();
  v.n_docs = (uint32_t)h.n_docs;
  v.doc_lo = (uint32_t)h.doc_lo;
  v.n_filter_words = (uint32_t)h.filters.size();
  {
    // K1 stage table: first block whose payload starts at or after s * kDecodeStageBytes
    const uint64_t pay_bytes = (uint64_t)n_gran * 16;
    const uint32_t n_stages = (uint32_t)((pay_bytes + kDecodeStageBytes - 1) / kDecodeStageBytes);
    std::vector<uint32_t> first((size_t)n_stages + 1, (uint32_t)h.blk_info.size());
    size_t b = 0;
    for (uint32_t st = 0; st < n_stages; st++) {
      const uint64_t lo16 = (uint64_t)st * (kDecodeStageBytes / 16);
      while (b < h.blk_info.size() && h.blk_info[b].payload_off16 < lo16) b++;
      first[st] = (uint32_t)b;
    }
    if (!cu(ix->d_k1_first.Ensure(first.size()), "cudaMalloc k1 stages") ||
        !cu(cudaMemcpy(ix->d_k1_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice), "H2D k1 stages"))
      return fail(e);
    v.k1_stage_first = ix->d_k1_first.p;
    v.k1_stages = n_stages;
    v.payload_granules = (uint32_t)n_gran;
    ix->hbm_bytes += (int64_t)first.size() * 4;
  }
  v.merge_ratio_x4 = kMergeRatioX4;
  if (const char *mr = getenv("WSR_MERGE_RATIO_X4")) v.merge_ratio_x4 = (uint32_t)std::max(0, atoi(mr));
  v.positions = nullptr;
  v.blk_pos = nullptr;
  v.grp_pos = nullptr;
  v.pos16 = 0;
  if (h.has_positions) {
    // 16-bit positions when they all fit (a document would need 65536+ tokens otherwise)
    bool fits16 = true;
    {
      const int T = HostThreads(h.positions.size(), 1 << 20);
      std::vector<uint8_t> big(T, 0);
      ParallelFor(T, [&](int t, int TT) {
        const size_t lo = h.positions.size() * t / TT, hi = h.positions.size() * (t + 1) / TT;
        uint32_t m = 0;
        for (size_t i = lo; i < hi; i++) m |= h.positions[i];
        big[t] = m > 0xFFFFu;
      });
      for (uint8_t x : big) fits16 = fits16 && !x;
    }
    const size_t n_pos = h.positions.size();
    const void *pos_src = h.positions.data();
    std::vector<uint16_t> pos16;
    if (fits16) {
      pos16.resize(n_pos);
      const int T = HostThreads(n_pos, 1 << 20);
      ParallelFor(T, [&](int t, int TT) {
        for (size_t i = n_pos * t / TT, e = n_pos * (t + 1) / TT; i < e; i++) pos16[i] = (uint16_t)h.positions[i];
      });
      std::vector<uint32_t>().swap(h.positions);
      pos_src = pos16.data();
    }
    const size_t pos_bytes = n_pos * (fits16 ? 2 : 4);
    if (!cu(ix->d_positions.Ensure(pos_bytes / 4 + 2), "cudaMalloc positions") ||
        !cu(ix->d_blk_pos.Ensure(h.blk_pos.size() + 1), "cudaMalloc blk_pos") ||
        !cu(ix->d_grp_pos.Ensure(h.grp_pos.size() + 1), "cudaMalloc grp_pos") ||
        !cu(cudaMemcpy(ix->d_grp_pos.p, h.grp_pos.data(), h.grp_pos.size() * 2, cudaMemcpyHostToDevice), "H2D grp_pos") ||
        !cu(cudaMemcpy(ix->d_positions.p, pos_src, pos_bytes, cudaMemcpyHostToDevice), "H2D positions") ||
        !cu(cudaMemcpy(ix->d_blk_pos.p, h.blk_pos.data(), h.blk_pos.size() * 4, cudaMemcpyHostToDevice), "H2D blk_pos"))
      return fail(e);
    v.positions = ix->d_positions.p;
    v.blk_pos = ix->d_blk_pos.p;
    v.grp_pos = ix->d_grp_pos.p;
    v.pos16 = fits16 ? 1u : 0u;
    ix->hbm_bytes += (int64_t)pos_bytes + (int64_t)h.blk_pos.size() * 4 + (int64_t)h.grp_pos.size() * 2;
    std::vector<uint32_t>().swap(h.positions);
    std::vector<uint32_t>().swap(h.blk_pos);
    std::vector<uint16_t>().swap(h.grp_pos);
  }
  // term dictionary in HBM for the query-log front end: the host table's slots with a 32-bit tag
  // (so most probes never touch the term bytes), 32-bit term offsets, the term arena
  if (h.term_arena.size() < 0xfffffff0ull && h.dict.Mask() < 0xffffffffull && !h.lists.empty()) {
    const std::vector<uint32_t> &slots = h.dict.Slots();
    std::vector<uint2> ds(slots.size());
    const int T = HostThreads(slots.size(), 1 << 16);
    ParallelFor(T, [&](int t, int TT) {
      const size_t lo = slots.size() * t / TT, hi = slots.size() * (t + 1) / TT;
      for (size_t i = lo; i < hi; i++) {
        const uint32_t id = slots[i];
        uint32_t tag = 0;
        if (id != 0xFFFFFFFFu)
          tag = (uint32_t)(TermDict::Hash(h.term_arena.data() + h.term_off[id],
                                          h.term_off[id + 1] - h.term_off[id]) >> 32);
        ds[i] = make_uint2(id, tag);
      }
    });
    std::vector<uint32_t> off32(h.term_off.begin(), h.term_off.end());
    if (!cu(ix->d_dict_slots.Ensure(ds.size()), "cudaMalloc dict") ||
        !cu(ix->d_term_off.Ensure(off32.size()), "cudaMalloc term_off") ||
        !cu(ix->d_arena.Ensure(h.term_arena.size() + 1), "cudaMalloc term arena") ||
        !cu(cudaMemcpy(ix->d_dict_slots.p, ds.data(), ds.size() * 8, cudaMemcpyHostToDevice), "H2D dict") ||
        !cu(cudaMemcpy(ix->d_term_off.p, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice), "H2D term_off") ||
        !cu(cudaMemcpy(ix->d_arena.p, h.term_arena.data(), h.term_arena.size(), cudaMemcpyHostToDevice), "H2D arena"))
      return fail(e);
    ix->dict.slots = ix->d_dict_slots.p;
    ix->dict.mask = (uint32_t)h.dict.Mask();
    ix->dict.term_off = ix->d_term_off.p;
    ix->dict.arena = ix->d_arena.p;
    ix->dict_on_device = true;
    ix->hbm_bytes += (int64_t)ds.size() * 8 + (int64_t)off32.size() * 4 + (int64_t)h.term_arena.size();
  }
  // the block arrays now live in HBM only
  std::vector<uint8_t>().swap(h.payload);
  std::vector<BlockInfo>().swap(h.blk_info);
  std::vector<uint32_t>().swap(h.blk_last);
  std::vector<uint16_t>().swap(h.blk_heads);
  std::vector<uint32_t>().swap(h.filters);
  return ix.release();
}

void wsr_index_close(wsr_index *idx) {
  if (!idx) return;
  cudaSetDevice(idx->device);
  for (wsr_batch *b : idx->pool) FreeBatch(b);
  delete idx;
}

int wsr_index_get_info(const wsr_index *idx, wsr_index_info *info) {
  if (!idx || !info) return Fail(WSR_ERR_ARG, "null argument");
  const HostIndex &h = idx->host;
  info->n_docs = h.n_docs;
  info->avg_doc_len = h.avg_len;
  info->n_terms = (int64_t)h.lists.size();
  info->n_postings = h.n_postings;
  info->n_postings_global = h.n_postings_global;
  info->n_blocks = idx->n_blocks;
  info->hbm_bytes = idx->hbm_bytes;
  info->payload_bytes = idx->payload_bytes;
  info->shard = h.shard;
  info->n_shards = h.n_shards;
  info->doc_lo = h.doc_lo;
  info->doc_hi = h.doc_hi;
  info->device = idx->device;
  return WSR_OK;
}

int wsr_index_set_global_stats(wsr_index *idx, int64_t doc_base, int64_t n_docs_global,
                               double avg_len_global, const uint32_t *df_global) {
  if (!idx || !df_global || doc_base < 0 || n_docs_global <= 0 || !(avg_len_global > 0) ||
      doc_base + (int64_t)idx->host.norms.size() > 0x7fffffffll)
    return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(idx->device));
  HostIndex &h = idx->host;
  h.n_docs = (int32_t)n_docs_global;
  h.avg_len = avg_len_global;
  const double k1 = 1.2, b = 0.75;                         // Bm25Similarity::BuildCache, scoring.h:85-90
  for (int i = 0; i < 256; i++) {
    const uint32_t m = i & 7;
    const int sh = (i >> 3) - 1;
    const uint32_t len = sh < 0 ? m : (m | 8u) << sh;
    h.cache[i] = k1 * (1 - b + b * len / h.avg_len);
  }
  h.n_postings_global = 0;
  for (size_t t = 0; t < h.lists.size(); t++) {
    if (df_global[t] < h.lists[t].df_shard) return Fail(WSR_ERR_ARG, "global df below the local df");
    h.lists[t].df_global = df_global[t];
    h.n_postings_global += df_global[t];
    const int doc_count = h.n_docs, doc_freq = (int)df_global[t];
    idx->idf[t] = log(1 + (doc_count - doc_freq + 0.5) / (doc_freq + 0.5));   // scoring.h:21-25
  }
  idx->doc_base = (uint32_t)doc_base;
  CU(cudaMemcpy(idx->d_lists.p, h.lists.data(), h.lists.size() * 16, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(idx->d_cache.p, h.cache, 256 * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(idx->d_idf.p, idx->idf.data(), idx->idf.size() * 8, cudaMemcpyHostToDevice));
  idx->view.n_docs = (uint32_t)h.n_docs;
  LaunchRefreshBlockMax(idx->view, (uint32_t)idx->n_blocks, idx->d_blk_info.p, idx->d_blk_max.p,
                        idx->sm_count, 0);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return WSR_OK;
}

int wsr_index_local_stats(const wsr_index *idx, uint32_t *df_local, uint32_t *ranks) {
  if (!idx) return Fail(WSR_ERR_ARG, "null argument");
  const HostIndex &h = idx->host;
  for (size_t t = 0; t < h.lists.size(); t++) {
    if (df_local) df_local[t] = h.lists[t].df_shard;
    if (ranks) {
      const char *s = h.term_arena.data() + h.term_off[t];
      const size_t len = h.term_off[t + 1] - h.term_off[t];
      if (len < 2 || len > 10 || s[0] != 't') return Fail(WSR_ERR_ARG, "term is not named t<rank>");
      uint32_t r = 0;
      for (size_t i = 1; i < len; i++) {
        if (s[i] < '0' || s[i] > '9') return Fail(WSR_ERR_ARG, "term is not named t<rank>");
        r = r * 10 + (uint32_t)(s[i] - '0');
      }
      ranks[t] = r;
    }
  }
  return WSR_OK;
}

int wsr_term_lookup(const wsr_index *idx, const char *term, size_t len, uint32_t *term_id,
                    uint32_t *df) {
  if (!idx || !term) return Fail(WSR_ERR_ARG, "null argument");
  const uint32_t t = idx->host.dict.Find(term, len);
  if (term_id) *term_id = t;
  if (t == WSR_TERM_ABSENT) {
    if (df) *df = 0;
    return 1;
  }
  if (df) *df = idx->host.lists[t].df_global;
  return 0;
}

int wsr_term_at(const wsr_index *idx, uint32_t term_id, char *buf, size_t cap, uint32_t *df) {
  if (!idx) return Fail(WSR_ERR_ARG, "null argument");
  const HostIndex &h = idx->host;
  if (term_id >= h.lists.size()) return Fail(WSR_ERR_ARG, "term id out of range");
  const size_t a = h.term_off[term_id], b = h.term_off[term_id + 1];
  if (buf) memcpy(buf, h.term_arena.data() + a, std::min(cap, b - a));
  if (df) *df = h.lists[term_id].df_global;
  return (int)(b - a);
}

int wsr_decode_list(wsr_index *idx, uint32_t term_id, uint32_t *docs, uint32_t *tfs, size_t cap,
                    size_t *n) {
  if (!idx || !n) return Fail(WSR_ERR_ARG, "null argument");
  if (term_id >= idx->host.lists.size()) return Fail(WSR_ERR_ARG, "term id out of range");
  CU(cudaSetDevice(idx->device));
  const ListInfo li = idx->host.lists[term_id];
  *n = li.df_shard;
  if (li.n_blocks == 0 || cap == 0 || (!docs && !tfs)) return WSR_OK;
  DevBuf<uint32_t> d_docs, d_tfs;
  CU(d_docs.Ensure((size_t)li.n_blocks * 128));
  CU(d_tfs.Ensure((size_t)li.n_blocks * 128));
  LaunchDecodeList(idx->view, li.first_block, li.n_blocks, d_docs.p, d_tfs.p, 0);
  CU(cudaGetLastError());
  const size_t m = std::min<size_t>(cap, li.df_shard);
  if (docs && m) CU(cudaMemcpy(docs, d_docs.p, m * 4, cudaMemcpyDeviceToHost));
  if (tfs && m) CU(cudaMemcpy(tfs, d_tfs.p, m * 4, cudaMemcpyDeviceToHost));
  CU(cudaDeviceSynchronize());
  return WSR_OK;
}

int wsr_decode_all(wsr_index *idx, uint64_t *checksum, float *kernel_ms) {
  if (!idx) return Fail(WSR_ERR_ARG, "null argument");
  CU(cudaSetDevice(idx->device));
  DevBuf<unsigned long long> d_sum;
  CU(d_sum.Ensure(1));
  CU(cudaMemset(d_sum.p, 0, 8));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, 0));
  LaunchDecodeAll(idx->view, (uint32_t)idx->n_blocks, d_sum.p, idx->sm_count, 0);
  CU(cudaEventRecord(e1, 0));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  unsigned long long s = 0;
  CU(cudaMemcpy(&s, d_sum.p, 8, cudaMemcpyDeviceToHost));
  if (checksum) *checksum = s;
  if (kernel_ms) *kernel_ms = ms;
  return WSR_OK;
}

namespace {
bool DeviceFrontEndUsable(const wsr_index *idx, size_t len, int k);
int PlanLogOnDevice(wsr_batch *b, const char *text, size_t len, int k, int cap_q);
int PlanQueriesOnDevice(wsr_batch *b, const wsr_query *queries, int n, int k_stride);
int FinishDevicePlan(wsr_batch *b, uint32_t n, int k_stride);
}  // namespace

wsr_batch *wsr_batch_create(wsr_index *idx, const wsr_query *queries, int n, int k_stride) {
  if (!idx || (!queries && n > 0) || n < 0 || k_stride < 1) {
    Fail(WSR_ERR_ARG, "bad argument");
    return nullptr;
  }
  wsr_batch *b = NewBatch(idx);
  if (!b) { Fail(WSR_ERR_CUDA, "cannot create CUDA stream"); return nullptr; }
  if (PlanBatch(b, queries, n, k_stride) != WSR_OK || UploadBatch(b) != WSR_OK ||
      cudaStreamSynchronize(b->stream) != cudaSuccess) {
    FreeBatch(b);
    return nullptr;
  }
  return b;
}

int wsr_batch_reset(wsr_batch *b, const wsr_query *queries, int n, int k_stride) {
  if (!b || (!queries && n > 0) || n < 0 || k_stride < 1) return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(b->idx->device));
  // the previous plan's H2D copies read the pinned staging this call is about to overwrite
  CU(cudaStreamSynchronize(b->stream));
  int rc = PlanBatch(b, queries, n, k_stride);
  if (rc == WSR_OK) rc = UploadBatch(b);
  if (rc != WSR_OK) InvalidateBatch(b);
  return rc;
}

int wsr_batch_reset_log(wsr_batch *b, const char *text, size_t len, int k, int *n_queries) {
  if (!b || (!text && len) || k < 1 || !n_queries) return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(b->idx->device));
  CU(cudaStreamSynchronize(b->stream));   // see wsr_batch_reset
  if (DeviceFrontEndUsable(b->idx, len, k)) {
    const int rc = PlanLogOnDevice(b, text, len, k, 0x7fffffff);
    if (rc == WSR_OK) *n_queries = b->n;
    else InvalidateBatch(b);
    return rc;
  }
  size_t lines = 1;
  for (size_t i = 0; i < len; i++) lines += text[i] == '\n';
  std::vector<wsr_query> qs(lines + 1);
  int n = 0;
  int rc = wsr_parse_query_log(b->idx, text, len, k, qs.data(), (int)qs.size(), &n);
  if (rc == WSR_OK) rc = PlanBatch(b, qs.data(), n, k);
  if (rc == WSR_OK) rc = UploadBatch(b);
  if (rc == WSR_OK) *n_queries = n;
  else InvalidateBatch(b);
  return rc;
}

void wsr_batch_destroy(wsr_batch *b) { FreeBatch(b); }

int wsr_batch_run(wsr_batch *b) {
  if (!b) return Fail(WSR_ERR_ARG, "null batch");
  CU(cudaSetDevice(b->idx->device));
  return EnqueueRun(b);
}

int wsr_batch_sync(wsr_batch *b) {
  if (!b) return Fail(WSR_ERR_ARG, "null batch");
  CU(cudaSetDevice(b->idx->device));
  CU(cudaStreamSynchronize(b->stream));
  return WSR_OK;
}

int wsr_batch_fetch(wsr_batch *b, wsr_hit *hits, int32_t *n_hits) {
  if (!b) return Fail(WSR_ERR_ARG, "null batch");
  CU(cudaSetDevice(b->idx->device));
  const size_t nh = (size_t)b->n * b->k_stride;
  const bool direct = hits && n_hits && IsPinned(hits) && IsPinned(n_hits);
  if (direct) {        // DMA straight into the caller's pinned buffers
    if (nh) CU(cudaMemcpyAsync(hits, b->out_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
    if (b->n) CU(cudaMemcpyAsync(n_hits, b->out_n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return WSR_OK;
  }
  // one D2H of [hits | n_hits] into pinned staging, then out to the caller's arrays
  const size_t bytes = b->out_hits_bytes + (size_t)b->n * 4;
  CU(b->h_out.Ensure(bytes + 16));
  if (b->n) CU(cudaMemcpyAsync(b->h_out.p, b->d_out.p, bytes, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  if (hits && nh) memcpy(hits, b->h_out.p, nh * sizeof(wsr_hit));
  if (n_hits && b->n) memcpy(n_hits, b->h_out.p + b->out_hits_bytes, (size_t)b->n * 4);
  return WSR_OK;
}

int wsr_batch_device_results(wsr_batch *b, void **d_hits, void **d_n_hits, void **stream) {
  if (!b) return Fail(WSR_ERR_ARG, "null batch");
  if (d_hits) *d_hits = b->out_hits;
  if (d_n_hits) *d_n_hits = b->out_n;
  if (stream) *stream = (void *)b->stream;
  return WSR_OK;
}

int wsr_batch_time(wsr_batch *b, int iters, float *ms_per_iter) {
  if (!b || iters < 1) return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(b->idx->device));
  CU(cudaStreamSynchronize(b->stream));
  CU(cudaEventRecord(b->ev0, b->stream));
  for (int i = 0; i < iters; i++) {
    int rc = EnqueueRun(b);
    if (rc) return rc;
  }
  CU(cudaEventRecord(b->ev1, b->stream));
  CU(cudaEventSynchronize(b->ev1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  if (ms_per_iter) *ms_per_iter = ms / iters;
  return WSR_OK;
}

int wsr_batch_profile(wsr_batch *b, float ms[6]) {
  if (!b || !ms) return Fail(WSR_ERR_ARG, "null argument");
  CU(cudaSetDevice(b->idx->device));
  struct Events {
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
  } evs;
  cudaEvent_t *ev = evs.ev;
  for (int i = 0; i < 6; i++) CU(cudaEventCreate(&ev[i]));
  int rc = EnqueueRun(b, ev);
  if (rc) return rc;
  CU(cudaStreamSynchronize(b->stream));
  for (int i = 0; i < 5; i++) CU(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  CU(cudaEventElapsedTime(&ms[5], ev[0], ev[5]));
  return WSR_OK;
}

int wsr_batch_count_work(wsr_batch *b) {
  if (!b) return Fail(WSR_ERR_ARG, "null argument");
  CU(cudaSetDevice(b->idx->device));
  int rc = EnqueueRun(b, nullptr, /*count_work=*/true);
  if (rc) return rc;
  CU(cudaStreamSynchronize(b->stream));
  return WSR_OK;
}

namespace {
// Parses the lines of text[begin, end) (begin is a line start) into out[0..]; returns the
// number of queries, or a negative status.
int ParseLines(const wsr_index *idx, const char *text, size_t begin, size_t end, int k,
               wsr_query *out, int cap) {
  int n = 0;
  size_t p = begin;
  while (p < end) {
    size_t e = p;
    while (e < end && text[e] != '\n') e++;
    // utils::trim + phrase quotes (query_pool.h:251-311)
    size_t a = p, b = e;
    while (a < b && isspace((unsigned char)text[a])) a++;
    while (b > a && isspace((unsigned char)text[b - 1])) b--;
    uint32_t flags = 0;
    if (b - a >= 1 && text[a] == '"' && text[b - 1] == '"') {
      flags = 1;
      a++;
      if (b > a) b--;
    }
    if (n >= cap) return WSR_ERR_ARG;
    wsr_query &q = out[n];
    memset(&q, 0, sizeof(q));
    q.k = (uint32_t)k;
    q.flags = flags;
    size_t t = a;
    while (t < b) {                      // utils::explode(line, ' '): empty pieces dropped
      while (t < b && text[t] == ' ') t++;
      size_t u = t;
      while (u < b && text[u] != ' ') u++;
      if (u > t) {
        if (q.n_terms >= WSR_MAX_TERMS) return WSR_ERR_UNSUPPORTED;
        q.term_ids[q.n_terms++] = idx->host.dict.Find(text + t, u - t);
      }
      t = u;
    }
    n++;
    p = e + 1;
  }
  return n;
}
}  // namespace

int wsr_parse_query_log(const wsr_index *idx, const char *text, size_t len, int k,
                        wsr_query *out, int cap, int *n_out) {
  if (!idx || (!text && len) || !out || !n_out || k < 0) return Fail(WSR_ERR_ARG, "bad argument");
  const int T = HostThreads(len, 1 << 14);
  // chunk boundaries on line starts, then per-chunk line counts give the output offsets
  std::vector<size_t> cut(T + 1, len);
  cut[0] = 0;
  for (int t = 1; t < T; t++) {
    size_t p = len * t / T;
    while (p < len && text[p] != '\n') p++;
    cut[t] = p < len ? p + 1 : len;
  }
  std::vector<int> lines(T, 0), got(T, 0);
  ParallelFor(T, [&](int t, int) {
    int c = 0;
    const char *p = text + cut[t], *e = text + cut[t + 1];
    while (p < e) {
      const char *nl = (const char *)memchr(p, '\n', e - p);
      c++;
      if (!nl) break;
      p = nl + 1;
    }
    lines[t] = c;
  });
  std::vector<int> off(T + 1, 0);
  for (int t = 0; t < T; t++) off[t + 1] = off[t] + lines[t];
  if (off[T] > cap) return Fail(WSR_ERR_ARG, "query buffer too small");
  ParallelFor(T, [&](int t, int) {
    got[t] = ParseLines(idx, text, cut[t], cut[t + 1], k, out + off[t], lines[t]);
  });
  for (int t = 0; t < T; t++) {
    if (got[t] == WSR_ERR_UNSUPPORTED) return Fail(WSR_ERR_UNSUPPORTED, "more than WSR_MAX_TERMS terms");
    if (got[t] != lines[t]) return Fail(WSR_ERR_ARG, "query log parse error");
  }
  *n_out = off[T];
  return WSR_OK;
}

int wsr_batch_get_stats(wsr_batch *b, wsr_batch_stats *s) {
  if (!b || !s) return Fail(WSR_ERR_ARG, "null argument");
  CU(cudaSetDevice(b->idx->device));
  DevCounters c;
  CU(cudaMemcpyAsync(&c, b->out_cnt, sizeof(c), cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  s->listed_postings = b->listed_postings;
  s->listed_bytes = b->listed_bytes;
  s->decoded_postings = c.decoded_postings;
  s->touched_bytes = c.touched_bytes;
  s->matches = c.matches;
  s->work_units = c.units;
  s->kernel_launches = b->launches;
  s->reserved = 0;
  s->probe_blocks = c.probe_blocks;
  return WSR_OK;
}

int wsr_search_batch(wsr_index *idx, const wsr_query *queries, int n, int k_stride,
                     wsr_hit *hits, int32_t *n_hits, uint32_t *doc_freqs,
                     int32_t *n_doc_freqs) {
  if (!idx || (!queries && n > 0) || n < 0 || k_stride < 1 || !hits || !n_hits)
    return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(idx->device));
  // doc_freqs: df of term i in query order, empty on the early-out paths (vacuum_engine.h:206-219)
  if (doc_freqs || n_doc_freqs) {
    for (int i = 0; i < n; i++) {
      const wsr_query &q = queries[i];
      bool ok = q.k > 0 && q.n_terms > 0 && q.n_terms <= WSR_MAX_TERMS;
      for (uint32_t t = 0; ok && t < q.n_terms; t++)
        ok = q.term_ids[t] != WSR_TERM_ABSENT && q.term_ids[t] < idx->host.lists.size();
      if (n_doc_freqs) n_doc_freqs[i] = ok ? (int32_t)q.n_terms : 0;
      if (doc_freqs && ok)
        for (uint32_t t = 0; t < q.n_terms; t++)
          doc_freqs[(size_t)i * WSR_MAX_TERMS + t] = idx->host.lists[q.term_ids[t]].df_global;
    }
  }
  PooledBatch pooled(idx);
  wsr_batch *b = pooled.b;
  if (!b) return Fail(WSR_ERR_CUDA, "cannot create batch");
  // large batches without collect-class queries are planned on the GPU, the rest by host threads
  static const bool host_planner = getenv("WSR_HOST_FRONTEND") && atoi(getenv("WSR_HOST_FRONTEND")) != 0;
  int rc;
  if (!host_planner && n >= 8192 && k_stride <= kMaxFastK) {
    rc = PlanQueriesOnDevice(b, queries, n, k_stride);
  } else {
    rc = PlanBatch(b, queries, n, k_stride);
    if (rc == WSR_OK) rc = UploadBatch(b);
  }
  if (rc == WSR_OK) rc = EnqueueRun(b);
  if (rc == WSR_OK) rc = wsr_batch_fetch(b, hits, n_hits);
  return rc;
}

namespace {

// wsr_search_log with the front end on the GPU: the log text is the only input that crosses
// PCIe; parsing, term lookup and planning run as kernels (frontend.cu), the host reads back 36
// bytes of totals to size the candidate buffers and launch the class kernels.
// Plans batch b from log text on the GPU (text H2D, frontend.cu kernels, 36 bytes of totals back)
// and prepares its buffers; the batch is then ready for EnqueueRun.
int PlanLogOnDevice(wsr_batch *b, const char *text, size_t len, int k, int cap_q) {
  wsr_index *idx = b->idx;
  // text -> device (staged through pinned memory unless the caller's buffer already is)
  CU(b->d_text.Ensure(len + 1));
  const char *src = text;
  if (!IsPinned(text)) {
    CU(b->h_text.Ensure(len + 1));
    memcpy(b->h_text.p, text, len);
    src = b->h_text.p;
  }
  CU(cudaMemcpyAsync(b->d_text.p, src, len, cudaMemcpyHostToDevice, b->stream));
  // line count (QueryLogReader semantics: every '\n' ends a line; a last line without one still
  // counts) by a kernel behind the copy: 4 bytes come back before the grids are sized. A host-side
  // count over the worker pool was as fast on an idle box but made the call's latency depend on
  // how quickly sleeping host threads wake up (plan time 0.3 - 1.0 ms in the round-2 traces).
  CU(b->d_fe_small.Ensure(2));
  CU(b->h_totals.Ensure(kPlanWords + 1));
  CU(cudaMemsetAsync(b->d_fe_small.p, 0, 8, b->stream));
  LaunchCountNewlines(b->d_text.p, (uint32_t)len, b->d_fe_small.p, b->stream);
  CU(cudaMemcpyAsync(b->h_totals.p, b->d_fe_small.p, 4, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  const size_t n_nl = b->h_totals.p[0];
  const size_t n_lines = n_nl + (text[len - 1] != '\n' ? 1 : 0);
  if (n_lines > (size_t)cap_q) return Fail(WSR_ERR_ARG, "result buffers too small");
  const uint32_t n = (uint32_t)n_lines;
  CU(b->d_nl.Ensure(n_nl + 1));
  CU(b->d_fe_small.Ensure(2));
  CU(b->d_tmp.Ensure(n));
  CU(b->d_item.Ensure(n));
  CU(b->d_excl.Ensure(n));
  CU(b->d_totals.Ensure(1));
  CU(b->d_queries.Ensure((size_t)n + 1));
  CU(b->d_multi.Ensure((size_t)n + 1));
  CU(b->h_totals.Ensure(kPlanWords + 1));
  const size_t cub_bytes = FrontEndTempBytes((uint32_t)len, n);
  CU(b->d_fe_cub.Ensure(cub_bytes));
  CU(cudaMemsetAsync(b->d_fe_small.p, 0, 8, b->stream));
  LaunchFrontEnd(b->d_text.p, (uint32_t)len, (uint32_t)n_nl, n, (uint32_t)k, idx->dict, idx->view,
                 b->d_nl.p, b->d_fe_small.p, b->d_tmp.p, b->d_item.p, b->d_excl.p, b->d_queries.p,
                 b->d_multi.p, b->d_totals.p, b->d_fe_small.p + 1, b->d_fe_cub.p, cub_bytes, b->stream);
  CU(cudaGetLastError());
  return FinishDevicePlan(b, n, k);
}

// Second half of both device planners: reads the 36 bytes of totals back, turns them into the
// class layout, sizes the batch's buffers and enqueues the unit -> query map.
int FinishDevicePlan(wsr_batch *b, uint32_t n, int k_stride) {
  CU(cudaMemcpyAsync(b->h_totals.p, b->d_totals.p, kPlanWords * 4, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaMemcpyAsync(b->h_totals.p + kPlanWords, b->d_fe_small.p + 1, 4, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  const uint32_t *tot = b->h_totals.p;
  const uint32_t errs = tot[kPlanWords];
  if (errs & 7u) InvalidateBatch(b);   // the plan in d_queries is not usable
  if (errs & 1u) return Fail(WSR_ERR_UNSUPPORTED, "more than WSR_MAX_TERMS terms");
  if (errs & 2u)
    return Fail(WSR_ERR_UNSUPPORTED, "phrase query on an index opened without WSR_OPEN_POSITIONS");
  if (errs & 4u) return Fail(WSR_ERR_ARG, "query k exceeds k_stride, or term id out of range");
  b->n = (int)n;
  b->k_stride = k_stride;
  b->planned.clear();
  b->multi.clear();
  uint32_t pos = 0;
  for (int c = 0; c < 3; c++) {
    b->class_begin[c] = pos;
    pos += tot[c];
    b->class_units[c] = tot[3 + c];
  }
  b->class_begin[3] = b->class_begin[4] = pos;
  b->class_units[3] = 0;
  b->np = pos;
  b->n_multi = tot[7];
  b->n_cand_units = tot[6];
  b->merge_units = tot[8];
  b->n_seg_entries = b->n_collect = 0;
  b->listed_postings = b->listed_bytes = 0;   // not tallied by the device planners
  const int rc = PrepareBatch(b, /*plan_on_host=*/false);
  if (rc) return rc;
  LaunchUnitMap(b->view, b->d_unit_query.p, b->np, b->stream);
  CU(cudaGetLastError());
  return WSR_OK;
}

// wsr_search_batch for large batches with k_stride <= kMaxFastK: the wsr_query array goes to the
// GPU as it is and is planned there (frontend.cu PlanQueriesKernel + scan + placement) instead of
// by host threads.
int PlanQueriesOnDevice(wsr_batch *b, const wsr_query *queries, int n_in, int k_stride) {
  wsr_index *idx = b->idx;
  const uint32_t n = (uint32_t)n_in;
  CU(b->d_wq.Ensure(n));
  const wsr_query *src = queries;
  if (!IsPinned(queries)) {
    CU(b->h_wq.Ensure(n));
    memcpy(b->h_wq.p, queries, (size_t)n * sizeof(wsr_query));
    src = b->h_wq.p;
  }
  CU(cudaMemcpyAsync(b->d_wq.p, src, (size_t)n * sizeof(wsr_query), cudaMemcpyHostToDevice, b->stream));
  CU(b->d_fe_small.Ensure(2));
  CU(b->d_tmp.Ensure(n));
  CU(b->d_item.Ensure(n));
  CU(b->d_excl.Ensure(n));
  CU(b->d_totals.Ensure(1));
  CU(b->d_queries.Ensure((size_t)n + 1));
  CU(b->d_multi.Ensure((size_t)n + 1));
  CU(b->h_totals.Ensure(kPlanWords + 1));
  const size_t cub_bytes = FrontEndTempBytes(16, n);
  CU(b->d_fe_cub.Ensure(cub_bytes));
  CU(cudaMemsetAsync(b->d_fe_small.p, 0, 8, b->stream));
  LaunchPlanQueries(b->d_wq.p, n, (uint32_t)k_stride, idx->view, b->d_tmp.p, b->d_item.p, b->d_excl.p,
                    b->d_queries.p, b->d_multi.p, b->d_totals.p, b->d_fe_small.p + 1, b->d_fe_cub.p, cub_bytes,
                    b->stream);
  CU(cudaGetLastError());
  return FinishDevicePlan(b, n, k_stride);
}

bool DeviceFrontEndUsable(const wsr_index *idx, size_t len, int k) {
  static const bool host_frontend = getenv("WSR_HOST_FRONTEND") && atoi(getenv("WSR_HOST_FRONTEND")) != 0;
  return !host_frontend && idx->dict_on_device && k <= kMaxFastK && len > 0 && len < 0x7ffffff0ull;   // cub counts in int
}

// doc_freqs of the log a batch was just planned from on the device: kernel + D2H on the batch
// stream, into the caller's arrays (through pinned staging unless they are pinned themselves).
int EnqueueDocFreqs(wsr_batch *b, uint32_t *doc_freqs, int32_t *n_doc_freqs, bool *staged) {
  const uint32_t n = (uint32_t)b->n;
  *staged = false;
  if (!n) return WSR_OK;
  CU(b->d_df.Ensure((size_t)n * WSR_MAX_TERMS));
  CU(b->d_ndf.Ensure(n));
  LaunchDocFreqs(b->d_tmp.p, n, b->idx->view, b->d_df.p, b->d_ndf.p, b->stream);
  CU(cudaGetLastError());
  uint32_t *df = doc_freqs;
  int32_t *ndf = n_doc_freqs;
  if (!IsPinned(doc_freqs) || !IsPinned(n_doc_freqs)) {
    CU(b->h_df.Ensure((size_t)n * WSR_MAX_TERMS));
    CU(b->h_ndf.Ensure(n));
    df = b->h_df.p;
    ndf = b->h_ndf.p;
    *staged = true;
  }
  CU(cudaMemcpyAsync(df, b->d_df.p, (size_t)n * WSR_MAX_TERMS * 4, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaMemcpyAsync(ndf, b->d_ndf.p, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
  return WSR_OK;
}

int SearchLogOnDevice(wsr_index *idx, const char *text, size_t len, int k, wsr_hit *hits,
                      int32_t *n_hits, uint32_t *doc_freqs, int32_t *n_doc_freqs, int cap_q,
                      int *n_queries) {
  // WSR_TRACE=1: host-side stage times of this call on stderr (where the end-to-end time goes)
  static const bool trace = getenv("WSR_TRACE") && atoi(getenv("WSR_TRACE")) != 0;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  PooledBatch pooled(idx);
  wsr_batch *b = pooled.b;
  if (!b) return Fail(WSR_ERR_CUDA, "cannot create batch");
  int rc = PlanLogOnDevice(b, text, len, k, cap_q);
  if (rc) return rc;
  const double t_plan = since();
  const uint32_t n = (uint32_t)b->n;
  const size_t nh = (size_t)n * k;
  const bool pinned = IsPinned(hits) && IsPinned(n_hits);
  // Pinned result buffers: the kernels write every query's top-k row and count STRAIGHT into the
  // caller's memory (page-locked host memory is device-addressable under UVA), a row when its query
  // finishes — the 16.4 MB result copy behind the last kernel (0.3 ms per 100k-query log) disappears
  // under the search kernels, and only rows that exist cross PCIe. doc_freqs (if asked for) are
  // written the same way by DocFreqsKernel on a side stream: they depend on the planned queries
  // only, which are complete at this point (the planner's totals have been read back).
  static const bool zero_copy = !(getenv("WSR_NO_ZEROCOPY") && atoi(getenv("WSR_NO_ZEROCOPY")) != 0);
  if (pinned && zero_copy) {
    const bool df_direct = doc_freqs && IsPinned(doc_freqs) && IsPinned(n_doc_freqs);
    bool df_staged = false;
    if (doc_freqs) {
      if (df_direct) {
        if (!b->side) CU(cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking));
        LaunchDocFreqs(b->d_tmp.p, n, idx->view, doc_freqs, n_doc_freqs, b->side);
        CU(cudaGetLastError());
      } else {
        rc = EnqueueDocFreqs(b, doc_freqs, n_doc_freqs, &df_staged);
        if (rc) return rc;
      }
    }
    if (n) memset(n_hits, 0, (size_t)n * 4);   // queries without work units never write their count
    b->view.hits = hits;
    b->view.n_hits = n_hits;
    rc = EnqueueRun(b);
    if (rc) return rc;
    const double t_enq = since();
    CU(cudaStreamSynchronize(b->stream));
    if (b->side) CU(cudaStreamSynchronize(b->side));
    const double t_done = since();
    b->view.hits = b->out_hits;
    b->view.n_hits = b->out_n;
    size_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += (size_t)n_hits[i];
    if (trace)
      fprintf(stderr, "[wsr trace] search_log n=%u: plan (H2D + front end + totals) %.0f us, enqueue %.0f us, "
                      "kernels + result stores %.0f us, tail %.0f us\n",
              n, t_plan, t_enq - t_plan, t_done - t_enq, since() - t_done);
    if (nh) idx->result_fill_ppm.store((int)(total * 1000000ull / nh), std::memory_order_relaxed);
    if (df_staged) {
      memcpy(doc_freqs, b->h_df.p, (size_t)n * WSR_MAX_TERMS * 4);
      memcpy(n_doc_freqs, b->h_ndf.p, (size_t)n * 4);
    }
    *n_queries = (int)n;
    return WSR_OK;
  }
  rc = EnqueueRun(b);
  if (rc) return rc;
  bool df_staged = false;
  if (doc_freqs && n_doc_freqs) {
    rc = EnqueueDocFreqs(b, doc_freqs, n_doc_freqs, &df_staged);
    if (rc) return rc;
  }
  int32_t *cnt = n_hits;
  if (!pinned) {
    CU(b->h_n.Ensure((size_t)n + 1));
    cnt = b->h_n.p;
  }
  // Two ways home for the results. Dense results: the [n, k] array is copied as it is, enqueued
  // right behind the kernels. Sparse results (most queries of the log return far fewer than k
  // hits): counts first, then the hits packed on the GPU, copied and scattered by host threads —
  // less PCIe traffic, but a host round trip between kernel and copy, and the first copy after
  // such a gap was measured ~350 us slower on the B200 boxes (4.7 MB: 444 us vs 93 us back to
  // back). So the packed path is taken only when the index's recent logs were under 10 % full.
  // (the packed path sums hit counts in int32: n*k must fit)
  const bool packed_path = idx->result_fill_ppm.load(std::memory_order_relaxed) < 100000 && nh < 0x7fffffffull;
  size_t total = 0;
  if (packed_path) {
    const size_t cub_bytes = PackTempBytes(n);
    CU(b->d_off.Ensure((size_t)n + 2));
    CU(b->d_pack_cub.Ensure(cub_bytes));
    CU(b->h_totals.Ensure(kPlanWords + 1));
    LaunchResultOffsets(b->out_n, b->d_off.p, n, b->d_pack_cub.p, cub_bytes, b->stream);
    CU(cudaMemcpyAsync(cnt, b->out_n, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(b->h_totals.p, b->d_off.p + n, 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    total = (size_t)(int32_t)b->h_totals.p[0];
    CU(b->d_packed.Ensure(total + 1));
    CU(b->h_packed.Ensure(total + 1));
    LaunchPackResults(b->out_hits, b->out_n, b->d_off.p, n, (uint32_t)k, b->d_packed.p, b->stream);
    CU(cudaGetLastError());
    if (total) CU(cudaMemcpyAsync(b->h_packed.p, b->d_packed.p, total * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    const int T = HostThreads(n, 4096);
    std::vector<size_t> start(T + 1, 0);
    ParallelFor(T, [&](int t, int TT) {
      size_t c = 0;
      for (size_t i = (size_t)n * t / TT, e = (size_t)n * (t + 1) / TT; i < e; i++) c += (size_t)cnt[i];
      start[t + 1] = c;
    });
    for (int t = 0; t < T; t++) start[t + 1] += start[t];
    const wsr_hit *src = b->h_packed.p;
    ParallelFor(T, [&](int t, int TT) {
      size_t at = start[t];
      for (size_t i = (size_t)n * t / TT, e = (size_t)n * (t + 1) / TT; i < e; i++) {
        const size_t c = (size_t)cnt[i];
        if (c) memcpy(hits + i * (size_t)k, src + at, c * sizeof(wsr_hit));
        at += c;
      }
    });
  } else {
    if (pinned) {
      CU(cudaMemcpyAsync(hits, b->out_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
      CU(cudaMemcpyAsync(cnt, b->out_n, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
      CU(cudaStreamSynchronize(b->stream));
    } else {
      CU(b->h_out.Ensure(b->out_hits_bytes + (size_t)n * 4 + 16));
      CU(cudaMemcpyAsync(b->h_out.p, b->d_out.p, b->out_hits_bytes + (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
      CU(cudaStreamSynchronize(b->stream));
      memcpy(hits, b->h_out.p, nh * sizeof(wsr_hit));
      memcpy(cnt, b->h_out.p + b->out_hits_bytes, (size_t)n * 4);
    }
    for (uint32_t i = 0; i < n; i++) total += (size_t)cnt[i];
  }
  if (nh) idx->result_fill_ppm.store((int)(total * 1000000ull / nh), std::memory_order_relaxed);
  if (!pinned) memcpy(n_hits, cnt, (size_t)n * 4);
  if (df_staged) {   // both result paths ended with a stream synchronize: the staging is complete
    memcpy(doc_freqs, b->h_df.p, (size_t)n * WSR_MAX_TERMS * 4);
    memcpy(n_doc_freqs, b->h_ndf.p, (size_t)n * 4);
  }
  *n_queries = (int)n;
  return WSR_OK;
}

}  // namespace

int wsr_search_log(wsr_index *idx, const char *text, size_t len, int k, wsr_hit *hits,
                   int32_t *n_hits, int cap_q, int *n_queries) {
  return wsr_search_log_ex(idx, text, len, k, hits, n_hits, nullptr, nullptr, cap_q, n_queries);
}

int wsr_search_log_ex(wsr_index *idx, const char *text, size_t len, int k, wsr_hit *hits,
                      int32_t *n_hits, uint32_t *doc_freqs, int32_t *n_doc_freqs, int cap_q,
                      int *n_queries) {
  if (!idx || (!text && len) || k < 1 || !hits || !n_hits || !n_queries || cap_q < 0 ||
      (doc_freqs == nullptr) != (n_doc_freqs == nullptr))
    return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(idx->device));
  // k <= kMaxFastK (no collect class) and a dictionary in HBM: parse and plan on the GPU.
  // WSR_HOST_FRONTEND=1 forces the host parser/planner (the same one wsr_search_batch uses).
  if (DeviceFrontEndUsable(idx, len, k))
    return SearchLogOnDevice(idx, text, len, k, hits, n_hits, doc_freqs, n_doc_freqs, cap_q, n_queries);
  // chunk boundaries on line starts: up to 4 chunks of at least 128 KiB of text each
  const int n_chunks = (int)std::max<size_t>(1, std::min<size_t>(4, len >> 17));
  std::vector<size_t> cut(n_chunks + 1, len);
  cut[0] = 0;
  for (int c = 1; c < n_chunks; c++) {
    size_t p = len * c / n_chunks;
    while (p < len && text[p] != '\n') p++;
    cut[c] = p < len ? p + 1 : len;
  }
  const bool pinned = IsPinned(hits) && IsPinned(n_hits);
  PooledBatch pool0(idx), pool1(idx);   // released (after their streams drain) on every return path
  wsr_batch *bt[2] = {pool0.b, pool1.b};
  if (!bt[0] || !bt[1]) return Fail(WSR_ERR_CUDA, "cannot create batch");
  struct Pend { int q0 = 0, n = 0; bool live = false; } pend[2];
  std::vector<wsr_query> qs;
  int done_q = 0, rc = WSR_OK;
  auto drain = [&](int s) -> int {      // waits for slot s and, if staged, copies its results out
    if (!pend[s].live) return WSR_OK;
    CU(cudaStreamSynchronize(bt[s]->stream));
    if (!pinned) {
      memcpy(hits + (size_t)pend[s].q0 * k, bt[s]->h_out.p, (size_t)pend[s].n * k * sizeof(wsr_hit));
      memcpy(n_hits + pend[s].q0, bt[s]->h_out.p + bt[s]->out_hits_bytes, (size_t)pend[s].n * 4);
    }
    pend[s].live = false;
    return WSR_OK;
  };
  for (int c = 0; c < n_chunks && rc == WSR_OK; c++) {
    const int s = c & 1;
    const char *ct = text + cut[c];
    const size_t cl = cut[c + 1] - cut[c];
    size_t lines = 0;
    for (const char *p = ct, *e = ct + cl; p < e;) {
      const char *nl = (const char *)memchr(p, '\n', e - p);
      lines++;
      if (!nl) break;
      p = nl + 1;
    }
    if (done_q + (int)lines > cap_q) { rc = Fail(WSR_ERR_ARG, "result buffers too small"); break; }
    qs.resize(lines + 1);
    int n = 0;
    rc = wsr_parse_query_log(idx, ct, cl, k, qs.data(), (int)qs.size(), &n);   // overlaps the GPU
    if (rc) break;
    if (doc_freqs) {   // vacuum_engine.h:206-219, as wsr_search_batch fills them
      for (int i = 0; i < n; i++) {
        const wsr_query &q = qs[i];
        bool ok = q.k > 0 && q.n_terms > 0;
        for (uint32_t t = 0; ok && t < q.n_terms; t++) ok = q.term_ids[t] != WSR_TERM_ABSENT;
        n_doc_freqs[done_q + i] = ok ? (int32_t)q.n_terms : 0;
        for (uint32_t t = 0; t < WSR_MAX_TERMS; t++)
          doc_freqs[(size_t)(done_q + i) * WSR_MAX_TERMS + t] =
              ok && t < q.n_terms ? idx->host.lists[q.term_ids[t]].df_global : 0u;
      }
    }
    rc = drain(s);                         // this slot's previous chunk must be finished
    if (rc) break;
    wsr_batch *b = bt[s];
    rc = PlanBatch(b, qs.data(), n, k);
    if (rc == WSR_OK) rc = UploadBatch(b);
    if (rc == WSR_OK) rc = EnqueueRun(b);
    if (rc) break;
    const size_t nh = (size_t)n * k;
    if (pinned) {
      if (nh) CU(cudaMemcpyAsync(hits + (size_t)done_q * k, b->out_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
      if (n) CU(cudaMemcpyAsync(n_hits + done_q, b->out_n, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
    } else {
      const size_t bytes = b->out_hits_bytes + (size_t)n * 4;
      CU(b->h_out.Ensure(bytes + 16));
      if (n) CU(cudaMemcpyAsync(b->h_out.p, b->d_out.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    }
    pend[s].q0 = done_q;
    pend[s].n = n;
    pend[s].live = true;
    done_q += n;
  }
  for (int s = 0; s < 2; s++) {
    const int r2 = drain(s);
    if (rc == WSR_OK) rc = r2;
  }
  *n_queries = done_q;
  return rc;
}

int wsr_search(wsr_index *idx, const char *const *terms, const size_t *term_lens, int n_terms,
               int k, unsigned flags, wsr_hit *hits, int *n_hits, uint32_t *doc_freqs,
               int *n_doc_freqs) {
  if (!idx || n_terms < 0 || k < 0 || !n_hits) return Fail(WSR_ERR_ARG, "bad argument");
  if (n_terms > WSR_MAX_TERMS) return Fail(WSR_ERR_UNSUPPORTED, "more than WSR_MAX_TERMS terms");
  *n_hits = 0;
  if (n_doc_freqs) *n_doc_freqs = 0;
  if (k == 0 || n_terms == 0) return WSR_OK;               // vacuum_engine.h:206-215
  wsr_query q;
  memset(&q, 0, sizeof(q));
  q.n_terms = (uint32_t)n_terms;
  q.k = (uint32_t)k;
  q.flags = flags;
  for (int t = 0; t < n_terms; t++) {
    uint32_t id, df;
    if (wsr_term_lookup(idx, terms[t], term_lens[t], &id, &df) != 0) return WSR_OK;  // missing term
    q.term_ids[t] = id;
  }
  uint32_t dfs[WSR_MAX_TERMS];
  int32_t ndf = 0, nh = 0;
  int rc = wsr_search_batch(idx, &q, 1, k, hits, &nh, dfs, &ndf);
  if (rc) return rc;
  *n_hits = nh;
  if (n_doc_freqs) *n_doc_freqs = ndf;
  if (doc_freqs) for (int t = 0; t < ndf; t++) doc_freqs[t] = dfs[t];
  return WSR_OK;
}

int wsr_merge_topk_device(const void *d_gathered_hits, const void *d_gathered_n_hits, int n_shards,
                          int n_queries, int k_stride, void *d_out_hits, void *d_out_n_hits,
                          void *stream) {
  if (!d_gathered_hits || !d_gathered_n_hits || !d_out_hits || !d_out_n_hits || n_shards < 1 ||
      n_queries < 0 || k_stride < 1)
    return Fail(WSR_ERR_ARG, "bad argument");
  LaunchMergeShards((const wsr_hit *)d_gathered_hits, (const int32_t *)d_gathered_n_hits, n_shards,
                    n_queries, k_stride, (wsr_hit *)d_out_hits, (int32_t *)d_out_n_hits,
                    (cudaStream_t)stream);
  CU(cudaGetLastError());
  return WSR_OK;
}

}  // extern "C"

#include "wsr_group.inl"
