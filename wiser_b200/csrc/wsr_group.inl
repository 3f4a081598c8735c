// Document-partitioned groups behind the C ABI: one partition directory per
// index, one or more partitions per GPU, one or more GPUs per process, optionally one rank of a
// multi-process job. Included at the end of wsr_capi.cu (it shares that file's batch internals).
//
// Per query batch and device: the search kernels of every local partition, a merge of the local
// partitions' top-k lists, then the cross-device exchange on the same stream — NCCL called
// directly (no Python in the path): one grouped send/recv in which device r receives every
// device's lists of ITS slice of the queries (hits and counts in the same group = one NCCL
// kernel), the slice merge kernel, one grouped all-gather of the merged slices. NCCL is loaded
// with dlopen so that libwsr.so has no link-time dependency on it (single-GPU users and the CPU
// symbol test never touch it); in a process that already holds a libnccl.so.2 that copy is used.
#include <dlfcn.h>

#include <unordered_map>

namespace {

struct NcclId { char internal[128]; };
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string error;
};
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

NcclApi *Nccl() {
  static NcclApi *api = []() {
    NcclApi *a = new NcclApi;
    const char *names[] = {getenv("WSR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !*n) continue;
      a->handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a->handle) break;
    }
    if (!a->handle) {
      a->error = "cannot load libnccl.so.2 (set WSR_NCCL_LIB to its path)";
      return a;
    }
    auto sym = [&](const char *s) -> void * {
      void *p = dlsym(a->handle, s);
      if (!p && a->error.empty()) a->error = std::string("libnccl lacks ") + s;
      return p;
    };
    a->GetUniqueId = (int (*)(NcclId *))sym("ncclGetUniqueId");
    a->CommInitRank = (int (*)(void **, int, NcclId, int))sym("ncclCommInitRank");
    a->CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
    a->Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))sym("ncclSend");
    a->Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))sym("ncclRecv");
    a->AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))sym("ncclAllGather");
    a->GroupStart = (int (*)())sym("ncclGroupStart");
    a->GroupEnd = (int (*)())sym("ncclGroupEnd");
    a->GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    return a;
  }();
  return api;
}

#define NC(call)                                                                              \
  do {                                                                                        \
    int r_ = (call);                                                                          \
    if (r_ != 0)                                                                              \
      return Fail(WSR_ERR_CUDA, std::string(#call) + ": " +                                   \
                                    (Nccl()->GetErrorString ? Nccl()->GetErrorString(r_) : "NCCL error")); \
  } while (0)

}  // namespace

// One rank of the shard exchange: an NCCL communicator bound to one device, and the exchange
// buffers (grown on demand, reused across batches).
struct wsr_comm {
  void *comm = nullptr;
  int rank = 0, world = 1, device = 0;
  DevBuf<wsr_hit> g_hits, m_hits, full_hits;   // gathered lists, merged slice, gathered merged slices
  DevBuf<int32_t> g_n, m_n, full_n;
  const wsr_hit *res_hits = nullptr;           // result of the last exchange (device pointers)
  const int32_t *res_n = nullptr;
  // doc_freqs travel with the lists when the caller wants them (same groups, same slices)
  DevBuf<uint32_t> g_df, m_df, full_df;
  DevBuf<int32_t> g_ndf, m_ndf, full_ndf;
  const uint32_t *res_df = nullptr;
  const int32_t *res_ndf = nullptr;
};

namespace {

// Slices of the scatter exchange: rank r merges queries [lo(r), lo(r + 1)); all slices have
// ceil(n / world) queries except the last non-empty one (later ranks may be empty).
inline uint32_t SliceLen(uint32_t n, int world) { return n ? (n + (uint32_t)world - 1) / (uint32_t)world : 0; }
inline uint32_t SliceLo(uint32_t n, int world, int r) { return std::min<uint32_t>(n, (uint32_t)r * SliceLen(n, world)); }

// Enqueues the exchange of (loc_hits[n*k], loc_n[n]) on `stream`; afterwards c->res_* point at the
// merged result (n queries) on every rank. mode 0: scatter exchange, 1: plain all-gather.
int ExchangeOnStream(wsr_comm *c, const wsr_hit *loc_hits, const int32_t *loc_n, uint32_t n, uint32_t k,
                     cudaStream_t stream, int mode, const uint32_t *loc_df = nullptr,
                     const int32_t *loc_ndf = nullptr) {
  c->res_df = loc_df;
  c->res_ndf = loc_ndf;
  if (c->world == 1) {
    c->res_hits = loc_hits;
    c->res_n = loc_n;
    return WSR_OK;
  }
  if (loc_df && mode == 1) return Fail(WSR_ERR_UNSUPPORTED, "doc_freqs travel with the scatter exchange only");
  NcclApi *nc = Nccl();
  if (!nc->error.empty()) return Fail(WSR_ERR_UNSUPPORTED, nc->error);
  const int W = c->world;
  const uint32_t s = SliceLen(n, W);
  if (mode == 1) {
    CU(c->g_hits.Ensure((size_t)W * n * k + 1));
    CU(c->g_n.Ensure((size_t)W * n + 1));
    CU(c->full_hits.Ensure((size_t)n * k + 1));
    CU(c->full_n.Ensure((size_t)n + 1));
    NC(nc->GroupStart());
    NC(nc->AllGather(loc_hits, c->g_hits.p, (size_t)n * k * sizeof(wsr_hit), kNcclUint8, c->comm, stream));
    NC(nc->AllGather(loc_n, c->g_n.p, (size_t)n * 4, kNcclUint8, c->comm, stream));
    NC(nc->GroupEnd());
    LaunchMergeShards(c->g_hits.p, c->g_n.p, W, (int)n, (int)k, c->full_hits.p, c->full_n.p, stream);
    CU(cudaGetLastError());
    c->res_hits = c->full_hits.p;
    c->res_n = c->full_n.p;
    return WSR_OK;
  }
  const uint32_t my_lo = SliceLo(n, W, c->rank), mine = SliceLo(n, W, c->rank + 1) - my_lo;
  CU(c->g_hits.Ensure((size_t)W * s * k + 1));
  CU(c->g_n.Ensure((size_t)W * s + 1));
  CU(c->m_hits.Ensure((size_t)s * k + 1));
  CU(c->m_n.Ensure((size_t)s + 1));
  CU(c->full_hits.Ensure((size_t)W * s * k + 1));
  CU(c->full_n.Ensure((size_t)W * s + 1));
  if (loc_df) {
    CU(c->g_df.Ensure((size_t)W * s * WSR_MAX_TERMS + 1));
    CU(c->g_ndf.Ensure((size_t)W * s + 1));
    CU(c->m_df.Ensure((size_t)s * WSR_MAX_TERMS + 1));
    CU(c->m_ndf.Ensure((size_t)s + 1));
    CU(c->full_df.Ensure((size_t)W * s * WSR_MAX_TERMS + 1));
    CU(c->full_ndf.Ensure((size_t)W * s + 1));
  }
  // (1) every rank hands rank j its lists of slice j: hits and counts travel in ONE group
  NC(nc->GroupStart());
  for (int j = 0; j < W; j++) {
    const uint32_t lo = SliceLo(n, W, j), cnt = SliceLo(n, W, j + 1) - lo;
    if (cnt) {
      NC(nc->Send(loc_hits + (size_t)lo * k, (size_t)cnt * k * sizeof(wsr_hit), kNcclUint8, j, c->comm, stream));
      NC(nc->Send(loc_n + lo, (size_t)cnt * 4, kNcclUint8, j, c->comm, stream));
      if (loc_df) {
        NC(nc->Send(loc_df + (size_t)lo * WSR_MAX_TERMS, (size_t)cnt * WSR_MAX_TERMS * 4, kNcclUint8, j, c->comm, stream));
        NC(nc->Send(loc_ndf + lo, (size_t)cnt * 4, kNcclUint8, j, c->comm, stream));
      }
    }
    if (mine) {
      NC(nc->Recv(c->g_hits.p + (size_t)j * mine * k, (size_t)mine * k * sizeof(wsr_hit), kNcclUint8, j, c->comm, stream));
      NC(nc->Recv(c->g_n.p + (size_t)j * mine, (size_t)mine * 4, kNcclUint8, j, c->comm, stream));
      if (loc_df) {
        NC(nc->Recv(c->g_df.p + (size_t)j * mine * WSR_MAX_TERMS, (size_t)mine * WSR_MAX_TERMS * 4, kNcclUint8, j, c->comm, stream));
        NC(nc->Recv(c->g_ndf.p + (size_t)j * mine, (size_t)mine * 4, kNcclUint8, j, c->comm, stream));
      }
    }
  }
  NC(nc->GroupEnd());
  // (2) merge this rank's slice, (3) all-gather the merged slices (padded to s queries; slice r
  // starts at query r*s, so the first n queries of full_* are the result)
  LaunchMergeShards(c->g_hits.p, c->g_n.p, W, (int)mine, (int)k, c->m_hits.p, c->m_n.p, stream);
  CU(cudaGetLastError());
  if (loc_df) {
    LaunchMergeDocFreqs(c->g_df.p, c->g_ndf.p, W, (int)mine, c->m_df.p, c->m_ndf.p, stream);
    CU(cudaGetLastError());
  }
  NC(nc->GroupStart());
  NC(nc->AllGather(c->m_hits.p, c->full_hits.p, (size_t)s * k * sizeof(wsr_hit), kNcclUint8, c->comm, stream));
  NC(nc->AllGather(c->m_n.p, c->full_n.p, (size_t)s * 4, kNcclUint8, c->comm, stream));
  if (loc_df) {
    NC(nc->AllGather(c->m_df.p, c->full_df.p, (size_t)s * WSR_MAX_TERMS * 4, kNcclUint8, c->comm, stream));
    NC(nc->AllGather(c->m_ndf.p, c->full_ndf.p, (size_t)s * 4, kNcclUint8, c->comm, stream));
  }
  NC(nc->GroupEnd());
  CU(cudaGetLastError());
  c->res_hits = c->full_hits.p;
  c->res_n = c->full_n.p;
  if (loc_df) {
    c->res_df = c->full_df.p;
    c->res_ndf = c->full_ndf.p;
  }
  return WSR_OK;
}

}  // namespace

extern "C" {

int wsr_comm_unique_id(char id[WSR_COMM_ID_BYTES]) {
  if (!id) return Fail(WSR_ERR_ARG, "null argument");
  NcclApi *nc = Nccl();
  if (!nc->error.empty()) return Fail(WSR_ERR_UNSUPPORTED, nc->error);
  NcclId u;
  NC(nc->GetUniqueId(&u));
  memcpy(id, u.internal, WSR_COMM_ID_BYTES);
  return WSR_OK;
}

wsr_comm *wsr_comm_init_rank(const char id[WSR_COMM_ID_BYTES], int rank, int world, int device) {
  if (world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) {
    Fail(WSR_ERR_ARG, "bad rank/world");
    return nullptr;
  }
  std::unique_ptr<wsr_comm> c(new wsr_comm);
  c->rank = rank;
  c->world = world;
  c->device = device;
  if (world > 1) {
    NcclApi *nc = Nccl();
    if (!nc->error.empty()) { Fail(WSR_ERR_UNSUPPORTED, nc->error); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { Fail(WSR_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    NcclId u;
    memcpy(u.internal, id, WSR_COMM_ID_BYTES);
    const int r = nc->CommInitRank(&c->comm, world, u, rank);
    if (r != 0) {
      Fail(WSR_ERR_CUDA, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
      return nullptr;
    }
  }
  return c.release();
}

void wsr_comm_destroy(wsr_comm *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm) Nccl()->CommDestroy(c->comm);
  delete c;
}

int wsr_batch_exchange(wsr_batch *b, wsr_comm *c, int mode) {
  if (!b || !c || mode < 0 || mode > 1) return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(b->idx->device));
  return ExchangeOnStream(c, b->out_hits, b->out_n, (uint32_t)b->n, (uint32_t)b->k_stride, b->stream, mode);
}

int wsr_batch_exchanged_results(wsr_batch *b, wsr_comm *c, void **d_hits, void **d_n_hits) {
  if (!b || !c || !c->res_hits) return Fail(WSR_ERR_ARG, "no exchange has run");
  if (d_hits) *d_hits = (void *)c->res_hits;
  if (d_n_hits) *d_n_hits = (void *)c->res_n;
  return WSR_OK;
}

int wsr_batch_fetch_exchanged(wsr_batch *b, wsr_comm *c, wsr_hit *hits, int32_t *n_hits) {
  if (!b || !c || !c->res_hits || !hits || !n_hits) return Fail(WSR_ERR_ARG, "bad argument");
  CU(cudaSetDevice(b->idx->device));
  const size_t nh = (size_t)b->n * b->k_stride;
  if (IsPinned(hits) && IsPinned(n_hits)) {
    if (nh) CU(cudaMemcpyAsync(hits, c->res_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
    if (b->n) CU(cudaMemcpyAsync(n_hits, c->res_n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return WSR_OK;
  }
  CU(b->h_out.Ensure(nh * sizeof(wsr_hit) + (size_t)b->n * 4 + 16));
  if (nh) CU(cudaMemcpyAsync(b->h_out.p, c->res_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, b->stream));
  if (b->n) CU(cudaMemcpyAsync(b->h_out.p + nh * sizeof(wsr_hit), c->res_n, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  if (nh) memcpy(hits, b->h_out.p, nh * sizeof(wsr_hit));
  if (b->n) memcpy(n_hits, b->h_out.p + nh * sizeof(wsr_hit), (size_t)b->n * 4);
  return WSR_OK;
}

}  // extern "C"

// ---- groups ------------------------------------------------------------------------------------
namespace {

struct GroupDevice {
  int device = 0;
  std::vector<int> parts;                 // indices into wsr_group::parts
  std::vector<wsr_batch *> batches;       // one per local partition; batches[0]'s stream leads
  std::vector<cudaEvent_t> done;          // search kernels of local partition p finished
  wsr_comm *comm = nullptr;               // rank of this device in the exchange
  DevBuf<wsr_hit> lg_hits, loc_hits;      // local partitions' lists gathered / merged
  DevBuf<int32_t> lg_n, loc_n;
  DevBuf<uint32_t> lg_df, loc_df;         // the same for doc_freqs
  DevBuf<int32_t> lg_ndf, loc_ndf;
  const wsr_hit *res_hits = nullptr;
  const int32_t *res_n = nullptr;
  const uint32_t *res_df = nullptr;
  const int32_t *res_ndf = nullptr;
  // pipelined runs (wsr_group_run): the exchange of pass i runs on its own stream, from a staged
  // copy of the local lists, while the search kernels of pass i+1 run on the leading stream
  cudaStream_t xstream = nullptr;
  cudaEvent_t ev_staged = nullptr, ev_xdone = nullptr;
  DevBuf<wsr_hit> stage_hits;
  DevBuf<int32_t> stage_n;
  bool x_pending = false;                 // an exchange on xstream has not been joined yet
  int rc = 0;
  std::string err;
};

}  // namespace

struct wsr_group {
  std::vector<wsr_index *> parts;
  std::vector<GroupDevice> devs;
  int rank0 = 0, world = 1;               // this process's first exchange rank / total ranks
  int n = 0, k = 0;                       // loaded log
  bool loaded = false;
  // persistent workers, one per device: the caller posts a job and waits for all of them
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  uint64_t job_id = 0;
  int pending = 0;
  bool stop = false;
  std::function<void(GroupDevice &, int)> job;

  void RunOnDevices(std::function<void(GroupDevice &, int)> fn) {
    if (devs.size() == 1) {   // no hand-off for the single-device case
      fn(devs[0], 0);
      return;
    }
    std::unique_lock<std::mutex> lk(mu);
    job = std::move(fn);
    pending = (int)devs.size();
    job_id++;
    cv_job.notify_all();
    cv_done.wait(lk, [&]() { return pending == 0; });
  }
  void WorkerLoop(int d) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(GroupDevice &, int)> fn;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&]() { return stop || job_id != seen; });
        if (stop) return;
        seen = job_id;
        fn = job;
      }
      fn(devs[d], d);
      std::lock_guard<std::mutex> g(mu);
      if (--pending == 0) cv_done.notify_all();
    }
  }
};

namespace {

int GroupFirstError(wsr_group *g) {
  for (GroupDevice &d : g->devs)
    if (d.rc) return Fail(d.rc, d.err);
  return WSR_OK;
}

// Plans the log on every local partition's batch (each partition has its own dictionary, so each
// parses the text itself, on its GPU).
void GroupLoadOn(wsr_group *g, GroupDevice &d, const char *text, size_t len, int k, int cap_q) {
  d.rc = 0;
  if (cudaSetDevice(d.device) != cudaSuccess) { d.rc = WSR_ERR_CUDA; d.err = "cudaSetDevice failed"; return; }
  for (size_t p = 0; p < d.batches.size(); p++) {
    wsr_batch *b = d.batches[p];
    cudaStreamSynchronize(b->stream);
    int rc;
    if (DeviceFrontEndUsable(b->idx, len, k)) {
      rc = PlanLogOnDevice(b, text, len, k, cap_q);
    } else {
      int nq = 0;
      rc = wsr_batch_reset_log(b, text, len, k, &nq);
    }
    if (rc) { d.rc = rc; d.err = g_err; InvalidateBatch(b); return; }
  }
}

// Enqueues one pass on a device: search kernels of every local partition, local merge, exchange.
void GroupRunOn(wsr_group *g, GroupDevice &d, int mode, bool with_df, bool pipelined) {
  d.rc = 0;
  auto fail = [&](int rc) { d.rc = rc; d.err = g_err; };
  if (cudaSetDevice(d.device) != cudaSuccess) { g_err = "cudaSetDevice failed"; return fail(WSR_ERR_CUDA); }
  const size_t P = d.batches.size();
  wsr_batch *lead = d.batches[0];
  const uint32_t n = (uint32_t)lead->n, k = (uint32_t)lead->k_stride;
  auto cu = [&](cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    fail(WSR_ERR_CUDA);
    return false;
  };
  for (size_t p = 0; p < P; p++) {
    wsr_batch *b = d.batches[p];
    int rc = EnqueueRun(b);
    if (rc) return fail(rc);
    if (with_df) {   // needs the device planner's unplaced queries (k <= 32)
      if (!b->d_tmp.p || !DeviceFrontEndUsable(b->idx, 1, (int)k)) {
        g_err = "doc_freqs of a group need the device front end (k <= 32)";
        return fail(WSR_ERR_UNSUPPORTED);
      }
      if (!cu(b->d_df.Ensure((size_t)n * WSR_MAX_TERMS + 1), "cudaMalloc") || !cu(b->d_ndf.Ensure((size_t)n + 1), "cudaMalloc")) return;
      LaunchDocFreqs(b->d_tmp.p, n, b->idx->view, b->d_df.p, b->d_ndf.p, b->stream);
    }
  }
  const wsr_hit *loc_hits = lead->out_hits;
  const int32_t *loc_n = lead->out_n;
  const uint32_t *loc_df = with_df ? lead->d_df.p : nullptr;
  const int32_t *loc_ndf = with_df ? lead->d_ndf.p : nullptr;
  if (P > 1) {
    if (!cu(d.lg_hits.Ensure(P * (size_t)n * k + 1), "cudaMalloc") || !cu(d.lg_n.Ensure(P * (size_t)n + 1), "cudaMalloc") ||
        !cu(d.loc_hits.Ensure((size_t)n * k + 1), "cudaMalloc") || !cu(d.loc_n.Ensure((size_t)n + 1), "cudaMalloc"))
      return;
    for (size_t p = 0; p < P; p++) {
      wsr_batch *b = d.batches[p];
      if (p) {
        if (!cu(cudaEventRecord(d.done[p], b->stream), "cudaEventRecord") ||
            !cu(cudaStreamWaitEvent(lead->stream, d.done[p], 0), "cudaStreamWaitEvent"))
          return;
      }
      if (!cu(cudaMemcpyAsync(d.lg_hits.p + p * (size_t)n * k, b->out_hits, (size_t)n * k * sizeof(wsr_hit),
                              cudaMemcpyDeviceToDevice, lead->stream), "D2D hits") ||
          !cu(cudaMemcpyAsync(d.lg_n.p + p * (size_t)n, b->out_n, (size_t)n * 4, cudaMemcpyDeviceToDevice,
                              lead->stream), "D2D counts"))
        return;
    }
    LaunchMergeShards(d.lg_hits.p, d.lg_n.p, (int)P, (int)n, (int)k, d.loc_hits.p, d.loc_n.p, lead->stream);
    if (!cu(cudaGetLastError(), "merge kernel")) return;
    if (with_df) {
      if (!cu(d.lg_df.Ensure(P * (size_t)n * WSR_MAX_TERMS + 1), "cudaMalloc") || !cu(d.lg_ndf.Ensure(P * (size_t)n + 1), "cudaMalloc") ||
          !cu(d.loc_df.Ensure((size_t)n * WSR_MAX_TERMS + 1), "cudaMalloc") || !cu(d.loc_ndf.Ensure((size_t)n + 1), "cudaMalloc"))
        return;
      for (size_t p = 0; p < P; p++) {   // (the waits on the partitions' streams were enqueued above)
        wsr_batch *b = d.batches[p];
        if (!cu(cudaMemcpyAsync(d.lg_df.p + p * (size_t)n * WSR_MAX_TERMS, b->d_df.p, (size_t)n * WSR_MAX_TERMS * 4,
                                cudaMemcpyDeviceToDevice, lead->stream), "D2D doc_freqs") ||
            !cu(cudaMemcpyAsync(d.lg_ndf.p + p * (size_t)n, b->d_ndf.p, (size_t)n * 4, cudaMemcpyDeviceToDevice,
                                lead->stream), "D2D doc_freq counts"))
          return;
      }
      LaunchMergeDocFreqs(d.lg_df.p, d.lg_ndf.p, (int)P, (int)n, d.loc_df.p, d.loc_ndf.p, lead->stream);
      if (!cu(cudaGetLastError(), "doc_freqs merge kernel")) return;
      loc_df = d.loc_df.p;
      loc_ndf = d.loc_ndf.p;
    }
    // the next pass of partition p > 0 must not overwrite its results before they were copied
    if (!cu(cudaEventRecord(d.done[0], lead->stream), "cudaEventRecord")) return;
    for (size_t p = 1; p < P; p++)
      if (!cu(cudaStreamWaitEvent(d.batches[p]->stream, d.done[0], 0), "cudaStreamWaitEvent")) return;
    loc_hits = d.loc_hits.p;
    loc_n = d.loc_n.p;
  }
  int rc;
  if (pipelined && d.comm->world > 1 && !with_df) {
    // stage the local lists (a device-to-device copy of n*k*16 B), hand the exchange to its own
    // stream and return: the next pass's search kernels overlap this pass's exchange. The staging
    // copy waits for the previous exchange, which still reads the staging buffers.
    if (!cu(d.stage_hits.Ensure((size_t)n * k + 1), "cudaMalloc") || !cu(d.stage_n.Ensure((size_t)n + 1), "cudaMalloc")) return;
    if (d.x_pending && !cu(cudaStreamWaitEvent(lead->stream, d.ev_xdone, 0), "cudaStreamWaitEvent")) return;
    if (!cu(cudaMemcpyAsync(d.stage_hits.p, loc_hits, (size_t)n * k * sizeof(wsr_hit), cudaMemcpyDeviceToDevice, lead->stream), "D2D stage") ||
        !cu(cudaMemcpyAsync(d.stage_n.p, loc_n, (size_t)n * 4, cudaMemcpyDeviceToDevice, lead->stream), "D2D stage") ||
        !cu(cudaEventRecord(d.ev_staged, lead->stream), "cudaEventRecord") ||
        !cu(cudaStreamWaitEvent(d.xstream, d.ev_staged, 0), "cudaStreamWaitEvent"))
      return;
    rc = ExchangeOnStream(d.comm, d.stage_hits.p, d.stage_n.p, n, k, d.xstream, mode);
    if (rc) return fail(rc);
    if (!cu(cudaEventRecord(d.ev_xdone, d.xstream), "cudaEventRecord")) return;
    d.x_pending = true;
  } else {
    if (d.x_pending && !cu(cudaStreamWaitEvent(lead->stream, d.ev_xdone, 0), "cudaStreamWaitEvent")) return;
    d.x_pending = false;
    rc = ExchangeOnStream(d.comm, loc_hits, loc_n, n, k, lead->stream, mode, loc_df, loc_ndf);
    if (rc) return fail(rc);
  }
  d.res_hits = d.comm->res_hits;
  d.res_n = d.comm->res_n;
  d.res_df = d.comm->res_df;
  d.res_ndf = d.comm->res_ndf;
}

}  // namespace

extern "C" {

wsr_group *wsr_group_open(const char *const *dirs, int n_dirs, const int *devices, int n_dev, int loader_threads,
                          unsigned flags, const wsr_group_dist *dist, char *err, size_t errlen) {
  auto fail = [&](const std::string &m) -> wsr_group * {
    g_err = m;
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!dirs || n_dirs < 1 || (n_dev > 0 && !devices)) return fail("bad argument");
  if (n_dev < 1) n_dev = 1;
  if (n_dirs % n_dev) return fail("the number of partitions must be a multiple of the number of devices");
  if (dist && (dist->world < 1 || dist->rank < 0 || dist->rank >= dist->world)) return fail("bad dist rank/world");
  if (dist && dist->world > 1 && n_dev != 1) return fail("a multi-process group drives one device per process");
  std::unique_ptr<wsr_group> g(new wsr_group);
  const int per_dev = n_dirs / n_dev;
  g->devs.resize(n_dev);
  auto cleanup = [&]() {
    for (GroupDevice &d : g->devs) {
      for (wsr_batch *b : d.batches) FreeBatch(b);
      for (cudaEvent_t e : d.done) if (e) cudaEventDestroy(e);
      if (d.xstream) cudaStreamDestroy(d.xstream);
      if (d.ev_staged) cudaEventDestroy(d.ev_staged);
      if (d.ev_xdone) cudaEventDestroy(d.ev_xdone);
      if (d.comm) wsr_comm_destroy(d.comm);
    }
    for (wsr_index *ix : g->parts) wsr_index_close(ix);
  };
  for (int p = 0; p < n_dirs; p++) {
    const int dv = devices ? devices[p / per_dev] : 0;
    char e[512] = {0};
    wsr_index *ix = wsr_index_open_ex(dirs[p], dv, 0, 1, loader_threads, flags, e, sizeof(e));
    if (!ix) { cleanup(); return fail(e); }
    g->parts.push_back(ix);
    GroupDevice &d = g->devs[p / per_dev];
    d.device = dv;
    d.parts.push_back(p);
  }
  // collection statistics over the LOCAL partitions (a multi-process job sets them afterwards
  // through wsr_index_set_global_stats on wsr_group_part(): it alone knows the other ranks')
  if (!(dist && dist->world > 1) && n_dirs > 1) {
    int64_t total = 0;
    double acc = 0.0;
    std::vector<int64_t> base(n_dirs);
    for (int p = 0; p < n_dirs; p++) {
      base[p] = total;
      total += g->parts[p]->host.n_docs;
      acc += g->parts[p]->host.avg_len * (double)g->parts[p]->host.n_docs;
    }
    const double avg = acc / (double)total;   // same arithmetic as wiser_b200/dist.py combine_partition_stats
    // global df: dense table over term ranks when every term is named t<rank> (the synthetic
    // corpora), else a string-keyed table
    std::vector<std::vector<uint32_t>> dfg(n_dirs);
    bool dense = true;
    std::vector<std::vector<uint32_t>> ranks(n_dirs);
    uint32_t max_rank = 0;
    for (int p = 0; p < n_dirs && dense; p++) {
      const size_t nt = g->parts[p]->host.lists.size();
      ranks[p].resize(nt);
      if (wsr_index_local_stats(g->parts[p], nullptr, ranks[p].data()) != WSR_OK) dense = false;
      else for (uint32_t r : ranks[p]) max_rank = std::max(max_rank, r);
    }
    if (dense) {
      std::vector<uint32_t> tab((size_t)max_rank + 1, 0u);
      for (int p = 0; p < n_dirs; p++)
        for (size_t t = 0; t < ranks[p].size(); t++) tab[ranks[p][t]] += g->parts[p]->host.lists[t].df_shard;
      for (int p = 0; p < n_dirs; p++) {
        dfg[p].resize(ranks[p].size());
        for (size_t t = 0; t < ranks[p].size(); t++) dfg[p][t] = tab[ranks[p][t]];
      }
    } else {
      std::unordered_map<std::string, uint32_t> tab;
      for (int p = 0; p < n_dirs; p++) {
        const HostIndex &h = g->parts[p]->host;
        for (size_t t = 0; t < h.lists.size(); t++)
          tab[std::string(h.term_arena.data() + h.term_off[t], h.term_off[t + 1] - h.term_off[t])] += h.lists[t].df_shard;
      }
      for (int p = 0; p < n_dirs; p++) {
        const HostIndex &h = g->parts[p]->host;
        dfg[p].resize(h.lists.size());
        for (size_t t = 0; t < h.lists.size(); t++)
          dfg[p][t] = tab[std::string(h.term_arena.data() + h.term_off[t], h.term_off[t + 1] - h.term_off[t])];
      }
    }
    for (int p = 0; p < n_dirs; p++)
      if (wsr_index_set_global_stats(g->parts[p], base[p], total, avg, dfg[p].data()) != WSR_OK) {
        const std::string m = g_err;
        cleanup();
        return fail("wsr_index_set_global_stats: " + m);
      }
  }
  // batches, events
  for (GroupDevice &d : g->devs) {
    for (int p : d.parts) {
      wsr_batch *b = NewBatch(g->parts[p]);
      if (!b) { cleanup(); return fail("cannot create batch"); }
      d.batches.push_back(b);
      cudaEvent_t e = nullptr;
      cudaSetDevice(d.device);
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cleanup(); return fail("cudaEventCreate failed"); }
      d.done.push_back(e);
    }
    cudaSetDevice(d.device);
    if (cudaStreamCreateWithFlags(&d.xstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.ev_staged, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.ev_xdone, cudaEventDisableTiming) != cudaSuccess) {
      cleanup();
      return fail("cannot create the exchange stream");
    }
  }
  // exchange ranks: the devices of this process are consecutive ranks starting at rank0
  g->world = dist && dist->world > 1 ? dist->world : n_dev;
  g->rank0 = dist && dist->world > 1 ? dist->rank : 0;
  char id[WSR_COMM_ID_BYTES] = {0};
  if (dist && dist->world > 1) memcpy(id, dist->comm_id, WSR_COMM_ID_BYTES);
  else if (n_dev > 1 && wsr_comm_unique_id(id) != WSR_OK) { const std::string m = g_err; cleanup(); return fail(m); }
  if (n_dev > 1)
    for (int d = 0; d < n_dev; d++) g->workers.emplace_back([gp = g.get(), d]() { gp->WorkerLoop(d); });
  // ncclCommInitRank blocks until every rank has called it: one thread per device
  g->RunOnDevices([&](GroupDevice &d, int di) {
    d.comm = wsr_comm_init_rank(id, g->rank0 + di, g->world, d.device);
    if (!d.comm) { d.rc = WSR_ERR_CUDA; d.err = g_err; }
  });
  for (GroupDevice &d : g->devs)
    if (!d.comm) {
      const std::string m = d.err;
      wsr_group *raw = g.release();
      wsr_group_close(raw);
      return fail("communicator: " + m);
    }
  return g.release();
}

void wsr_group_close(wsr_group *g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true;
  }
  g->cv_job.notify_all();
  for (std::thread &t : g->workers) t.join();
  for (GroupDevice &d : g->devs) {
    cudaSetDevice(d.device);
    if (d.xstream) cudaStreamSynchronize(d.xstream);
    for (wsr_batch *b : d.batches) FreeBatch(b);
    for (cudaEvent_t e : d.done) if (e) cudaEventDestroy(e);
    if (d.xstream) cudaStreamDestroy(d.xstream);
    if (d.ev_staged) cudaEventDestroy(d.ev_staged);
    if (d.ev_xdone) cudaEventDestroy(d.ev_xdone);
    if (d.comm) wsr_comm_destroy(d.comm);
  }
  for (wsr_index *ix : g->parts) wsr_index_close(ix);
  delete g;
}

int wsr_group_n_parts(const wsr_group *g) { return g ? (int)g->parts.size() : 0; }
wsr_index *wsr_group_part(wsr_group *g, int i) {
  if (!g || i < 0 || (size_t)i >= g->parts.size()) { Fail(WSR_ERR_ARG, "bad partition index"); return nullptr; }
  return g->parts[i];
}

int wsr_group_load_log(wsr_group *g, const char *text, size_t len, int k, int *n_queries) {
  if (!g || (!text && len) || k < 1 || !n_queries) return Fail(WSR_ERR_ARG, "bad argument");
  g->loaded = false;
  g->RunOnDevices([&](GroupDevice &d, int) { GroupLoadOn(g, d, text, len, k, 0x7fffffff); });
  const int rc = GroupFirstError(g);
  if (rc) return rc;
  g->n = g->devs[0].batches[0]->n;
  g->k = k;
  for (GroupDevice &d : g->devs)
    for (wsr_batch *b : d.batches)
      if (b->n != g->n) return Fail(WSR_ERR_ARG, "partitions disagree on the number of log lines");
  g->loaded = true;
  *n_queries = g->n;
  return WSR_OK;
}

int wsr_group_run(wsr_group *g, int mode) {
  if (!g || !g->loaded || mode < 0 || mode > 1) return Fail(WSR_ERR_ARG, "no log loaded");
  g->RunOnDevices([&](GroupDevice &d, int) { GroupRunOn(g, d, mode, false, true); });
  return GroupFirstError(g);
}

int wsr_group_join(wsr_group *g) {
  if (!g) return Fail(WSR_ERR_ARG, "null group");
  for (GroupDevice &d : g->devs) {
    if (!d.x_pending) continue;
    CU(cudaSetDevice(d.device));
    CU(cudaStreamWaitEvent(d.batches[0]->stream, d.ev_xdone, 0));
    d.x_pending = false;
  }
  return WSR_OK;
}

int wsr_group_sync(wsr_group *g) {
  if (!g) return Fail(WSR_ERR_ARG, "null group");
  for (GroupDevice &d : g->devs) {
    CU(cudaSetDevice(d.device));
    for (wsr_batch *b : d.batches) CU(cudaStreamSynchronize(b->stream));
    CU(cudaStreamSynchronize(d.xstream));
  }
  return WSR_OK;
}

int wsr_group_stream(wsr_group *g, void **stream) {
  if (!g || !stream) return Fail(WSR_ERR_ARG, "null argument");
  *stream = (void *)g->devs[0].batches[0]->stream;
  return WSR_OK;
}

int wsr_group_fetch(wsr_group *g, wsr_hit *hits, int32_t *n_hits) {
  if (!g || !g->loaded || !hits || !n_hits) return Fail(WSR_ERR_ARG, "bad argument");
  GroupDevice &d = g->devs[0];
  if (!d.res_hits) return Fail(WSR_ERR_ARG, "wsr_group_run has not run");
  { const int jrc = wsr_group_join(g); if (jrc) return jrc; }
  CU(cudaSetDevice(d.device));
  wsr_batch *lead = d.batches[0];
  const size_t nh = (size_t)g->n * g->k;
  const bool pinned = IsPinned(hits) && IsPinned(n_hits);
  if (!pinned) CU(lead->h_out.Ensure(nh * sizeof(wsr_hit) + (size_t)g->n * 4 + 16));
  wsr_hit *hdst = pinned ? hits : reinterpret_cast<wsr_hit *>(lead->h_out.p);
  int32_t *ndst = pinned ? n_hits : reinterpret_cast<int32_t *>(lead->h_out.p + nh * sizeof(wsr_hit));
  if (nh) CU(cudaMemcpyAsync(hdst, d.res_hits, nh * sizeof(wsr_hit), cudaMemcpyDeviceToHost, lead->stream));
  if (g->n) CU(cudaMemcpyAsync(ndst, d.res_n, (size_t)g->n * 4, cudaMemcpyDeviceToHost, lead->stream));
  CU(cudaStreamSynchronize(lead->stream));
  if (!pinned) {
    if (nh) memcpy(hits, hdst, nh * sizeof(wsr_hit));
    if (g->n) memcpy(n_hits, ndst, (size_t)g->n * 4);
  }
  return WSR_OK;
}

int wsr_group_stats(wsr_group *g, uint64_t *listed_postings, uint64_t *n_postings, uint64_t *hbm_bytes, int64_t *n_docs) {
  if (!g) return Fail(WSR_ERR_ARG, "null group");
  uint64_t np = 0, hb = 0;
  for (wsr_index *ix : g->parts) { np += (uint64_t)ix->host.n_postings; hb += (uint64_t)ix->hbm_bytes; }
  if (n_postings) *n_postings = np;
  if (hbm_bytes) *hbm_bytes = hb;
  if (n_docs) *n_docs = g->parts[0]->host.n_docs;
  if (listed_postings) {
    uint64_t l = 0;
    for (GroupDevice &d : g->devs)
      for (wsr_batch *b : d.batches) l += b->listed_postings;
    *listed_postings = l;
  }
  return WSR_OK;
}

int wsr_group_search_log(wsr_group *g, const char *text, size_t len, int k, wsr_hit *hits, int32_t *n_hits,
                         uint32_t *doc_freqs, int32_t *n_doc_freqs, int cap_q, int *n_queries) {
  if (!g || (!text && len) || k < 1 || (hits == nullptr) != (n_hits == nullptr) || !n_queries || cap_q < 0 ||
      (doc_freqs == nullptr) != (n_doc_freqs == nullptr) || (!hits && doc_freqs))
    return Fail(WSR_ERR_ARG, "bad argument");
  int n = 0;
  int rc = wsr_group_load_log(g, text, len, k, &n);
  if (rc) return rc;
  if (hits && n > cap_q) return Fail(WSR_ERR_ARG, "result buffers too small");
  // doc_freqs are computed by every partition from its own dictionary and merged with the lists
  // (a term may be missing from one partition's dictionary and present in another's). Every rank
  // of a multi-process job must make the same choice, so they always travel when k allows it.
  const bool with_df = k <= kMaxFastK;
  if (doc_freqs && !with_df) return Fail(WSR_ERR_UNSUPPORTED, "doc_freqs of a group need the device front end (k <= 32)");
  g->RunOnDevices([&](GroupDevice &d, int) { GroupRunOn(g, d, 0, with_df, false); });
  rc = GroupFirstError(g);
  if (rc) return rc;
  if (!hits) {   // a rank of a multi-process job that does not face the client: search + exchange only
    rc = wsr_group_sync(g);
    if (rc == WSR_OK) *n_queries = n;
    return rc;
  }
  GroupDevice &d0 = g->devs[0];
  wsr_batch *lead = d0.batches[0];
  bool df_staged = false;
  if (doc_freqs) {
    CU(cudaSetDevice(d0.device));
    uint32_t *df = doc_freqs;
    int32_t *ndf = n_doc_freqs;
    if (!IsPinned(doc_freqs) || !IsPinned(n_doc_freqs)) {
      CU(lead->h_df.Ensure((size_t)n * WSR_MAX_TERMS + 1));
      CU(lead->h_ndf.Ensure((size_t)n + 1));
      df = lead->h_df.p;
      ndf = lead->h_ndf.p;
      df_staged = true;
    }
    if (n) {
      CU(cudaMemcpyAsync(df, d0.res_df, (size_t)n * WSR_MAX_TERMS * 4, cudaMemcpyDeviceToHost, lead->stream));
      CU(cudaMemcpyAsync(ndf, d0.res_ndf, (size_t)n * 4, cudaMemcpyDeviceToHost, lead->stream));
    }
  }
  rc = wsr_group_fetch(g, hits, n_hits);
  if (rc) return rc;
  rc = wsr_group_sync(g);
  if (rc) return rc;
  if (df_staged && n) {
    memcpy(doc_freqs, lead->h_df.p, (size_t)n * WSR_MAX_TERMS * 4);
    memcpy(n_doc_freqs, lead->h_ndf.p, (size_t)n * 4);
  }
  *n_queries = n;
  return WSR_OK;
}

}  // extern "C"
