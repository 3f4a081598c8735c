#include "gpu_vacuum_engine.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace wsr {

namespace {
[[noreturn]] void Fatal(const std::string &msg) {
  // The reference reports misuse and corrupt input with LOG(FATAL), which aborts.
  fprintf(stderr, "[F gpu_vacuum_engine] %s\n", msg.c_str());
  abort();
}
}  // namespace

struct GpuVacuumEngine::Pending {
  wsr_query q;
  const SearchQuery *src = nullptr;   // group mode: the query text is sent, not term ids
  std::string err;                    // message of a failed batch (set by the leader's thread)
  std::vector<wsr_hit> hits;
  int32_t n_hits = 0;
  bool done = false;
  int rc = 0;
  // own wake-up: a finished batch wakes exactly its callers. The condition variable belongs to
  // the calling THREAD (shared ownership), so a leader may notify after releasing the lock
  // without racing the caller's return.
  std::shared_ptr<std::condition_variable> cv;
};

GpuVacuumEngine::GpuVacuumEngine(const std::string engine_dir_path, int bloom_enable_factor,
                                 GpuEngineOptions opt)
    : dir_(engine_dir_path), bloom_enable_factor_(bloom_enable_factor), opt_(opt) {}

GpuVacuumEngine::~GpuVacuumEngine() {
  if (group_) wsr_group_close(group_);
  else if (idx_) wsr_index_close(idx_);
}

bool GpuVacuumEngine::LookupAnyPartition(const std::string &term, uint32_t *df) const {
  // a term may be missing from one partition's dictionary and present in another's; the df a
  // partition reports is the collection-wide one (set by the group's statistics exchange)
  const int n = group_ ? wsr_group_n_parts(group_) : 1;
  for (int p = 0; p < n; p++) {
    uint32_t id;
    wsr_index *ix = group_ ? wsr_group_part(group_, p) : idx_;
    if (wsr_term_lookup(ix, term.data(), term.size(), &id, df) == 0) return true;
  }
  return false;
}

void GpuVacuumEngine::Load() {
  if (idx_) Fatal("Engine is already loaded.");                     // vacuum_engine.h:145
  char err[512] = {0};
  if (!opt_.partition_dirs.empty()) {
    std::vector<const char *> dirs;
    for (const std::string &d : opt_.partition_dirs) dirs.push_back(d.c_str());
    std::vector<int> devs = opt_.devices.empty() ? std::vector<int>{opt_.device} : opt_.devices;
    group_ = wsr_group_open(dirs.data(), (int)dirs.size(), devs.data(), (int)devs.size(), opt_.loader_threads,
                            opt_.load_positions ? WSR_OPEN_POSITIONS : 0u, nullptr, err, sizeof(err));
    if (!group_) Fatal(std::string("wsr_group_open: ") + err);
    idx_ = wsr_group_part(group_, 0);
    return;
  }
  idx_ = wsr_index_open_ex(dir_.c_str(), opt_.device, opt_.shard, opt_.n_shards, opt_.loader_threads,
                           opt_.load_positions ? WSR_OPEN_POSITIONS : 0u, err, sizeof(err));
  if (!idx_) Fatal(std::string("wsr_index_open: ") + err);
}

int GpuVacuumEngine::TermCount() const {
  wsr_index_info info;
  if (!idx_ || wsr_index_get_info(idx_, &info) != 0) Fatal("Engine is not yet loaded");
  return (int)info.n_terms;
}

std::map<std::string, int> GpuVacuumEngine::PostinglistSizes(const TermList &terms) {
  std::map<std::string, int> ret;
  for (auto &term : terms) {
    uint32_t df;
    if (LookupAnyPartition(term, &df)) ret[term] = (int)df;
  }
  return ret;
}

// Returns false when the reference would return the empty result early.
bool GpuVacuumEngine::ToWsrQuery(const SearchQuery &q, wsr_query *out) const {
  memset(out, 0, sizeof(*out));
  if (q.n_results <= 0 || q.terms.empty()) return false;           // vacuum_engine.h:206-208
  if (q.terms.size() > WSR_MAX_TERMS) Fatal("more than WSR_MAX_TERMS query terms");
  out->n_terms = (uint32_t)q.terms.size();
  out->k = (uint32_t)q.n_results;
  out->flags = q.is_phrase ? WSR_QUERY_PHRASE : 0u;
  for (size_t t = 0; t < q.terms.size(); t++) {
    uint32_t id, df;
    if (wsr_term_lookup(idx_, q.terms[t].data(), q.terms[t].size(), &id, &df) != 0)
      return false;                                                  // vacuum_engine.h:213-215
    out->term_ids[t] = id;
  }
  return true;
}

SearchResult GpuVacuumEngine::Search(const SearchQuery &query) {
  SearchResult result;
  if (!idx_) Fatal("Engine is not yet loaded");
  Pending p;
  if (group_) {
    // group mode: every partition resolves the terms in its own dictionary on the GPU; here only
    // the early-outs and doc_freqs (vacuum_engine.h:206-219)
    if (query.n_results <= 0 || query.terms.empty()) return result;
    if (query.terms.size() > WSR_MAX_TERMS) Fatal("more than WSR_MAX_TERMS query terms");
    std::vector<int> dfs;
    for (const Term &t : query.terms) {
      uint32_t df;
      if (!LookupAnyPartition(t, &df)) return result;
      dfs.push_back((int)df);
    }
    result.doc_freqs = dfs;
    memset(&p.q, 0, sizeof(p.q));
    p.q.k = (uint32_t)query.n_results;
    p.src = &query;
  } else {
    // one dictionary lookup per term: ids for the query, dfs for doc_freqs (vacuum_engine.h:217-219)
    memset(&p.q, 0, sizeof(p.q));
    if (query.n_results <= 0 || query.terms.empty()) return result;           // vacuum_engine.h:206-208
    if (query.terms.size() > WSR_MAX_TERMS) Fatal("more than WSR_MAX_TERMS query terms");
    p.q.n_terms = (uint32_t)query.terms.size();
    p.q.k = (uint32_t)query.n_results;
    p.q.flags = query.is_phrase ? WSR_QUERY_PHRASE : 0u;
    std::vector<int> dfs;
    for (size_t t = 0; t < query.terms.size(); t++) {
      uint32_t id, df;
      if (wsr_term_lookup(idx_, query.terms[t].data(), query.terms[t].size(), &id, &df) != 0)
        return result;                                                         // vacuum_engine.h:213-215
      p.q.term_ids[t] = id;
      dfs.push_back((int)df);
    }
    result.doc_freqs = dfs;
  }
  p.hits.resize(p.q.k);
  thread_local std::shared_ptr<std::condition_variable> my_cv = std::make_shared<std::condition_variable>();
  p.cv = my_cv;
  {
    std::vector<std::shared_ptr<std::condition_variable>> wake;
    std::unique_lock<std::mutex> lk(mu_);
    pending_.push_back(&p);
    std::vector<Pending *> take;
    while (!p.done) {
      // (a group runs one batch at a time: its per-device batches and exchange buffers are single)
      if (inflight_ < (group_ ? 1 : std::max(1, opt_.max_inflight)) && !pending_.empty()) {
        // lead: everything queued so far (this caller is in it unless another leader took it)
        inflight_++;
        if (opt_.coalesce_window_us > 0 && (int)pending_.size() < opt_.coalesce_max_batch)
          p.cv->wait_for(lk, std::chrono::microseconds(opt_.coalesce_window_us),
                        [&]() { return (int)pending_.size() >= opt_.coalesce_max_batch; });
        take.clear();
        if ((int)pending_.size() <= opt_.coalesce_max_batch) {
          take.swap(pending_);
        } else {
          take.assign(pending_.begin(), pending_.begin() + opt_.coalesce_max_batch);
          pending_.erase(pending_.begin(), pending_.begin() + opt_.coalesce_max_batch);
        }
        lk.unlock();
        if (group_) RunGroupBatch(take);
        else RunBatch(take);
        lk.lock();
        inflight_--;
        wake.clear();
        for (Pending *t : take) {
          t->done = true;
          if (t != &p) wake.push_back(t->cv);
        }
        // hand the free slot to a queued caller (arrivals that find a free slot lead themselves)
        if (!pending_.empty()) wake.push_back(pending_.front()->cv);
        lk.unlock();
        for (auto &c : wake) c->notify_one();     // outside the lock: woken callers do not pile up on it
        lk.lock();
      } else {
        p.cv->wait(lk);
      }
    }
  }
  if (p.rc != 0) Fatal(std::string("search batch failed: ") + p.err);
  for (int i = 0; i < p.n_hits; i++) {
    SearchResultEntry e;
    e.doc_id = p.hits[i].doc_id;
    e.doc_score = p.hits[i].score;
    result.entries.push_back(e);
  }
  return result;
}

// One coalesced batch on the calling (leader) thread: submit, then hand every caller its hits.
void GpuVacuumEngine::RunBatch(const std::vector<Pending *> &take) {
  thread_local std::vector<wsr_query> qs;
  thread_local std::vector<wsr_hit> hits;
  thread_local std::vector<int32_t> n_hits;
  uint32_t k_stride = 1;
  qs.clear();
  for (Pending *p : take) {
    qs.push_back(p->q);
    if (p->q.k > k_stride) k_stride = p->q.k;
  }
  hits.resize(qs.size() * (size_t)k_stride);
  n_hits.assign(qs.size(), 0);
  const int rc = wsr_search_batch(idx_, qs.data(), (int)qs.size(), (int)k_stride, hits.data(),
                                  n_hits.data(), nullptr, nullptr);
  const std::string msg = rc ? wsr_last_error() : "";   // this (the leader's) thread's message
  for (size_t i = 0; i < take.size(); i++) {
    Pending *p = take[i];
    p->rc = rc;
    p->err = msg;
    p->n_hits = rc == 0 ? n_hits[i] : 0;
    for (int j = 0; j < p->n_hits; j++) p->hits[j] = hits[i * (size_t)k_stride + j];
  }
}

// Group mode: the coalesced queries travel as query-log text (one line each; a phrase in double
// quotes) to wsr_group_search_log; a query asking for fewer results than the batch's k takes the
// head of its list.
void GpuVacuumEngine::RunGroupBatch(const std::vector<Pending *> &take) {
  thread_local std::string text;
  thread_local std::vector<wsr_hit> hits;
  thread_local std::vector<int32_t> n_hits;
  uint32_t k = 1;
  text.clear();
  for (Pending *p : take) {
    if (p->q.k > k) k = p->q.k;
    if (p->src->is_phrase) text += '"';
    for (size_t t = 0; t < p->src->terms.size(); t++) {
      if (t) text += ' ';
      text += p->src->terms[t];
    }
    if (p->src->is_phrase) text += '"';
    text += '\n';
  }
  hits.resize(take.size() * (size_t)k);
  n_hits.assign(take.size(), 0);
  int got = 0;
  int rc = wsr_group_search_log(group_, text.data(), text.size(), (int)k, hits.data(), n_hits.data(), nullptr,
                                nullptr, (int)take.size(), &got);
  std::string msg = rc ? wsr_last_error() : "";
  if (rc == 0 && got != (int)take.size()) { rc = WSR_ERR_ARG; msg = "query text produced a different number of lines"; }
  for (size_t i = 0; i < take.size(); i++) {
    Pending *p = take[i];
    p->rc = rc;
    p->err = msg;
    p->n_hits = rc == 0 ? std::min<int32_t>(n_hits[i], (int32_t)p->q.k) : 0;
    for (int j = 0; j < p->n_hits; j++) p->hits[j] = hits[i * (size_t)k + j];
  }
}

std::vector<SearchResult> GpuVacuumEngine::SearchBatch(const std::vector<SearchQuery> &queries) {
  std::vector<SearchResult> out(queries.size());
  if (group_) {
    std::vector<Pending> pend(queries.size());
    std::vector<Pending *> take;
    for (size_t i = 0; i < queries.size(); i++) {
      const SearchQuery &q = queries[i];
      if (q.n_results <= 0 || q.terms.empty()) continue;
      if (q.terms.size() > WSR_MAX_TERMS) Fatal("more than WSR_MAX_TERMS query terms");
      bool ok = true;
      std::vector<int> dfs;
      for (const Term &t : q.terms) {
        uint32_t df;
        if (!LookupAnyPartition(t, &df)) { ok = false; break; }
        dfs.push_back((int)df);
      }
      if (!ok) continue;
      out[i].doc_freqs = dfs;
      memset(&pend[i].q, 0, sizeof(wsr_query));
      pend[i].q.k = (uint32_t)q.n_results;
      pend[i].src = &q;
      pend[i].hits.resize(q.n_results);
      take.push_back(&pend[i]);
    }
    if (!take.empty()) RunGroupBatch(take);
    for (size_t i = 0; i < queries.size(); i++) {
      if (pend[i].rc != 0) Fatal(std::string("search batch failed: ") + pend[i].err);
      for (int j = 0; j < pend[i].n_hits; j++) {
        SearchResultEntry e;
        e.doc_id = pend[i].hits[j].doc_id;
        e.doc_score = pend[i].hits[j].score;
        out[i].entries.push_back(e);
      }
    }
    return out;
  }
  std::vector<wsr_query> qs(queries.size());
  uint32_t k_stride = 1;
  for (size_t i = 0; i < queries.size(); i++) {
    if (ToWsrQuery(queries[i], &qs[i])) {
      for (auto &t : queries[i].terms) {
        uint32_t id, df;
        wsr_term_lookup(idx_, t.data(), t.size(), &id, &df);
        out[i].doc_freqs.push_back((int)df);
      }
      if (qs[i].k > k_stride) k_stride = qs[i].k;
    } else {
      memset(&qs[i], 0, sizeof(wsr_query));
    }
  }
  std::vector<wsr_hit> hits(queries.size() * (size_t)k_stride);
  std::vector<int32_t> n_hits(queries.size(), 0);
  if (wsr_search_batch(idx_, qs.data(), (int)qs.size(), (int)k_stride, hits.data(), n_hits.data(),
                       nullptr, nullptr) != 0)
    Fatal(std::string("wsr_search_batch: ") + wsr_last_error());
  for (size_t i = 0; i < queries.size(); i++)
    for (int j = 0; j < n_hits[i]; j++) {
      SearchResultEntry e;
      e.doc_id = hits[i * (size_t)k_stride + j].doc_id;
      e.doc_score = hits[i * (size_t)k_stride + j].score;
      out[i].entries.push_back(e);
    }
  return out;
}

void GpuVacuumEngine::AddDocument(const DocInfo) { Fatal("Not implemented in VacuumEngine."); }
int GpuVacuumEngine::LoadLocalDocuments(const std::string &, int, const std::string) {
  Fatal("Not implemented in VacuumEngine.");
}
void GpuVacuumEngine::Serialize(std::string) const { Fatal("Not implemented in VacuumEngine."); }
void GpuVacuumEngine::Deserialize(std::string) { Fatal("Not implemented in VacuumEngine."); }

std::unique_ptr<SearchEngineServiceNew> CreateSearchEngine(std::string engine_type,
                                                           int bloom_enable_factor) {
  const std::string prefix = "gpu:vacuum_dump:";
  if (engine_type.compare(0, prefix.size(), prefix) == 0)
    return std::unique_ptr<SearchEngineServiceNew>(
        new GpuVacuumEngine(engine_type.substr(prefix.size()), bloom_enable_factor));
  throw std::runtime_error("Wrong engine type: " + engine_type);     // engine_factory.h:47
}

}  // namespace wsr
