// C++ host side of the drop-in: class GpuVacuumEngine implements the reference's engine seam
// SearchEngineServiceNew (reference src/qq_mem/src/engine_services.h:14-27) on top of the
// C ABI in include/wsr.h. URL scheme for the factory: "gpu:vacuum_dump:<dir>"
// (the reference's own is "vacuum:vacuum_dump:<dir>", engine_factory.h:21-50).
//
// Built stand-alone, the interface types below mirror the reference's (same names, members,
// defaults: types.h:67-79, 205-218, 259-345). Built inside the reference tree, define
// WSR_WITH_REFERENCE_HEADERS and the reference's own engine_services.h / types.h are used
// instead, so GpuVacuumEngine derives from the reference's real abstract class
// (see INTEGRATION.md).
#ifndef WSR_GPU_VACUUM_ENGINE_H
#define WSR_GPU_VACUUM_ENGINE_H

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "wsr.h"

#ifdef WSR_WITH_REFERENCE_HEADERS
#include "engine_services.h"   // reference: SearchEngineServiceNew, SearchQuery, SearchResult
#else
typedef std::string Term;
typedef std::vector<Term> TermList;
typedef int DocIdType;
typedef double qq_float;

struct DocInfo {};   // only named by the unimplemented AddDocument (as in VacuumEngine)

struct SearchQuery {                        // types.h:205-218
  SearchQuery() {}
  SearchQuery(const TermList &terms_in) : terms(terms_in) {}
  SearchQuery(const TermList &terms_in, const bool &return_snippets_in)
      : terms(terms_in), return_snippets(return_snippets_in) {}
  TermList terms;
  int n_results = 5;
  bool return_snippets = false;
  int n_snippet_passages = 3;
  bool is_phrase = false;
};

struct SearchResultEntry {                  // types.h:259-263
  std::string snippet;
  DocIdType doc_id;
  qq_float doc_score;
};

struct SearchResult {                       // types.h:301-345
  std::vector<SearchResultEntry> entries;
  std::vector<int> doc_freqs;               // doc freq of terms queried
  const SearchResultEntry &operator[](int i) const { return entries[i]; }
  std::size_t Size() const { return entries.size(); }
};

class SearchEngineServiceNew {              // engine_services.h:14-27
 public:
  virtual ~SearchEngineServiceNew() {}
  virtual void AddDocument(const DocInfo doc_info) = 0;
  virtual int LoadLocalDocuments(const std::string &line_doc_path, int n_rows,
                                 const std::string loader) = 0;
  virtual void Load() = 0;
  virtual int TermCount() const = 0;
  virtual std::map<std::string, int> PostinglistSizes(const TermList &terms) = 0;
  virtual SearchResult Search(const SearchQuery &query) = 0;
  virtual void Serialize(std::string dir_path) const = 0;
  virtual void Deserialize(std::string dir_path) = 0;
};
#endif  // WSR_WITH_REFERENCE_HEADERS

namespace wsr {

struct GpuEngineOptions {
  int device = 0;
  // Document-partitioned deployment behind the same seam: when partition_dirs is not
  // empty the engine serves the partitions (one vacuum directory each, local doc ids) as ONE
  // collection through the group API of wsr.h — spread evenly over `devices` (default: `device`),
  // collection statistics exchanged by the library, per-partition top-k merged on the device and
  // across devices over NCCL. engine_dir_path is then only a label.
  std::vector<std::string> partition_dirs;
  std::vector<int> devices;
  int shard = 0, n_shards = 1;       // document partition held by this engine
  int loader_threads = 0;            // 0 = all cores
  bool load_positions = true;        // position column in HBM: needed by phrase queries
  int coalesce_max_batch = 4096;     // Search() callers coalesced per launch
  int coalesce_window_us = 0;        // extra wait for more callers before a launch (0: none —
                                     // callers that arrive while a batch runs form the next one)
  int max_inflight = 2;              // batches in flight at once (host work of one overlaps the
                                     // GPU work of the other)
};

class GpuVacuumEngine : public SearchEngineServiceNew {
 public:
  explicit GpuVacuumEngine(const std::string engine_dir_path, int bloom_enable_factor = 1,
                           GpuEngineOptions opt = GpuEngineOptions());
  ~GpuVacuumEngine();   // the reference base has no virtual destructor

  // ---- SearchEngineServiceNew
  void Load() override;
  int TermCount() const override;
  std::map<std::string, int> PostinglistSizes(const TermList &terms) override;
  // Re-entrant: concurrent callers are coalesced into GPU batches (the reference serves Search
  // from N threads on one shared engine, grpc_server_impl.h:309-328). Leader/follower combining:
  // a caller that finds fewer than max_inflight batches running takes everything queued
  // (itself included), runs it as one batch on its own thread and wakes the others; callers
  // arriving meanwhile form the next batch. A lone caller pays no hand-off at all.
  SearchResult Search(const SearchQuery &query) override;
  // As VacuumEngine (vacuum_engine.h:260-276): not implemented, fatal.
  void AddDocument(const DocInfo doc_info) override;
  int LoadLocalDocuments(const std::string &line_doc_path, int n_rows,
                         const std::string loader) override;
  void Serialize(std::string dir_path) const override;
  void Deserialize(std::string dir_path) override;

  // ---- batch interface used by the replay driver
  std::vector<SearchResult> SearchBatch(const std::vector<SearchQuery> &queries);
  wsr_index *handle() const { return idx_; }
  wsr_group *group() const { return group_; }

 private:
  struct Pending;
  bool ToWsrQuery(const SearchQuery &q, wsr_query *out) const;
  void RunBatch(const std::vector<Pending *> &take);
  void RunGroupBatch(const std::vector<Pending *> &take);
  bool LookupAnyPartition(const std::string &term, uint32_t *df) const;

  std::string dir_;
  int bloom_enable_factor_;
  GpuEngineOptions opt_;
  wsr_index *idx_ = nullptr;     // single-index mode; in group mode: partition 0 (borrowed from group_)
  wsr_group *group_ = nullptr;
  // request coalescer
  std::mutex mu_;
  std::vector<Pending *> pending_;
  int inflight_ = 0;
};

// engine_factory.h:33-50 extended with the gpu: scheme; throws std::runtime_error otherwise.
std::unique_ptr<SearchEngineServiceNew> CreateSearchEngine(std::string engine_type,
                                                           int bloom_enable_factor = 1);

}  // namespace wsr
#endif
