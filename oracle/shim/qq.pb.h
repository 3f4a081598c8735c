// Build shim (test infrastructure): hand-written value types with the accessor
// names protoc would generate for the reference's qq.proto messages, so that
// types.h / utils.cc compile without protobuf. Field list per qq.proto:42-56.
#ifndef WSR_SHIM_QQ_PB_H
#define WSR_SHIM_QQ_PB_H
// protobuf-generated headers pull these in transitively; the reference relies on that.
#include <unistd.h>
#include <algorithm>
#include <functional>
#include <iostream>
#include <memory>
#include <cctype>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <climits>
#include <cassert>
#include <cmath>
#include <map>
#include <set>
#include <unordered_map>
#include <cstdint>
#include <string>
#include <vector>
namespace qq {
class SearchRequest {
 public:
  int terms_size() const { return (int)terms_.size(); }
  const std::string &terms(int i) const { return terms_[i]; }
  void add_terms(const std::string &t) { terms_.push_back(t); }
  void clear_terms() { terms_.clear(); }
  int32_t n_results() const { return n_results_; }
  void set_n_results(int32_t v) { n_results_ = v; }
  bool return_snippets() const { return return_snippets_; }
  void set_return_snippets(bool v) { return_snippets_ = v; }
  int32_t n_snippet_passages() const { return n_snippet_passages_; }
  void set_n_snippet_passages(int32_t v) { n_snippet_passages_ = v; }
  bool is_phrase() const { return is_phrase_; }
  void set_is_phrase(bool v) { is_phrase_ = v; }
 private:
  std::vector<std::string> terms_;
  int32_t n_results_ = 0;
  bool return_snippets_ = false;
  int32_t n_snippet_passages_ = 0;
  bool is_phrase_ = false;
};
class SearchReplyEntry {
 public:
  int32_t doc_id() const { return doc_id_; }
  void set_doc_id(int32_t v) { doc_id_ = v; }
  const std::string &snippet() const { return snippet_; }
  void set_snippet(const std::string &s) { snippet_ = s; }
  float doc_score() const { return doc_score_; }
  void set_doc_score(float v) { doc_score_ = v; }
 private:
  int32_t doc_id_ = 0;
  std::string snippet_;
  float doc_score_ = 0;
};
class SearchReply {
 public:
  int entries_size() const { return (int)entries_.size(); }
  const SearchReplyEntry &entries(int i) const { return entries_[i]; }
  SearchReplyEntry *add_entries() { entries_.emplace_back(); return &entries_.back(); }
  void clear_entries() { entries_.clear(); }
 private:
  std::vector<SearchReplyEntry> entries_;
};
}  // namespace qq
#endif
