// Build shim (test infrastructure): just enough of boost::locale for the
// reference's snippet highlighter to COMPILE. The hot path under test never
// generates snippets (return_snippets=false), so none of this executes.
#ifndef WSR_SHIM_BOOST_LOCALE_H
#define WSR_SHIM_BOOST_LOCALE_H
#include <locale>
#include <string>
#include <vector>
namespace boost { namespace locale {
class generator {
 public:
  std::locale operator()(const std::string &) const { return std::locale::classic(); }
};
namespace boundary {
enum boundary_type { character, word, sentence, line };
enum rule_type { sentence_term = 1, sentence_any = 3 };
class sboundary_point_index {
 public:
  typedef std::vector<std::string::const_iterator>::iterator iterator;
  void map(boundary_type, std::string::const_iterator b, std::string::const_iterator e,
           const std::locale & = std::locale()) {
    pts_.clear();
    pts_.push_back(b);
    for (auto it = b; it != e; ++it)
      if (*it == '.' || *it == '!' || *it == '?') pts_.push_back(it + 1);
    if (pts_.back() != e) pts_.push_back(e);
  }
  void rule(rule_type) {}
  iterator begin() { return pts_.begin(); }
  iterator end() { return pts_.end(); }
  iterator find(std::string::const_iterator p) {
    for (auto it = pts_.begin(); it != pts_.end(); ++it)
      if (*it >= p) return it;
    return pts_.end() - 1;
  }
 private:
  std::vector<std::string::const_iterator> pts_;
};
}  // namespace boundary
}}  // namespace boost::locale
#endif
