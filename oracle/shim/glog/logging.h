// Build shim (test infrastructure): minimal stand-in for <glog/logging.h> so the
// UNMODIFIED reference engine sources under /root/reference compile in this
// image, which has no glog. Only the macros the reference uses are provided.
// Semantics kept: LOG(FATAL) aborts; DLOG* vanish under NDEBUG.
#ifndef WSR_SHIM_GLOG_LOGGING_H
#define WSR_SHIM_GLOG_LOGGING_H
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <gflags/gflags.h>  // real glog pulls gflags in; the reference relies on that

extern int FLAGS_minloglevel;   // defined by the tool's main translation unit
extern int FLAGS_logtostderr;

namespace shim_glog {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class LogMessage {
 public:
  LogMessage(int sev, const char *file, int line) : sev_(sev) {
    ss_ << "[" << "IWEF"[sev] << " " << file << ":" << line << "] ";
  }
  ~LogMessage() {
    if (sev_ >= FLAGS_minloglevel || sev_ == FATAL) std::cerr << ss_.str() << std::endl;
    if (sev_ == FATAL) std::abort();
  }
  std::ostream &stream() { return ss_; }
 private:
  int sev_;
  std::ostringstream ss_;
};
struct Voidify { void operator&(std::ostream &) {} };
}  // namespace shim_glog

#define LOG_IF(sev, cond) \
  !(cond) ? (void)0 : shim_glog::Voidify() & shim_glog::LogMessage(shim_glog::sev, __FILE__, __LINE__).stream()
#define LOG(sev) \
  LOG_IF(sev, (shim_glog::sev >= FLAGS_minloglevel || shim_glog::sev == shim_glog::FATAL))
#ifdef NDEBUG
#define DLOG(sev) LOG_IF(sev, false)
#define DLOG_IF(sev, cond) LOG_IF(sev, false && (cond))
#else
#define DLOG(sev) LOG(sev)
#define DLOG_IF(sev, cond) LOG_IF(sev, cond)
#endif

namespace google {
inline void InitGoogleLogging(const char *) {}
}
#endif
