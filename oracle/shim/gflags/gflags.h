// Build shim (test infrastructure): stand-in for <gflags/gflags.h>. Flags become
// plain globals FLAGS_<name>; there is no command-line parsing.
#ifndef WSR_SHIM_GFLAGS_H
#define WSR_SHIM_GFLAGS_H
#include <cstdint>
#include <string>
#define DEFINE_bool(name, val, doc) bool FLAGS_##name = (val)
#define DEFINE_int32(name, val, doc) int32_t FLAGS_##name = (val)
#define DEFINE_double(name, val, doc) double FLAGS_##name = (val)
#define DEFINE_string(name, val, doc) std::string FLAGS_##name = (val)
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
namespace gflags {
inline void ParseCommandLineFlags(int *, char ***, bool) {}
}
namespace google {
inline void ParseCommandLineFlags(int *, char ***, bool) {}
}
#endif
