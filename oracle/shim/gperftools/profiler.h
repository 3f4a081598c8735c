// Build shim (test infrastructure): no-op gperftools profiler hooks.
#ifndef WSR_SHIM_PROFILER_H
#define WSR_SHIM_PROFILER_H
inline int ProfilerStart(const char *) { return 0; }
inline void ProfilerStop() {}
#endif
