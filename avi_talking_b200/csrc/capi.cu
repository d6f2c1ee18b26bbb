// Library-level entry points: version, thread-local error string, launch counter.
#include "common.cuh"

namespace avi {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace avi

extern "C" int avi_version(void) { return AVI_B200_VERSION; }
extern "C" const char* avi_last_error(void) { return avi::g_err; }
extern "C" int64_t avi_launch_count(void) { return avi::g_launches.load(); }
