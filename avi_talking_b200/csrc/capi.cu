// Library-level entry points: version, thread-local error string, launch counter, the dynamic-tile switch.
#include <cstdlib>

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

// AVI_DYNAMIC_TILES=<bit mask> in the environment (1 GEMM, 2 conv0). Default 0 = the static walk: measured on the configs[1] step the
// dynamic scheduler is correct (bit-identical) but not faster (profiles/r2/clc_dynamic_tiles_ab.txt); it is for callers that share the
// GPU with other work for longer than this step does.
static std::atomic<int> g_dynamic_tiles{getenv("AVI_DYNAMIC_TILES") != nullptr ? atoi(getenv("AVI_DYNAMIC_TILES")) : 0};
int dynamic_tiles_mask() { return g_dynamic_tiles.load(std::memory_order_relaxed); }

// programmatic dependent launch of the GEMM / attention / LayerNorm kernels (common.cuh launch_pdl)
static std::atomic<int> g_pdl{getenv("AVI_PDL") != nullptr ? atoi(getenv("AVI_PDL")) : 0};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

}  // namespace avi

extern "C" int avi_set_pdl(int32_t on) { return avi::g_pdl.exchange(on ? 1 : 0, std::memory_order_relaxed); }
extern "C" int avi_set_dynamic_tiles(int32_t mask) { return avi::g_dynamic_tiles.exchange(mask, std::memory_order_relaxed); }
extern "C" int avi_version(void) { return AVI_B200_VERSION; }
extern "C" const char* avi_last_error(void) { return avi::g_err; }
extern "C" int64_t avi_launch_count(void) { return avi::g_launches.load(); }
