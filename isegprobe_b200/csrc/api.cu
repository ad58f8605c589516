// Error text, version and launch counter of libisp_b200.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace isp {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace isp

extern "C" int isp_version(void) { return ISP_ABI_VERSION; }
extern "C" const char* isp_last_error(void) { return isp::g_err; }
extern "C" unsigned long long isp_launch_count(void) { return isp::g_launches.load(); }
