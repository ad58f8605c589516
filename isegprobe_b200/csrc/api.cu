// Error text, version and launch counter of libisp_b200.
#include <stdarg.h>
#include <atomic>
#include <mutex>

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

namespace {
constexpr int kMaxDev = 64, kMaxOptIn = 512;
std::mutex g_dev_mu;
int g_sms[kMaxDev] = {};
struct OptIn { int dev; const void* fn; int bytes; };
OptIn g_optin[kMaxOptIn];
int g_noptin = 0;
}  // namespace

int device_sm_count(int* sms) {
  int dev = 0;
  ISP_CUDA(cudaGetDevice(&dev));
  ISP_REQUIRE(dev >= 0 && dev < kMaxDev, ISP_ERR_UNSUPPORTED, "device ordinal %d not supported", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_sms[dev]) ISP_CUDA(cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  *sms = g_sms[dev];
  return ISP_OK;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
  int dev = 0;
  ISP_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  for (int i = 0; i < g_noptin; ++i)
    if (g_optin[i].dev == dev && g_optin[i].fn == kernel && g_optin[i].bytes >= bytes) return ISP_OK;
  ISP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (g_noptin < kMaxOptIn) g_optin[g_noptin++] = {dev, kernel, bytes};  // table full: the attribute is simply re-applied
  return ISP_OK;
}
}  // namespace isp

extern "C" int isp_version(void) { return ISP_ABI_VERSION; }
extern "C" const char* isp_last_error(void) { return isp::g_err; }
extern "C" unsigned long long isp_launch_count(void) { return isp::g_launches.load(); }
