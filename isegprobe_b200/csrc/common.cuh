// Shared helpers for libisp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/isp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libisp_b200 targets sm_100a only (no multi-arch dispatch)"
#endif

namespace isp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Per-device launch state (api.cu).  One process may drive several GPUs, so nothing about a device is cached in a
// process-wide static: the SM count is looked up per current device, and the > 48 KB dynamic shared-memory opt-in of a
// kernel is applied once per (device, kernel).  Both are thread-safe.  Return 0 / an ISP_ERR_* code (error text set).
int device_sm_count(int* sms);
int ensure_dynamic_smem(const void* kernel, int bytes);

#define ISP_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) {                            \
      ::isp::set_error(__VA_ARGS__);          \
      return (code);                          \
    }                                         \
  } while (0)

#define ISP_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      ::isp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return ISP_ERR_CUDA;                                                  \
    }                                                                       \
    ::isp::count_launch(1);                                                 \
  } while (0)

#define ISP_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      ::isp::set_error("%s failed: %s", #call, cudaGetErrorString(e__));    \
      return ISP_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline cudaStream_t as_stream(isp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// exact (erf) GELU, as torch.nn.GELU() default
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// the same function with erf to ~1.5e-7 absolute (Abramowitz-Stegun 7.1.26): 1 RCP + 1 EX2 + 7 FMA instead of erff's ~40
// instructions, for kernels whose instruction stream is dominated by the GELU (JBU range / fix-up projections)
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = x * 0.70710678118654752440f, az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-az * az * 1.4426950408889634f));
  const float erfz = copysignf(fmaf(-poly * t, e, 1.f), z);
  return 0.5f * x * (1.f + erfz);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {  // torch 'reflect' padding (no edge repeat)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

}  // namespace isp
