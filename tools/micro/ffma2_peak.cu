// Measured fp32 FMA ceiling of the SM (packed FFMA2 with a broadcast scalar operand, the instruction the
// AdaptiveConv inner loop is made of), to put next to the HBM roofline: nvcc -arch=sm_100a -O3 ffma2_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int PACKED>
__global__ void __launch_bounds__(256) k(float2* out, float s, int iters) {
  float2 a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 x = make_float2(1.0001f + s, 0.9999f - s);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (PACKED) a[i] = __ffma2_rn(a[i], x, make_float2(s, s));
      else { a[i].x = fmaf(a[i].x, x.x, s); a[i].y = fmaf(a[i].y, x.y, s); }
    }
  }
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) { r.x += a[i].x; r.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float2* out; cudaMalloc(&out, sizeof(float2) * sms * 8 * 256);
  const int iters = 20000;
  for (int packed = 0; packed < 2; ++packed) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (packed) k<1><<<sms * 8, 256>>>(out, 1e-6f, iters); else k<0><<<sms * 8, 256>>>(out, 1e-6f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * 8 * 256 * iters * 32.0;
      printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"tfma_per_s\": %.2f, \"tflops\": %.2f, \"sms\": %d}\n", packed ? "FFMA2" : "FFMA", ms,
             fma / ms / 1e9, 2 * fma / ms / 1e9, sms);
    }
  }
  return 0;
}
