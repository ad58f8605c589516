// What would a tensor-core AdaptiveConv cost?    The banded formulation is
//   out[c, n] = sum_i sum_col x[c, y+i, x0+col] * Band_i[col, n]     M = 128 channels, N = pixels of one output row,
//   K = 7 input rows x (N + 8) columns, three kind::tf32 MMAs per k-step for the 2-term split under the 1e-3 contract,
// i.e. very many SMALL-N MMAs.  This program measures what one SM sustains for back-to-back
// `tcgen05.mma.cta_group::1.kind::tf32` of shape 128 x N x 8 issued by one thread, with the A operand in shared memory
// (SS) and in tensor memory (TS), and prints cycles per MMA plus the projected time of the stage-512 AdaptiveConv launch
// of the bench (B=16, 512^2, C=384: 1.61e9 outputs over 148 SMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32_rate umma_tf32_rate.cu && ./umma_tf32_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, 128-byte swizzle: rows of 32 tf32 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor: tf32 x tf32 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool TS>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];  // A: 128 rows x 128 B = 16 KB at 0; B: 256 rows x 128 B = 32 KB at 16 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (TS) {  // zero the A columns (256..319) so the products are finite
    const uint32_t z = 0;
    for (int c = 0; c < 64; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((threadIdx.x >> 5) * 32u << 16) + 256 + c), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    const uint64_t adesc = desc_sw128(s32(smem)), bdesc = desc_sw128(s32(smem + 16 * 1024));
    const uint32_t idesc = idesc_tf32(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t ks = it & 3;  // 4 k-steps of 32 bytes inside the 128-byte swizzle atom
      const uint32_t acc = it ? 1u : 0u;
      if (TS) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm),
            "r"(tm + 256 + (it & 7) * 8), "l"(bdesc + 2 * ks), "r"(idesc), "r"(acc)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
            "l"(adesc + 2 * ks), "l"(bdesc + 2 * ks), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
      if (!ok && ++spins > (1u << 26)) __trap();
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  long long* cyc;
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 8192;
  const double outputs = 16.0 * 512 * 512 * 384;  // stage-512 AdaptiveConv launch of the bench
  printf("# back-to-back tcgen05.mma kind::tf32 128 x N x 8 from one thread per SM, %d SMs, %d MMAs each, max clock %.0f MHz\n", sms,
         iters, khz / 1000.0);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 32, 64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (ts) k<true><<<sms, 128, 48 * 1024>>>(N, iters, cyc);
        else k<false><<<sms, 128, 48 * 1024>>>(N, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      }
      double avg = 0;
      for (int i = 0; i < sms; ++i) avg += (double)cyc[i];
      avg /= sms;
      const double per = avg / iters;
      // banded AdaptiveConv with N pixels per MMA: 7 rows x ceil((N + 6) / 8) k-steps x 3 split terms per 128 x N outputs
      const int ksteps = 7 * ((N + 6 + 7) / 8);
      const double mmas_per_out = 3.0 * ksteps / (128.0 * N);
      const double clk_per_out = mmas_per_out * per;
      const double ms = outputs * clk_per_out / sms / 1.8e9 * 1e3;  // 1.8 GHz: the clock the bench boxes sustain
      printf("{\"A_operand\": \"%s\", \"N\": %d, \"cycles_per_mma\": %.1f, \"tf32_tflops_at_1.8GHz\": %.0f, \"banded_ksteps_per_row_group\": %d, "
             "\"mmas_per_output\": %.5f, \"cycles_per_output_per_sm\": %.3f, \"projected_ms_stage512_b16\": %.2f}\n",
             ts ? "tmem" : "smem", N, per, 2.0 * 128 * N * 8 / per * 1.8e9 * sms / 1e12, ksteps, mmas_per_out, clk_per_out, ms);
    }
  return 0;
}
