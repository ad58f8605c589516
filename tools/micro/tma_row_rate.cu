// How fast does one SM's TMA unit serve 2-D box loads, as a function of the box shape?  One CTA per SM, one thread issues
// `cp.async.bulk.tensor.2d` loads of a [rows x cols] bf16 box into a ring of `in_flight` stages and waits for them; the
// source matrix is small enough to stay in L2 (the question is the unit's service rate, not DRAM).  Prints cycles per box,
// per box row and bytes per cycle for boxes of equal size (16 KB) but different row length, and for the GEMM kernel's two
// operand shapes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_row_rate tma_row_rate.cu && ./tma_row_rate
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int box_rows, int box_cols, int nboxes_rows,
                                            int nboxes_cols, int iters, uint32_t box_bytes, int depth, int issuers, int same_lane_warp, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4][32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 128; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0][0] + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // issuer w: lane 0 of warp w (different warps), or lane w of warp 0 (same warp, divergent lanes)
  const int w = same_lane_warp ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
  const bool is_issuer = same_lane_warp ? (threadIdx.x < issuers) : ((threadIdx.x & 31) == 0 && w < issuers);
  if (is_issuer) {
    uint8_t* my = smem + (size_t)w * depth * box_bytes;
    const int nb = nboxes_rows * nboxes_cols;
    uint32_t b = blockIdx.x * 7919u + w * 1237u;
    const long long t0 = clock64();
    int s = 0, round = 0;
    for (int it = 0; it < iters + depth; ++it) {
      if (it >= depth) {  // the load issued `depth` iterations ago into this stage has landed
        const uint32_t par = (round - 1) & 1;
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                       : "=r"(ok) : "r"(s32(&bar[w][s])), "r"(par) : "memory");
      }
      if (it < iters) {
        b = (b + 1) % (uint32_t)nb;
        const int c0 = (int)(b % nboxes_cols) * box_cols, c1 = (int)(b / nboxes_cols) * box_rows;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[w][s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(my + s * box_bytes)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(&bar[w][s])), "r"(c0), "r"(c1)
                     : "memory");
      }
      if (++s == depth) { s = 0; ++round; }
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  encode_fn enc = (encode_fn)fp;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int R = 16384, C = 512;  // 16 MB bf16 matrix: L2-resident
  void* d;
  cudaMalloc(&d, (size_t)R * C * 2);
  cudaMemset(d, 0, (size_t)R * C * 2);
  long long* cyc;
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
  struct Shape { int rows, cols; CUtensorMapSwizzle sw; const char* what; int depth, issuers, same; };
  const Shape shapes[] = {
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B, swizzle 128B (A tile of the GEMM)", 4, 1, 0},
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B, swizzle 128B (A tile of the GEMM)", 12, 1, 0},
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B, swizzle 128B", 6, 2, 0},
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B, swizzle 128B", 3, 4, 0},
      {128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B, swizzle 128B", 6, 2, 1},
      {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, " 32 rows x 128 B, swizzle 128B (epilogue chunk)", 12, 1, 0},
      {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, " 32 rows x 128 B, swizzle 128B", 12, 2, 0},
      {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, " 32 rows x 128 B, swizzle 128B", 12, 4, 0},
      {32, 64, CU_TENSOR_MAP_SWIZZLE_128B, " 32 rows x 128 B, swizzle 128B", 12, 4, 1},
      {256, 32, CU_TENSOR_MAP_SWIZZLE_64B, "256 rows x  64 B, swizzle 64B", 12, 1, 0},
      {32, 256, CU_TENSOR_MAP_SWIZZLE_NONE, " 32 rows x 512 B, no swizzle", 12, 1, 0},
  };

  const int iters = 4000;
  printf("# one thread per SM issuing 2-D TMA box loads (bf16, L2-resident source), in_flight loads outstanding, %d SMs, %d loads each\n", sms, iters);
  for (const Shape& sh : shapes) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, str[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)sh.cols, (cuuint32_t)sh.rows}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sh.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed for %s (%d)\n", sh.what, (int)r); continue; }
    const uint32_t bytes = (uint32_t)sh.rows * sh.cols * 2;
    for (int rep = 0; rep < 2; ++rep) {
      k<<<sms, 128, (size_t)sh.depth * bytes * sh.issuers>>>(tm, sh.rows, sh.cols, R / sh.rows, C / sh.cols, iters, bytes, sh.depth, sh.issuers, sh.same, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)cyc[i];
    avg /= sms;
    const double per_box = avg / iters;
    printf("{\"box\": \"%s\", \"issuing_threads\": %d, \"same_warp\": %d, \"in_flight_per_thread\": %d, \"bytes\": %u, \"cycles_per_box_per_thread\": %.1f, \"cycles_per_box_sm\": %.1f, \"bytes_per_cycle_per_sm\": %.1f}\n",
           sh.what, sh.issuers, sh.same, sh.depth, bytes, per_box, per_box / sh.issuers, bytes * sh.issuers / per_box);
  }
  return 0;
}
