// JBU fix-up MLP on the tensor cores (third generation of kernel B of jbu_filters_v2.cu, padded filter layout only).
//   filters[p, slot] = k[p, slot] + 0.1 * ( W1 gelu(W0 [k; g] + b0) + b1 )[slot]      (upstream JBULearnedRange.fixup_proj:
//   Conv2d(52, 49, 1) -> GELU -> Dropout2d -> Conv2d(49, 49, 1); restated in oracle/jbu.py)
// The 52 -> 49 -> 49 MLP per pixel is the one GEMM-shaped piece of the JBU stage (5.0 kFMA of its 6.5 kFMA per pixel): the
// SIMT kernel spent 1.6 ms on it for B=16 at 512^2, at 62 % of the FFMA2 peak.  Here a tile is 128 pixels:
//   * A operand of layer 1 = the pixel's row of `filters` AS STORED ([7][8] slots, 224 bytes): one TMA box pair drops 128 rows
//     straight into the canonical 128B-swizzled K-major layout; the three guidance values ride in the zero pad slots 7, 15, 23
//     (W0 is re-indexed to that slot order once per launch), so K = 56 = 7 k-steps of `tcgen05.mma.kind::tf32`;
//   * precision: the tensor pipe truncates fp32 operands to tf32, so the activations go in as hi + lo (lo = x - trunc(x), a
//     second A tile; two MMAs per k-step) and the weights are rounded to tf32 once (a fixed 2^-12 relative perturbation of W,
//     scaled by the 0.1 in front of the MLP: ~3e-5 of the stack's output against a 1e-3 contract);
//   * epilogue 1 (one thread per pixel row): TMEM -> + b0 -> erf-form GELU (erf to 1.5e-7) -> hi / lo tiles of the hidden layer back into
//     shared memory in operand layout; layer 2 has its OUTPUT columns in slot order, so epilogue 2 adds 0.1 * o onto k in
//     place in the shared-memory tile and one TMA box pair stores it;
//   * a CTA runs TWO independent 128-thread groups (own tiles, barriers, TMEM columns, issuing thread): one group's MMAs and
//     TMA round trips overlap the other's epilogues.  224 KB of shared memory: one CTA per SM, persistent over the tiles.
#include "tc_common.cuh"

namespace isp {
namespace jf3 {

constexpr int TP = 128;               // pixels per tile
constexpr int ATOM_A = TP * 128;      // one swizzle atom of an activation tile: 128 rows x 128 B
constexpr int A_BYTES = 2 * ATOM_A;   // K padded to 64 floats = two atoms
constexpr int ATOM_W = 64 * 128;      // weights: 64 rows (N) x 128 B
constexpr int W_BYTES = 2 * ATOM_W;
constexpr int GROUPS = 2, GT = 128;
constexpr int SMEM_BYTES = GROUPS * 3 * A_BYTES + 2 * W_BYTES + 1024;  // + alignment slack
constexpr int KSTEPS = 7;             // 56 floats

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// byte offset of 16-byte unit `unit` (0..7) of row `row` inside one 128B-swizzled atom
__device__ __forceinline__ uint32_t sw_off(int row, int unit) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((unit ^ (row & 7)) << 4));
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(m), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }
__device__ __forceinline__ int slot_tap(int slot) { return (slot >> 3) * 7 + (slot & 7); }  // valid for (slot & 7) < 7

__global__ void __launch_bounds__(GROUPS * GT, 1)
jbu_fixup_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float4* __restrict__ g4, long long npix,
                    const float* __restrict__ fw0, const float* __restrict__ fb0, const float* __restrict__ fw1,
                    const float* __restrict__ fb1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* W0 = tiles + GROUPS * 3 * A_BYTES;
  uint8_t* W1 = W0 + W_BYTES;
  __shared__ __align__(8) uint64_t full_bar[GROUPS], mma1_bar[GROUPS], mma2_bar[GROUPS];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float b0_s[64], b1_s[64];

  const int tid = threadIdx.x, g = tid >> 7, t = tid & (GT - 1), wg = t >> 5;
  if (tid == 0) {
    for (int i = 0; i < GROUPS; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&mma1_bar[i], 1); tc::mbar_init(&mma2_bar[i], 1); }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmF);
  }
  if (tid < 32) tc::tmem_alloc(&tmem_base_s, 256);
  // weights -> operand layout (rounded to tf32), biases
  for (int idx = tid; idx < 64 * 64; idx += GROUPS * GT) {
    const int n = idx >> 6, k = idx & 63;
    float w0 = 0.f, w1 = 0.f;
    if (n < 49 && k < 56) {          // layer 1: row n = hidden unit, column k = input slot
      const int i = k >> 3, c = k & 7;
      if (c < 7) w0 = __ldg(fw0 + n * 52 + i * 7 + c);
      else if (i < 3) w0 = __ldg(fw0 + n * 52 + 49 + i);
    }
    if (n < 56 && (n & 7) < 7 && k < 49) w1 = __ldg(fw1 + slot_tap(n) * 49 + k);  // layer 2: row n = output SLOT, column k = hidden unit
    const uint32_t off = (uint32_t)(k >> 5) * ATOM_W + sw_off(n, (k & 31) >> 2) + (k & 3) * 4;
    *reinterpret_cast<float*>(W0 + off) = tf32_rn(w0);
    *reinterpret_cast<float*>(W1 + off) = tf32_rn(w1);
  }
  if (tid < 64) {
    b0_s[tid] = tid < 49 ? __ldg(fb0 + tid) : 0.f;
    b1_s[tid] = (tid < 56 && (tid & 7) < 7) ? __ldg(fb1 + slot_tap(tid)) : 0.f;
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  uint8_t* Ak = tiles + g * 3 * A_BYTES;  // k (+ g in the pad slots); becomes the output tile
  uint8_t* Alo = Ak + A_BYTES;            // low parts: of the input, then of the hidden layer
  uint8_t* Hhi = Ak + 2 * A_BYTES;        // hidden layer
  const uint32_t d1 = tmem_base + g * 128, d2 = d1 + 64;
  const uint32_t lane_addr = (uint32_t)(wg * 32) << 16;
  const uint32_t idesc = idesc_tf32(128, 64);
  const int row = t;
  const long long ntiles = (npix + TP - 1) / TP;
  uint32_t ph = 0;
  for (long long tile = (long long)blockIdx.x * GROUPS + g; tile < ntiles; tile += (long long)gridDim.x * GROUPS, ph ^= 1) {
    const long long pix0 = tile * TP;
    if (t == 0) {
      tc::tma_store_wait_read<0>();  // the previous tile's store has read Ak
      tc::mbar_arrive_expect_tx(&full_bar[g], A_BYTES);
      tc::tma_load_2d(Ak, &tmF, &full_bar[g], 0, (int)pix0);
      tc::tma_load_2d(Ak + ATOM_A, &tmF, &full_bar[g], 32, (int)pix0);  // columns 56..63 are out of bounds: zero-filled
      const long long nxt = tile + (long long)gridDim.x * GROUPS;  // the group's next tile: HBM -> L2 while this one is worked on
      if (nxt < ntiles) {
        tma_prefetch_l2_2d(&tmF, 0, (int)(nxt * TP));
        tma_prefetch_l2_2d(&tmF, 32, (int)(nxt * TP));
      }
    }
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix0 + row < npix) gv = __ldg(g4 + pix0 + row);
    tc::mbar_wait(&full_bar[g], ph);
    // guidance into the pad slots 7 / 15 / 23, low parts of the row
#pragma unroll
    for (int u = 0; u < 14; ++u) {
      const uint32_t off = (uint32_t)(u >> 3) * ATOM_A + sw_off(row, u & 7);
      float4 v = *reinterpret_cast<const float4*>(Ak + off);
      if (u == 1) v.w = gv.x;
      if (u == 3) v.w = gv.y;
      if (u == 5) v.w = gv.z;
      if (u == 1 || u == 3 || u == 5) *reinterpret_cast<float4*>(Ak + off) = v;
      *reinterpret_cast<float4*>(Alo + off) =
          make_float4(v.x - tf32_trunc(v.x), v.y - tf32_trunc(v.y), v.z - tf32_trunc(v.z), v.w - tf32_trunc(v.w));
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    group_sync(g);
    if (t == 0) {
      tc::tc_fence_after();
#pragma unroll
      for (int s = 0; s < KSTEPS; ++s) {
        const uint32_t ao = (uint32_t)(s >> 2) * ATOM_A, wo = (uint32_t)(s >> 2) * ATOM_W;
        const uint64_t dw = tc::smem_desc_k_sw128(tc::smem_u32(W0 + wo)) + 2 * (s & 3);
        umma_tf32(d1, tc::smem_desc_k_sw128(tc::smem_u32(Ak + ao)) + 2 * (s & 3), dw, idesc, s ? 1u : 0u);
        umma_tf32(d1, tc::smem_desc_k_sw128(tc::smem_u32(Alo + ao)) + 2 * (s & 3), dw, idesc, 1u);
      }
      tc::umma_commit(&mma1_bar[g]);
    }
    tc::mbar_wait(&mma1_bar[g], ph);
    tc::tc_fence_after();
    {  // hidden layer: h = gelu(acc + b0) -> hi tile (full fp32: the tensor pipe truncates) and lo tile
      uint32_t a0[32], a1[32];
      tc::tmem_ld32(d1 + lane_addr, a0);
      tc::tmem_ld32(d1 + lane_addr + 32, a1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 14; ++u) {
        float h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = 4 * u + e;
          const float acc = __uint_as_float(n < 32 ? a0[n & 31] : a1[n & 31]);
          h[e] = n < 49 ? gelu_erf_fast(acc + b0_s[n]) : 0.f;
        }
        const uint32_t off = (uint32_t)(u >> 3) * ATOM_A + sw_off(row, u & 7);
        *reinterpret_cast<float4*>(Hhi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(Alo + off) =
            make_float4(h[0] - tf32_trunc(h[0]), h[1] - tf32_trunc(h[1]), h[2] - tf32_trunc(h[2]), h[3] - tf32_trunc(h[3]));
      }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    group_sync(g);
    if (t == 0) {
      tc::tc_fence_after();
#pragma unroll
      for (int s = 0; s < KSTEPS; ++s) {
        const uint32_t ao = (uint32_t)(s >> 2) * ATOM_A, wo = (uint32_t)(s >> 2) * ATOM_W;
        const uint64_t dw = tc::smem_desc_k_sw128(tc::smem_u32(W1 + wo)) + 2 * (s & 3);
        umma_tf32(d2, tc::smem_desc_k_sw128(tc::smem_u32(Hhi + ao)) + 2 * (s & 3), dw, idesc, s ? 1u : 0u);
        umma_tf32(d2, tc::smem_desc_k_sw128(tc::smem_u32(Alo + ao)) + 2 * (s & 3), dw, idesc, 1u);
      }
      tc::umma_commit(&mma2_bar[g]);
    }
    tc::mbar_wait(&mma2_bar[g], ph);
    tc::tc_fence_after();
    {  // filters = k + 0.1 * (acc + b1), in place in the tile (slot order on both sides); pad slots are written as zeros
      uint32_t a0[32], a1[32];
      tc::tmem_ld32(d2 + lane_addr, a0);
      tc::tmem_ld32(d2 + lane_addr + 32, a1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 14; ++u) {
        const uint32_t off = (uint32_t)(u >> 3) * ATOM_A + sw_off(row, u & 7);
        const float4 k4 = *reinterpret_cast<const float4*>(Ak + off);
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = 4 * u + e;
          const float acc = __uint_as_float(n < 32 ? a0[n & 31] : a1[n & 31]);
          o[e] = (n & 7) == 7 ? 0.f : fmaf(0.1f, acc + b1_s[n], kk[e]);
        }
        *reinterpret_cast<float4*>(Ak + off) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    group_sync(g);
    if (t == 0) {
      tc::tma_store_2d(&tmF, Ak, 0, (int)pix0);
      tc::tma_store_2d(&tmF, Ak + ATOM_A, 32, (int)pix0);  // clipped at column 56 and at the last pixel
      tc::tma_store_commit();
    }
  }
  if (t == 0) tc::tma_store_wait_all();
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace jf3

// filters [npix][56] (the padded layout, k already written by jbu_range_kernel) += 0.1 * fixup MLP; called by isp_jbu_filters
int jbu_fixup_tc_launch(const float* g, float* filters, long long npix, const float* fw0, const float* fb0, const float* fw1,
                        const float* fb1, cudaStream_t stream) {
  ISP_REQUIRE(npix > 0 && npix < (1LL << 31), ISP_ERR_UNSUPPORTED, "jbu_filters: pixel count %lld", npix);
  CUtensorMap tmF;
  const uint64_t dims[2] = {56, (uint64_t)npix}, str[2] = {4, 56 * 4};
  const uint32_t box[2] = {32, (uint32_t)jf3::TP};
  if (int e = make_tmap(&tmF, 4, filters, 2, dims, str, box, "jbu_filters(fixup)", true)) return e;
  if (int e = ensure_dynamic_smem((const void*)jf3::jbu_fixup_tc_kernel, jf3::SMEM_BYTES)) return e;
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  const long long ntiles = (npix + jf3::TP - 1) / jf3::TP;
  const int grid = (int)((ntiles + jf3::GROUPS - 1) / jf3::GROUPS < sms ? (ntiles + jf3::GROUPS - 1) / jf3::GROUPS : sms);
  jf3::jbu_fixup_tc_kernel<<<grid, jf3::GROUPS * jf3::GT, jf3::SMEM_BYTES, stream>>>(
      tmF, reinterpret_cast<const float4*>(g), npix, fw0, fb0, fw1, fb1);
  ISP_CHECK_LAUNCH("jbu_fixup_tc_kernel");
  return ISP_OK;
}

}  // namespace isp
