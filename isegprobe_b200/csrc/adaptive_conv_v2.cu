// AdaptiveConv, second generation (the one isp_adaptive_conv_fwd dispatches to).
//   out[b,y,x,c] = sum_{i,j<7} in[b,y+i,x+j,c] * filt[b,y,x,i*7+j]        (NHWC, fp32)
//
// v1 (adaptive_conv.cu) measured 2.0 TB/s: ncu showed it bound by shared-memory wavefronts
// (a broadcast LDS.128 costs 2 wavefronts, so the 49 weights alone cost 26 wavefronts per
// 64-channel pixel) and by the instructions of its cp.async fill loop.  v2 changes:
//   * half-warp = one output row, lane = 4 channels (one float4): a weight fetched from smem
//     now feeds 2 rows x 64 channels per wavefront pair, i.e. half the weight traffic per FMA;
//   * row passes: for each of the 7 filter rows the warp slides a 7-wide float4 window along
//     its 16 output pixels (1 new LDS.128 per pixel), accumulators for all 16 pixels stay in
//     registers (64) -- instead of a 7x7 window per pixel;
//   * the (8+6)x(16+6)x64-channel input tile arrives by ONE 4-D TMA box per channel group
//     (un-swizzled, OOB zero-filled): no fill instructions.  The tile is single-buffered
//     (105 KB per CTA) so that TWO CTAs are resident per SM: one computes while the other's TMA
//     is in flight, and every SMSP has two warps to hide shared-memory latency;
//   * every FMA is still a packed FFMA2 with the weight as scalar-broadcast operand.
#include "tc_common.cuh"

namespace isp {
namespace ac2 {

constexpr int TH = 8, TW = 16, CG = 64;
constexpr int PH = TH + 6, PW = TW + 6;
constexpr int kThreads = 128;
constexpr uint32_t kStageBytes = PH * PW * CG * 4;        // 78848
constexpr int kWRow = 8;                                   // taps per filter row in smem (7 + pad)
constexpr uint32_t kWBytes = TH * TW * 7 * kWRow * 4;      // 28672
constexpr uint32_t kSmem = kStageBytes + kWBytes;          // 107520 -> two CTAs per SM

__global__ void __launch_bounds__(kThreads, 2)
adaptive_conv_v2_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmF,
                        const float* __restrict__ filt, float* __restrict__ out, int H, int W, int C, int filt_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar, filt_bar;
  float* in_s = reinterpret_cast<float*>(smem);                       // [PH][PW][64]
  float* w_s = reinterpret_cast<float*>(smem + kStageBytes);          // [TH*TW][7][8]
  const int b = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const int ncg = C / CG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmIn);
    tc::mbar_init(&full_bar, 1);
    tc::mbar_init(&filt_bar, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(&full_bar, kStageBytes);
    tc::tma_load_4d(smem, &tmIn, &full_bar, 0, tx0, ty0, b);
    if (filt_ld == 56) {  // row-padded filters [B,H,W,7,8]: the smem image is exactly one TMA box
      tc::prefetch_tmap(&tmF);
      tc::mbar_arrive_expect_tx(&filt_bar, kWBytes);
      tc::tma_load_4d(smem + kStageBytes, &tmF, &filt_bar, 0, tx0, ty0, b);
    }
  }
  // per-pixel filters of the tile -> smem [pixel][filter row][8] (rows padded to two float4)
  // (each tile row is one contiguous run of 16*49 floats in global memory: coalesced loads,
  //  7 independent loads in flight per thread)
  if (filt_ld != 56) {
    constexpr int kRun = TW * 49;              // 784 floats per tile row
    constexpr int kPerThread = (TH * kRun + kThreads - 1) / kThreads;  // 49
    for (int it0 = 0; it0 < kPerThread; it0 += 7) {
      float v[7];
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int e = (it0 + u) * kThreads + threadIdx.x;
        const int r = e / kRun, o = e - r * kRun;
        const int y = ty0 + r, x = tx0 + o / 49;
        v[u] = (e < TH * kRun && y < H && x < W) ? __ldg(filt + (((size_t)b * H + y) * W + tx0) * 49 + o) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int e = (it0 + u) * kThreads + threadIdx.x;
        if (e < TH * kRun) {
          const int r = e / kRun, o = e - r * kRun;
          const int px = o / 49, tap = o - px * 49;
          w_s[((r * TW + px) * 7 + tap / 7) * kWRow + tap % 7] = v[u];
        }
      }
    }
    for (int idx = threadIdx.x; idx < TH * TW * 7; idx += kThreads) w_s[idx * kWRow + 7] = 0.f;  // pad tap
  }
  __syncthreads();
  if (filt_ld == 56) tc::mbar_wait(&filt_bar, 0);

  const int hl = lane & 15;               // float4 index inside the 64-channel group
  const int row = warp * 2 + (lane >> 4);  // output row of the tile owned by this half-warp
  const int y = ty0 + row;
  for (int cg = 0; cg < ncg; ++cg) {
    tc::mbar_wait(&full_bar, cg & 1);
    const float4* tile = reinterpret_cast<const float4*>(in_s) + hl;
    float2 acc[TW][2];
#pragma unroll
    for (int x = 0; x < TW; ++x) acc[x][0] = acc[x][1] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
      const float4* rowp = tile + (row + i) * PW * 16;
      const float4* wrow = reinterpret_cast<const float4*>(w_s + (row * TW) * 7 * kWRow + i * kWRow);
      float4 win[7];
#pragma unroll
      for (int j = 0; j < 6; ++j) win[j] = rowp[j * 16];
#pragma unroll
      for (int x = 0; x < TW; ++x) {
        win[(x + 6) % 7] = rowp[(x + 6) * 16];
        const float4 wa = wrow[x * (7 * kWRow / 4)], wb = wrow[x * (7 * kWRow / 4) + 1];
        const float wj[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float4 v = win[(x + j) % 7];
          acc[x][0] = __ffma2_rn(make_float2(v.x, v.y), make_float2(wj[j], wj[j]), acc[x][0]);
          acc[x][1] = __ffma2_rn(make_float2(v.z, v.w), make_float2(wj[j], wj[j]), acc[x][1]);
        }
      }
    }
    if (y < H) {
      float* orow = out + (((size_t)b * H + y) * W + tx0) * C + cg * CG + 4 * hl;
#pragma unroll
      for (int x = 0; x < TW; ++x)
        if (tx0 + x < W)
          *reinterpret_cast<float4*>(orow + (size_t)x * C) = make_float4(acc[x][0].x, acc[x][0].y, acc[x][1].x, acc[x][1].y);
    }
    __syncthreads();  // tile fully consumed (accumulators are in registers; stores may still be in flight)
    if (threadIdx.x == 0 && cg + 1 < ncg) {
      tc::mbar_arrive_expect_tx(&full_bar, kStageBytes);
      tc::tma_load_4d(smem, &tmIn, &full_bar, (cg + 1) * CG, tx0, ty0, b);
    }
  }
}

}  // namespace ac2

// Third generation: the same data movement, but every half-warp owns TWO adjacent output rows of an 8-pixel-wide
// strip (tile = 16 rows x 8 pixels).  ncu on v2: shared-memory pipe 72 % busy (155 wavefronts per 224 FFMA2 of a
// filter-row sweep, 88 of them the input window loads) against 59 % of the FFMA2 peak -- the kernel was bound by
// shared-memory bandwidth.  An input row loaded into the sliding window now feeds filter row i of the upper output
// row AND filter row i-1 of the lower one: 0.29 instead of 0.39 input wavefronts per FFMA2.
namespace ac3 {

constexpr int TH = 16, TW = 8, CG = 64;
constexpr int PH = TH + 6, PW = TW + 6;                    // 22 x 14
constexpr int kThreads = 128;
constexpr uint32_t kStageBytes = PH * PW * CG * 4;        // 78848
constexpr int kWRow = 8;
constexpr uint32_t kWBytes = TH * TW * 7 * kWRow * 4;      // 28672
constexpr uint32_t kSmem = kStageBytes + kWBytes;          // 107520 -> two CTAs per SM

// One input row of the strip: slide the 7-wide float4 window over its 14 pixels; DO0 / DO1 say whether the row
// contributes to the upper (filter row i0) / lower (filter row i0 - 1) output row.
template <bool DO0, bool DO1>
__device__ __forceinline__ void row_pass(const float4* __restrict__ rowp, const float4* __restrict__ w0,
                                         const float4* __restrict__ w1, float2 (&acc)[2][TW][2]) {
  float4 win[7];
#pragma unroll
  for (int j = 0; j < 6; ++j) win[j] = rowp[j * 16];
#pragma unroll
  for (int x = 0; x < TW; ++x) {
    win[(x + 6) % 7] = rowp[(x + 6) * 16];
    if (DO0) {
      const float4 wa = w0[x * (7 * kWRow / 4)], wb = w0[x * (7 * kWRow / 4) + 1];
      const float wj[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float4 v = win[(x + j) % 7];
        acc[0][x][0] = __ffma2_rn(make_float2(v.x, v.y), make_float2(wj[j], wj[j]), acc[0][x][0]);
        acc[0][x][1] = __ffma2_rn(make_float2(v.z, v.w), make_float2(wj[j], wj[j]), acc[0][x][1]);
      }
    }
    if (DO1) {
      const float4 wa = w1[x * (7 * kWRow / 4)], wb = w1[x * (7 * kWRow / 4) + 1];
      const float wj[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float4 v = win[(x + j) % 7];
        acc[1][x][0] = __ffma2_rn(make_float2(v.x, v.y), make_float2(wj[j], wj[j]), acc[1][x][0]);
        acc[1][x][1] = __ffma2_rn(make_float2(v.z, v.w), make_float2(wj[j], wj[j]), acc[1][x][1]);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2)
adaptive_conv_v3_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmF,
                        float* __restrict__ out, int H, int W, int C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar, filt_bar;
  float* in_s = reinterpret_cast<float*>(smem);                       // [PH][PW][64]
  float* w_s = reinterpret_cast<float*>(smem + kStageBytes);          // [TH*TW][7][8]
  const int b = blockIdx.z, ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const int ncg = C / CG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmIn);
    tc::prefetch_tmap(&tmF);
    tc::mbar_init(&full_bar, 1);
    tc::mbar_init(&filt_bar, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(&full_bar, kStageBytes);
    tc::tma_load_4d(smem, &tmIn, &full_bar, 0, tx0, ty0, b);
    tc::mbar_arrive_expect_tx(&filt_bar, kWBytes);   // OOB pixels of an edge tile arrive as zeros
    tc::tma_load_4d(smem + kStageBytes, &tmF, &filt_bar, 0, tx0, ty0, b);
  }
  __syncthreads();
  tc::mbar_wait(&filt_bar, 0);
  const int hl = lane & 15;                      // float4 index inside the 64-channel group
  const int r0 = 2 * (warp * 2 + (lane >> 4));   // upper output row of this half-warp's pair
  for (int cg = 0; cg < ncg; ++cg) {
    // the tile is single-buffered: pull the next channel group into L2 now, so that the TMA load issued after this
    // group's sweep is an L2 hit (ncu: 26 % of the stall samples were the wait for that load)
    if (threadIdx.x == 0 && cg + 1 < ncg) tc::tma_prefetch_l2_4d(&tmIn, (cg + 1) * CG, tx0, ty0, b);
    tc::mbar_wait(&full_bar, cg & 1);
    const float4* tile = reinterpret_cast<const float4*>(in_s) + hl;
    const float4* wr0 = reinterpret_cast<const float4*>(w_s + (r0 * TW) * 7 * kWRow);
    const float4* wr1 = reinterpret_cast<const float4*>(w_s + ((r0 + 1) * TW) * 7 * kWRow);
    float2 acc[2][TW][2];
#pragma unroll
    for (int x = 0; x < TW; ++x) acc[0][x][0] = acc[0][x][1] = acc[1][x][0] = acc[1][x][1] = make_float2(0.f, 0.f);
    row_pass<true, false>(tile + (r0 + 0) * PW * 16, wr0, wr1, acc);
#pragma unroll 1
    for (int ir = 1; ir < 7; ++ir)
      row_pass<true, true>(tile + (r0 + ir) * PW * 16, wr0 + ir * (kWRow / 4), wr1 + (ir - 1) * (kWRow / 4), acc);
    row_pass<false, true>(tile + (r0 + 7) * PW * 16, wr0, wr1 + 6 * (kWRow / 4), acc);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = ty0 + r0 + rr;
      if (y < H) {
        float* orow = out + (((size_t)b * H + y) * W + tx0) * C + cg * CG + 4 * hl;
#pragma unroll
        for (int x = 0; x < TW; ++x)
          if (tx0 + x < W)
            *reinterpret_cast<float4*>(orow + (size_t)x * C) =
                make_float4(acc[rr][x][0].x, acc[rr][x][0].y, acc[rr][x][1].x, acc[rr][x][1].y);
      }
    }
    __syncthreads();  // tile fully consumed
    if (threadIdx.x == 0 && cg + 1 < ncg) {
      tc::mbar_arrive_expect_tx(&full_bar, kStageBytes);
      tc::tma_load_4d(smem, &tmIn, &full_bar, (cg + 1) * CG, tx0, ty0, b);
    }
  }
}

}  // namespace ac3
}  // namespace isp

using namespace isp;

extern "C" int isp_adaptive_conv_fwd(const float* in_padded, const float* filters, float* out, int B, int H, int W,
                                     int C, int filt_ld, isp_stream_t stream) {
  ISP_REQUIRE(in_padded && filters && out, ISP_ERR_BAD_SHAPE, "adaptive_conv_fwd: null pointer");
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, ISP_ERR_BAD_SHAPE, "adaptive_conv_fwd: bad shape B=%d H=%d W=%d C=%d", B,
              H, W, C);
  ISP_REQUIRE(filt_ld == 49 || filt_ld == 56, ISP_ERR_BAD_SHAPE, "adaptive_conv_fwd: filt_ld must be 49 or 56 (got %d)", filt_ld);
  ISP_REQUIRE(filt_ld == 49 || aligned16(filters), ISP_ERR_MISALIGNED, "adaptive_conv_fwd: padded filters must be 16-byte aligned");
  ISP_REQUIRE(C % ac2::CG == 0, ISP_ERR_UNSUPPORTED, "adaptive_conv_fwd: NHWC path needs C %% 64 == 0 (C=%d)", C);
  ISP_REQUIRE(aligned16(in_padded) && aligned16(out), ISP_ERR_MISALIGNED, "adaptive_conv_fwd: 16-byte alignment");
  ISP_REQUIRE(B <= 65535 && cdiv(H, ac2::TH) <= 65535, ISP_ERR_UNSUPPORTED, "adaptive_conv_fwd: grid too large");
  if (filt_ld == 56) {  // row-padded filters: third-generation kernel (two output rows per half-warp)
    CUtensorMap tmI, tmW;
    const uint64_t Hp = H + 6, Wp = W + 6;
    const uint64_t dims[4] = {(uint64_t)C, Wp, Hp, (uint64_t)B};
    const uint64_t str[4] = {4, (uint64_t)C * 4, Wp * C * 4, Hp * Wp * C * 4};
    const uint32_t box[4] = {ac3::CG, ac3::PW, ac3::PH, 1};
    if (int e = make_tmap(&tmI, 4, in_padded, 4, dims, str, box, "adaptive_conv_fwd(in)", false)) return e;
    const uint64_t fdims[4] = {56, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t fstr[4] = {4, 56 * 4, (uint64_t)W * 56 * 4, (uint64_t)H * W * 56 * 4};
    const uint32_t fbox[4] = {56, ac3::TW, ac3::TH, 1};
    if (int e = make_tmap(&tmW, 4, filters, 4, fdims, fstr, fbox, "adaptive_conv_fwd(filters)", false)) return e;
    if (int e = ensure_dynamic_smem((const void*)ac3::adaptive_conv_v3_kernel, (int)ac3::kSmem)) return e;
    ISP_REQUIRE(cdiv(H, ac3::TH) <= 65535, ISP_ERR_UNSUPPORTED, "adaptive_conv_fwd: grid too large");
    dim3 grid3(cdiv(W, ac3::TW), cdiv(H, ac3::TH), B);
    ac3::adaptive_conv_v3_kernel<<<grid3, ac3::kThreads, ac3::kSmem, as_stream(stream)>>>(tmI, tmW, out, H, W, C);
    ISP_CHECK_LAUNCH("adaptive_conv_v3_kernel");
    return ISP_OK;
  }
  CUtensorMap tm, tmF;
  {
    const uint64_t Hp = H + 6, Wp = W + 6;
    const uint64_t dims[4] = {(uint64_t)C, Wp, Hp, (uint64_t)B};
    const uint64_t str[4] = {4, (uint64_t)C * 4, Wp * C * 4, Hp * Wp * C * 4};
    const uint32_t box[4] = {ac2::CG, ac2::PW, ac2::PH, 1};
    if (int e = make_tmap(&tm, 4, in_padded, 4, dims, str, box, "adaptive_conv_fwd(in)", false)) return e;
  }
  tmF = tm;
  if (filt_ld == 56) {
    const uint64_t dims[4] = {56, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {4, 56 * 4, (uint64_t)W * 56 * 4, (uint64_t)H * W * 56 * 4};
    const uint32_t box[4] = {56, ac2::TW, ac2::TH, 1};
    if (int e = make_tmap(&tmF, 4, filters, 4, dims, str, box, "adaptive_conv_fwd(filters)", false)) return e;
  }
  if (int e = ensure_dynamic_smem((const void*)ac2::adaptive_conv_v2_kernel, (int)ac2::kSmem)) return e;
  dim3 grid(cdiv(W, ac2::TW), cdiv(H, ac2::TH), B);
  ac2::adaptive_conv_v2_kernel<<<grid, ac2::kThreads, ac2::kSmem, as_stream(stream)>>>(tm, tmF, filters, out, H, W, C, filt_ld);
  ISP_CHECK_LAUNCH("adaptive_conv_v2_kernel");
  return ISP_OK;
}
