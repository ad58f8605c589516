// BatchNorm2d in TRAINING mode for the frozen LoftUp / LiFT conv stacks.  The reference's trainer calls net.train()
// on the whole model (core/training/trainer.py:213-214), which flips the frozen upsampler's BatchNorm layers
// (loftup/loftup.py:55-65, LiFT.py:17-24) to batch statistics + running-statistic updates.  With batch
// statistics the normalisation cannot be folded into the conv epilogue (the statistics are those of the conv output
// itself), so the conv writes its raw output and two bandwidth kernels follow:
//   isp_col_moments_bf16    per-channel (sum, sum of squares) partials over slabs of 1024 pixels (deterministic: no atomics;
//                           the host adds the slab partials in fp64)
//   isp_bn_relu_rows_bf16   y = max(x * scale[c] + shift[c], 0) in place (+ the per-pixel (sum, sum of squares) of the stored
//                           values, the row statistics the next GEMM's fused LayerNorm expects)
#include "common.cuh"

namespace isp {

constexpr int kSlabRows = ISP_COL_MOMENTS_SLAB_ROWS;

// grid.x = slabs; thread t owns the bf16 column pair (2t, 2t+1); rows of the slab are walked with fully coalesced loads
__global__ void __launch_bounds__(256) col_moments_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long M,
                                                          int C, float* __restrict__ partial, int ldp) {
  const long long r0 = (long long)blockIdx.x * kSlabRows;
  const long long r1 = r0 + kSlabRows < M ? r0 + kSlabRows : M;
  for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const __nv_bfloat16* px = x + r0 * ld + c;
    if (c + 1 < C) {
#pragma unroll 8
      for (long long r = r0; r < r1; ++r, px += ld) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px));
        s0 += v.x; q0 = fmaf(v.x, v.x, q0);
        s1 += v.y; q1 = fmaf(v.y, v.y, q1);
      }
    } else {
      for (long long r = r0; r < r1; ++r, px += ld) {
        const float v = __bfloat162float(*px);
        s0 += v; q0 = fmaf(v, v, q0);
      }
    }
    float* out = partial + (long long)blockIdx.x * 2 * ldp;
    out[c] = s0; out[ldp + c] = q0;
    if (c + 1 < C) { out[c + 1] = s1; out[ldp + c + 1] = q1; }
  }
}

// one warp per row; lanes stride over bf16 pairs
__global__ void __launch_bounds__(256) bn_relu_rows_kernel(__nv_bfloat16* __restrict__ x, long long ld, long long M, int C,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           float2* __restrict__ stats) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* px = x + row * ld;
  float s = 0.f, q = 0.f;
  for (int c = 2 * lane; c < C; c += 64) {
    if (c + 1 < C) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(px + c));
      const float a = fmaxf(fmaf(v.x, __ldg(scale + c), __ldg(shift + c)), 0.f);
      const float b = fmaxf(fmaf(v.y, __ldg(scale + c + 1), __ldg(shift + c + 1)), 0.f);
      const __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
      *reinterpret_cast<__nv_bfloat162*>(px + c) = o;
      const float2 w = __bfloat1622float2(o);  // statistics of the values as stored
      s += w.x + w.y; q = fmaf(w.x, w.x, fmaf(w.y, w.y, q));
    } else {
      const float a = fmaxf(fmaf(__bfloat162float(px[c]), __ldg(scale + c), __ldg(shift + c)), 0.f);
      const __nv_bfloat16 o = __float2bfloat16_rn(a);
      px[c] = o;
      const float w = __bfloat162float(o);
      s += w; q = fmaf(w, w, q);
    }
  }
  if (stats) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      q += __shfl_xor_sync(0xffffffffu, q, d);
    }
    if (lane == 0) stats[row] = make_float2(s, q);
  }
}

}  // namespace isp

using namespace isp;

// partial: fp32 [ceil(M / ISP_COL_MOMENTS_SLAB_ROWS)][2][ldp]; row 0 of a slab = column sums, row 1 = sums of squares
extern "C" int isp_col_moments_bf16(const void* x, long long ld, long long M, int C, float* partial, int ldp,
                                    isp_stream_t stream) {
  ISP_REQUIRE(x && partial, ISP_ERR_BAD_SHAPE, "col_moments_bf16: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && ld >= C && ldp >= C, ISP_ERR_BAD_SHAPE, "col_moments_bf16: bad shape");
  ISP_REQUIRE(ld % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3u) == 0, ISP_ERR_MISALIGNED,
              "col_moments_bf16: x must be 4-byte aligned with an even row stride");
  const long long slabs = (M + kSlabRows - 1) / kSlabRows;
  ISP_REQUIRE(slabs < (1ll << 31), ISP_ERR_UNSUPPORTED, "col_moments_bf16: too many rows");
  col_moments_kernel<<<(unsigned)slabs, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, M, C, partial,
                                                                   ldp);
  ISP_CHECK_LAUNCH("col_moments_kernel");
  return ISP_OK;
}

// stats: optional fp32 [M][2] (sum, sum of squares of the stored row) -- one slot per row for isp_gemm_bf16_tc_ex(ln_slots=1)
extern "C" int isp_bn_relu_rows_bf16(void* x, long long ld, const float* scale, const float* shift, long long M, int C,
                                     float* stats, isp_stream_t stream) {
  ISP_REQUIRE(x && scale && shift, ISP_ERR_BAD_SHAPE, "bn_relu_rows_bf16: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && ld >= C, ISP_ERR_BAD_SHAPE, "bn_relu_rows_bf16: bad shape");
  ISP_REQUIRE(ld % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3u) == 0, ISP_ERR_MISALIGNED,
              "bn_relu_rows_bf16: x must be 4-byte aligned with an even row stride");
  const long long blocks = (M + 7) / 8;
  ISP_REQUIRE(blocks < (1ll << 31), ISP_ERR_UNSUPPORTED, "bn_relu_rows_bf16: too many rows");
  bn_relu_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(x), ld, M, C, scale, shift,
                                                                     reinterpret_cast<float2*>(stats));
  ISP_CHECK_LAUNCH("bn_relu_rows_kernel");
  return ISP_OK;
}
