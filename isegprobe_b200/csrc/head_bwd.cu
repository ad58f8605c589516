// Backward helpers of the IS head (ConvSegHead, core/model/heads/conv_heads.py:48-73; classifier
// core/model/heads/base_head.py:8-18) that are not GEMMs: the 1x1 classifier's backward fused with the
// ReLU mask of the layer below it, and per-channel sums (bias gradients).  HBM-bound: one pass each.
#include "common.cuh"

namespace isp {

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// logits[m] = sum_c a[m,c] wc[c] + bc  (num_classes == 1).  Given dlog[m]:
//   dz[m,c]  = a[m,c] > 0 ? dlog[m] * wc[c] : 0      (gradient wrt the pre-ReLU output of the last 3x3 conv)
//   dwc[c]  += sum_m dlog[m] * a[m,c];   dbc += sum_m dlog[m];   dbz[c] += sum_m dz[m,c]  (that conv's bias grad)
// Warp per row (grid-stride), lane owns V groups of 8 channels; column sums live in registers, are
// combined across the block's warps in smem and leave as one atomicAdd per channel and block.
template <int V>
__global__ void __launch_bounds__(256) classifier_bwd_kernel(const __nv_bfloat16* __restrict__ a, long long lda,
                                                             const float* __restrict__ dlog, const float* __restrict__ wc,
                                                             __nv_bfloat16* __restrict__ dz, long long ldz,
                                                             float* __restrict__ dwc, float* __restrict__ dbc,
                                                             float* __restrict__ dbz, long long M, int C) {
  __shared__ float red[8][V * 256 + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * 8;
  float w[V][8], sw[V][8], sz[V][8], sb = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (lane + 32 * i) * 8 + e;
      w[i][e] = c < C ? __ldg(wc + c) : 0.f;
      sw[i][e] = sz[i][e] = 0.f;
    }
  for (long long m = (long long)blockIdx.x * 8 + wid; m < M; m += nwarps) {
    const float g = __ldg(dlog + m);
    sb += g;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c0 = (lane + 32 * i) * 8;
      if (c0 >= C) continue;
      float x[8], y[8];
      unpack8(*reinterpret_cast<const uint4*>(a + m * lda + c0), x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        y[e] = x[e] > 0.f ? g * w[i][e] : 0.f;
        sw[i][e] = fmaf(g, x[e], sw[i][e]);
        sz[i][e] += y[e];
      }
      *reinterpret_cast<uint4*>(dz + m * ldz + c0) = pack8(y);
    }
  }
  // two rounds through one smem buffer: classifier weight gradient (+ its bias), then the conv bias gradient
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[wid][(lane + 32 * i) * 8 + e] = sw[i][e];
  if (lane == 0) red[wid][V * 256] = sb;
  __syncthreads();
  for (int idx = threadIdx.x; idx < V * 256 + 1; idx += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][idx];
    if (idx == V * 256) atomicAdd(dbc, s);
    else if (idx < C) atomicAdd(dwc + idx, s);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[wid][(lane + 32 * i) * 8 + e] = sz[i][e];
  __syncthreads();
  for (int idx = threadIdx.x; idx < V * 256 && idx < C; idx += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][idx];
    atomicAdd(dbz + idx, s);
  }
}


// out[c] += sum_m x[m,c]   (bf16 rows): bias gradient of a conv layer from its output gradient
template <int V>
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                          float* __restrict__ out, long long M, int C) {
  __shared__ float red[8][V * 256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * 8;
  float s[V][8];
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) s[i][e] = 0.f;
  for (long long m = (long long)blockIdx.x * 8 + wid; m < M; m += nwarps) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c0 = (lane + 32 * i) * 8;
      if (c0 >= C) continue;
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(x + m * ld + c0), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[i][e] += v[e];
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[wid][(lane + 32 * i) * 8 + e] = s[i][e];
  __syncthreads();
  for (int c = threadIdx.x; c < V * 256 && c < C; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][c];
    atomicAdd(out + c, t);
  }
}

}  // namespace isp

using namespace isp;

extern "C" int isp_head_classifier_bwd(const void* act_bf16, long long lda, const float* dlogits, const float* wc,
                                       void* dz_bf16, long long ldz, float* dwc, float* dbc, float* dbz, long long M,
                                       int C, isp_stream_t stream) {
  ISP_REQUIRE(act_bf16 && dlogits && wc && dz_bf16 && dwc && dbc && dbz, ISP_ERR_BAD_SHAPE, "head_classifier_bwd: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 1024 && lda >= C && ldz >= C, ISP_ERR_BAD_SHAPE,
              "head_classifier_bwd: bad shape (C must be a multiple of 8, <= 1024)");
  ISP_REQUIRE(lda % 8 == 0 && ldz % 8 == 0 && aligned16(act_bf16) && aligned16(dz_bf16), ISP_ERR_MISALIGNED,
              "head_classifier_bwd: rows must be 16-byte aligned");
  const int V = (C / 8 + 31) / 32;
  const int grid = (int)(M / 8 + 1 < 148 * 4 ? M / 8 + 1 : 148 * 4);
  auto a = reinterpret_cast<const __nv_bfloat16*>(act_bf16);
  auto z = reinterpret_cast<__nv_bfloat16*>(dz_bf16);
  if (V == 1) classifier_bwd_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(a, lda, dlogits, wc, z, ldz, dwc, dbc, dbz, M, C);
  else if (V == 2) classifier_bwd_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(a, lda, dlogits, wc, z, ldz, dwc, dbc, dbz, M, C);
  else if (V == 3) classifier_bwd_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(a, lda, dlogits, wc, z, ldz, dwc, dbc, dbz, M, C);
  else classifier_bwd_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(a, lda, dlogits, wc, z, ldz, dwc, dbc, dbz, M, C);
  ISP_CHECK_LAUNCH("classifier_bwd_kernel");
  return ISP_OK;
}

extern "C" int isp_colsum_bf16(const void* x_bf16, long long ld, float* out, long long M, int C, isp_stream_t stream) {
  ISP_REQUIRE(x_bf16 && out && M > 0 && C > 0 && C % 8 == 0 && C <= 1024 && ld >= C, ISP_ERR_BAD_SHAPE,
              "colsum_bf16: bad shape (C must be a multiple of 8, <= 1024)");
  ISP_REQUIRE(ld % 8 == 0 && aligned16(x_bf16), ISP_ERR_MISALIGNED, "colsum_bf16: rows must be 16-byte aligned");
  const int V = (C / 8 + 31) / 32;
  const int grid = (int)(M / 8 + 1 < 148 * 4 ? M / 8 + 1 : 148 * 4);
  auto x = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  if (V == 1) colsum_bf16_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, ld, out, M, C);
  else if (V == 2) colsum_bf16_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(x, ld, out, M, C);
  else if (V == 3) colsum_bf16_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(x, ld, out, M, C);
  else colsum_bf16_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, ld, out, M, C);
  ISP_CHECK_LAUNCH("colsum_bf16_kernel");
  return ISP_OK;
}
