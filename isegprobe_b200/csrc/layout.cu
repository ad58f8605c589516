// Layout / dtype movers and the align_corners=True bilinear resize (a14/a16 of
// core/model/iseg_probe_model.py:120-129, iseg_base_model.py:75-80).
#include "common.cuh"

namespace isp {

// [B,C,H,W] (element strides sb,sc,sh,sw) -> dense [B,H*W,Cpad]; 32x32 smem transpose
template <typename OutT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, OutT* __restrict__ out, int C,
                                                           int H, int W, int Cpad, long long sb, long long sc,
                                                           long long sh, long long sw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int HW = H * W;
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) v = in[b * sb + c * sc + (p / W) * sh + (p % W) * sw];
    tile[k][tx] = v;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    if (p < HW && c < Cpad) {
      const float v = tile[tx][k];
      if constexpr (sizeof(OutT) == 2) out[((size_t)b * HW + p) * Cpad + c] = __float2bfloat16(v);
      else out[((size_t)b * HW + p) * Cpad + c] = v;
    }
  }
}

// bilinear, align_corners=True (ATen upsample_bilinear2d): NHWC f32 -> NHWC f32|bf16
// DUAL: additionally writes a bf16 copy (the tensor-core operand of the next GEMM) in the same pass.
template <typename OutT, bool DUAL = false>
__global__ void __launch_bounds__(256) bilinear_ac_kernel(const float* __restrict__ in, OutT* __restrict__ out, int B,
                                                          int C, int Hin, int Win, int Hout, int Wout, int Cpad,
                                                          float sy, float sx, __nv_bfloat16* __restrict__ out2 = nullptr,
                                                          const float* __restrict__ bias = nullptr) {
  const int C4 = Cpad / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Hout * Wout * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const long long p = idx / C4;
  const int ox = (int)(p % Wout), oy = (int)((p / Wout) % Hout), b = (int)(p / ((long long)Wout * Hout));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < C) {
    const float fy = sy * (float)oy, fx = sx * (float)ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float4* s = reinterpret_cast<const float4*>(in) + (size_t)b * Hin * Win * (C / 4) + c4;
    const int Ci4 = C / 4;
    const float4 v00 = __ldg(s + ((size_t)y0 * Win + x0) * Ci4), v01 = __ldg(s + ((size_t)y0 * Win + x1) * Ci4);
    const float4 v10 = __ldg(s + ((size_t)y1 * Win + x0) * Ci4), v11 = __ldg(s + ((size_t)y1 * Win + x1) * Ci4);
    r.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
    r.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
    r.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
    r.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
    if (bias) {  // per-channel constant added after the interpolation
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
    }
  }
  if constexpr (sizeof(OutT) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r.x, r.y), c = __floats2bfloat162_rn(r.z, r.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&a);
    u.y = *reinterpret_cast<unsigned*>(&c);
    reinterpret_cast<uint2*>(out)[idx] = u;
  } else {
    reinterpret_cast<float4*>(out)[idx] = r;
  }
  if constexpr (DUAL) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r.x, r.y), c = __floats2bfloat162_rn(r.z, r.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&a);
    u.y = *reinterpret_cast<unsigned*>(&c);
    reinterpret_cast<uint2*>(out2)[idx] = u;
  }
}

// Same resize for the JBU stack's last pass (fp32 -> fp32 + per-channel bias), organised for the memory system: a block is
// 32 lanes (4 channels each = one 512-byte run of a pixel) x 8 adjacent output columns and marches down a strip of RS output
// rows.  The horizontally interpolated value of a source row is kept in registers and reused when the next output row's
// upper source row is this one's lower row (always, for up-scaling; 7 rows of 8 for 512 -> 448), and the two source
// columns of neighbouring output columns are fetched by the block at the same time (L1 hits): ~1.5 source reads per
// output from L2 instead of 4.  Same expression tree as bilinear_ac_kernel.
constexpr int RS = 16;
template <typename OutT>
__global__ void __launch_bounds__(256) bilinear_ac_march_kernel(const float* __restrict__ in, OutT* __restrict__ out,
                                                                const float* __restrict__ bias, int C, int Hin, int Win,
                                                                int Hout, int Wout, float sy, float sx) {
  const int C4 = C / 4;
  const int c4 = blockIdx.z % (C4 / 32) * 32 + threadIdx.x, b = blockIdx.z / (C4 / 32);
  const int ox = blockIdx.x * 8 + threadIdx.y;
  if (ox >= Wout) return;
  const float fx = sx * (float)ox;
  const int x0 = (int)fx, x1 = x0 + (x0 < Win - 1 ? 1 : 0);
  const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
  const float4* s = reinterpret_cast<const float4*>(in) + (size_t)b * Hin * Win * C4 + c4;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  auto hrow = [&](int y) {
    const float4 a = __ldg(s + ((size_t)y * Win + x0) * C4), c = __ldg(s + ((size_t)y * Win + x1) * C4);
    return make_float4(lx0 * a.x + lx1 * c.x, lx0 * a.y + lx1 * c.y, lx0 * a.z + lx1 * c.z, lx0 * a.w + lx1 * c.w);
  };
  int have = -1;
  float4 hc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int oy_end = min(Hout, (int)(blockIdx.y + 1) * RS);
  const size_t o0 = ((size_t)b * Hout * Wout + ox) * C4 + c4;  // in 4-channel units
#pragma unroll 2
  for (int oy = blockIdx.y * RS; oy < oy_end; ++oy) {
    const float fy = sy * (float)oy;
    const int y0 = (int)fy, y1 = y0 + (y0 < Hin - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    const float4 h0 = (y0 == have) ? hc : hrow(y0);
    const float4 h1 = (y1 == y0) ? h0 : hrow(y1);
    have = y1;
    hc = h1;
    float4 r;
    r.x = ly0 * h0.x + ly1 * h1.x + bb.x;
    r.y = ly0 * h0.y + ly1 * h1.y + bb.y;
    r.z = ly0 * h0.z + ly1 * h1.z + bb.z;
    r.w = ly0 * h0.w + ly1 * h1.w + bb.w;
    const size_t oi = o0 + (size_t)oy * Wout * C4;
    if constexpr (sizeof(OutT) == 2) {  // bf16 features for a head that rounds its input to bf16 anyway
      __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
      __stcs(reinterpret_cast<uint2*>(out) + oi, make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi)));
    } else {
      __stcs(reinterpret_cast<float4*>(out) + oi, r);
    }
  }
}

// Adjoint of bilinear_ac_kernel (gradient w.r.t. its input): every input pixel gathers, with the forward's own
// fp32 coordinates and weights, from the output pixels it contributed to.  VEC = 4: float4 over channels.
template <int VEC>
__global__ void __launch_bounds__(256) bilinear_ac_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int B,
                                                              int C, int Hin, int Win, int Hout, int Wout, float sy,
                                                              float sx) {
  const int CV = C / VEC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Hin * Win * CV;
  if (idx >= total) return;
  const int cv = (int)(idx % CV);
  const long long p = idx / CV;
  const int xi = (int)(p % Win), yi = (int)((p / Win) % Hin), b = (int)(p / ((long long)Win * Hin));
  // output rows / columns whose source coordinate lies within one pixel of (yi, xi)
  const int oy0 = sy > 0.f ? max(0, (int)floorf((float)(yi - 1) / sy)) : 0;
  const int oy1 = sy > 0.f ? min(Hout - 1, (int)ceilf((float)(yi + 1) / sy)) : Hout - 1;
  const int ox0 = sx > 0.f ? max(0, (int)floorf((float)(xi - 1) / sx)) : 0;
  const int ox1 = sx > 0.f ? min(Wout - 1, (int)ceilf((float)(xi + 1) / sx)) : Wout - 1;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  const float* g = gout + (size_t)b * Hout * Wout * C + (size_t)cv * VEC;
  for (int oy = oy0; oy <= oy1; ++oy) {
    const float fy = sy * (float)oy;
    const int y0 = (int)fy, y1 = y0 + (y0 < Hin - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0;
    const float wy = (y0 == yi ? 1.f - ly1 : 0.f) + (y1 == yi ? ly1 : 0.f);
    if (wy == 0.f) continue;
    for (int ox = ox0; ox <= ox1; ++ox) {
      const float fx = sx * (float)ox;
      const int x0 = (int)fx, x1 = x0 + (x0 < Win - 1 ? 1 : 0);
      const float lx1 = fx - (float)x0;
      const float wx = (x0 == xi ? 1.f - lx1 : 0.f) + (x1 == xi ? lx1 : 0.f);
      if (wx == 0.f) continue;
      const float wgt = wy * wx;
      const float* gp = g + ((size_t)oy * Wout + ox) * C;
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gp));
        acc[0] = fmaf(wgt, t.x, acc[0]); acc[1] = fmaf(wgt, t.y, acc[1]);
        acc[2 % VEC] = fmaf(wgt, t.z, acc[2 % VEC]); acc[3 % VEC] = fmaf(wgt, t.w, acc[3 % VEC]);
      } else {
        acc[0] = fmaf(wgt, __ldg(gp), acc[0]);
      }
    }
  }
  float* o = gin + ((size_t)(b * Hin + yi) * Win + xi) * C + (size_t)cv * VEC;
#pragma unroll
  for (int v = 0; v < VEC; ++v) o[v] = acc[v];
}

// C[M,N] = alpha * (A[M,K] W[N,K]^T + bias) + resid   -- fp32 SIMT, 64x64x16 tiles, 4x4 per thread
__global__ void __launch_bounds__(256) gemm_f32_simt_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ resid, float alpha,
                                                            float* __restrict__ Cm, long long M, int N, int K) {
  __shared__ float As[16][65], Ws[16][65];
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      As[k][r] = (m0 + r < M && k0 + k < K) ? A[(m0 + r) * K + k0 + k] : 0.f;
      Ws[k][r] = (n0 + r < N && k0 + k < K) ? Wt[(size_t)(n0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      v *= alpha;
      if (resid) v += resid[m * N + n];
      Cm[m * N + n] = v;
    }
  }
}

}  // namespace isp

using namespace isp;

extern "C" int isp_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W, long long sb, long long sc,
                                    long long sh, long long sw, isp_stream_t stream) {
  ISP_REQUIRE(in && out && B > 0 && C > 0 && H > 0 && W > 0, ISP_ERR_BAD_SHAPE, "nchw_to_nhwc_f32: bad arguments");
  ISP_REQUIRE(B <= 65535 && cdiv(C, 32) <= 65535, ISP_ERR_UNSUPPORTED, "nchw_to_nhwc_f32: grid too large");
  dim3 grid(cdiv((long long)H * W, 32), cdiv(C, 32), B);
  nchw_to_nhwc_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(in, out, C, H, W, C, sb, sc, sh, sw);
  ISP_CHECK_LAUNCH("nchw_to_nhwc_kernel<float>");
  return ISP_OK;
}

extern "C" int isp_nchw_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, int Cpad, long long sb,
                                     long long sc, long long sh, long long sw, isp_stream_t stream) {
  ISP_REQUIRE(in && out && B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, ISP_ERR_BAD_SHAPE,
              "nchw_to_nhwc_bf16: bad arguments");
  ISP_REQUIRE(B <= 65535 && cdiv(Cpad, 32) <= 65535, ISP_ERR_UNSUPPORTED, "nchw_to_nhwc_bf16: grid too large");
  dim3 grid(cdiv((long long)H * W, 32), cdiv(Cpad, 32), B);
  nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(in, reinterpret_cast<__nv_bfloat16*>(out), C,
                                                                         H, W, Cpad, sb, sc, sh, sw);
  ISP_CHECK_LAUNCH("nchw_to_nhwc_kernel<bf16>");
  return ISP_OK;
}

extern "C" int isp_bilinear_ac_nhwc(const float* in, void* out, int B, int C, int Hin, int Win, int Hout, int Wout,
                                    int out_bf16, int Cpad, isp_stream_t stream) {
  ISP_REQUIRE(in && out && B > 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, ISP_ERR_BAD_SHAPE,
              "bilinear_ac_nhwc: bad arguments");
  ISP_REQUIRE(C % 4 == 0 && Cpad % 4 == 0 && Cpad >= C, ISP_ERR_UNSUPPORTED, "bilinear_ac_nhwc: C, Cpad %% 4 == 0");
  ISP_REQUIRE(aligned16(in) && aligned16(out), ISP_ERR_MISALIGNED, "bilinear_ac_nhwc: 16-byte alignment");
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  const long long total = (long long)B * Hout * Wout * (Cpad / 4);
  if (out_bf16)
    bilinear_ac_kernel<__nv_bfloat16><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(
        in, reinterpret_cast<__nv_bfloat16*>(out), B, C, Hin, Win, Hout, Wout, Cpad, sy, sx);
  else
    bilinear_ac_kernel<float><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<float*>(out), B, C,
                                                                               Hin, Win, Hout, Wout, Cpad, sy, sx);
  ISP_CHECK_LAUNCH("bilinear_ac_kernel");
  return ISP_OK;
}

// Gradient of isp_bilinear_ac_nhwc w.r.t. its input: gin [B,Hin,Win,C] = resize^T gout [B,Hout,Wout,C] (fp32 NHWC).
extern "C" int isp_bilinear_ac_nhwc_bwd(const float* gout, float* gin, int B, int C, int Hin, int Win, int Hout, int Wout,
                                        isp_stream_t stream) {
  ISP_REQUIRE(gout && gin && B > 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, ISP_ERR_BAD_SHAPE,
              "bilinear_ac_nhwc_bwd: bad arguments");
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  if (C % 4 == 0 && aligned16(gout)) {
    const long long total = (long long)B * Hin * Win * (C / 4);
    bilinear_ac_bwd_kernel<4><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(gout, gin, B, C, Hin, Win, Hout, Wout, sy, sx);
  } else {
    const long long total = (long long)B * Hin * Win * C;
    bilinear_ac_bwd_kernel<1><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(gout, gin, B, C, Hin, Win, Hout, Wout, sy, sx);
  }
  ISP_CHECK_LAUNCH("bilinear_ac_bwd_kernel");
  return ISP_OK;
}

extern "C" int isp_bilinear_ac_nhwc_dual(const float* in, float* out_f32, void* out_bf16, int B, int C, int Hin, int Win,
                                         int Hout, int Wout, isp_stream_t stream) {
  ISP_REQUIRE(in && out_f32 && out_bf16 && B > 0 && C > 0 && C % 4 == 0, ISP_ERR_BAD_SHAPE,
              "bilinear_ac_nhwc_dual: bad arguments (C must be a multiple of 4)");
  ISP_REQUIRE(Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, ISP_ERR_BAD_SHAPE, "bilinear_ac_nhwc_dual: bad size");
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  const long long total = (long long)B * Hout * Wout * (C / 4);
  bilinear_ac_kernel<float, true><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(
      in, out_f32, B, C, Hin, Win, Hout, Wout, C, sy, sx, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  ISP_CHECK_LAUNCH("bilinear_ac_kernel(dual)");
  return ISP_OK;
}

// out[b,y,x,c] = resize(in)[b,y,x,c] + bias[c]  (fp32 NHWC; bias may be NULL): the last pass of the JBU stack, whose
// final 1x1 conv has been commuted to the source -- only its bias is left to add here.
extern "C" int isp_bilinear_ac_nhwc_bias(const float* in, void* out, int out_bf16, const float* bias, int B, int C, int Hin,
                                         int Win, int Hout, int Wout, isp_stream_t stream) {
  ISP_REQUIRE(in && out && B > 0 && C > 0 && C % 4 == 0, ISP_ERR_BAD_SHAPE,
              "bilinear_ac_nhwc_bias: bad arguments (C must be a multiple of 4)");
  ISP_REQUIRE(Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, ISP_ERR_BAD_SHAPE, "bilinear_ac_nhwc_bias: bad size");
  ISP_REQUIRE(aligned16(in) && aligned16(out) && aligned16(bias), ISP_ERR_MISALIGNED, "bilinear_ac_nhwc_bias: 16-byte alignment");
  const float sy = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f;
  const float sx = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
  if (C % 128 == 0 && (long long)B * (C / 128) <= 65535 && cdiv(Hout, RS) <= 65535) {
    dim3 grid(cdiv(Wout, 8), cdiv(Hout, RS), B * (C / 128));
    if (out_bf16)
      bilinear_ac_march_kernel<__nv_bfloat16><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(
          in, reinterpret_cast<__nv_bfloat16*>(out), bias, C, Hin, Win, Hout, Wout, sy, sx);
    else
      bilinear_ac_march_kernel<float><<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, reinterpret_cast<float*>(out), bias, C,
                                                                                   Hin, Win, Hout, Wout, sy, sx);
    ISP_CHECK_LAUNCH("bilinear_ac_march_kernel");
    return ISP_OK;
  }
  const long long total = (long long)B * Hout * Wout * (C / 4);
  if (out_bf16)
    bilinear_ac_kernel<__nv_bfloat16><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(
        in, reinterpret_cast<__nv_bfloat16*>(out), B, C, Hin, Win, Hout, Wout, C, sy, sx, nullptr, bias);
  else
    bilinear_ac_kernel<float><<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(in, reinterpret_cast<float*>(out), B, C, Hin, Win,
                                                                               Hout, Wout, C, sy, sx, nullptr, bias);
  ISP_CHECK_LAUNCH("bilinear_ac_kernel(bias)");
  return ISP_OK;
}

extern "C" int isp_gemm_f32_simt(const float* A, const float* W, const float* bias, const float* resid, float alpha,
                                 float* C, long long M, int N, int K, isp_stream_t stream) {
  ISP_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, ISP_ERR_BAD_SHAPE, "gemm_f32_simt: bad arguments");
  dim3 grid(cdiv(M, 64), cdiv(N, 64));
  ISP_REQUIRE(grid.y <= 65535, ISP_ERR_UNSUPPORTED, "gemm_f32_simt: N too large");
  gemm_f32_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, W, bias, resid, alpha, C, M, N, K);
  ISP_CHECK_LAUNCH("gemm_f32_simt_kernel");
  return ISP_OK;
}
