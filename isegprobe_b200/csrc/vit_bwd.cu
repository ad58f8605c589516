// Row / element kernels of the activation backward through the frozen ViT (the reference trains the click embedding
// through the frozen backbone: core/model/featurizers/DINOv2.py:518-523, core/model/iseg_probe_model.py:93-99,
// trainer backward core/training/trainer.py:213-221).  The GEMMs of that backward run on gemm_tc_kernel (transposed
// packed weights, isp_gemm_bf16_tc_batched for the per-head attention products); what is left is bandwidth work:
//   layernorm_bwd_kernel   dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ residual gradient)
//   gelu_bwd_kernel        dpre = dh * (Phi(pre) + pre * phi(pre))          (nn.GELU, exact erf form)
//   softmax_rows_kernel    P = softmax(S) over the valid keys of a score row (recomputed from Q K^T)
//   attn_ds_kernel         dS = P * (dP - sum_j P dP)                       (softmax backward, row-wise)
//   transpose_kernel       [Z][R][C] -> [Z][C][R] bf16 (P^T, dS^T and d_emb^T operands)
#include "common.cuh"

#include <cuda_bf16.h>

namespace isp {
namespace vb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per row; C <= 32 * kMaxPerLane
constexpr int kMaxPerLane = 32;

template <bool X_BF16>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, long long lddy,
                                                            const void* __restrict__ xv_, long long ldx,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ resid, long long ldr,
                                                            float* __restrict__ dx, long long lddx,
                                                            __nv_bfloat16* __restrict__ dx_bf, long long ldb, long long M,
                                                            int C, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = reinterpret_cast<const float*>(xv_) + row * ldx;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv_) + row * ldx;
  const float* dr = dy + row * lddy;
  float xv[kMaxPerLane], gv[kMaxPerLane];
  float s = 0.f;
  const int n = (C + 31) / 32;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i < n) {
      const int c = lane + 32 * i;
      xv[i] = c < C ? (X_BF16 ? __bfloat162float(xb[c]) : xr[c]) : 0.f;
      gv[i] = c < C ? dr[c] * __ldg(gamma + c) : 0.f;
      s += xv[i];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < n) {
      const float d = (lane + 32 * i < C) ? xv[i] - mean : 0.f;
      var += d * d;
    }
  const float rstd = rsqrtf(warp_sum(var) / (float)C + eps);
  float a = 0.f, b = 0.f;  // sum(g*dy), sum(g*dy*xhat)
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < n) {
      const float xh = (xv[i] - mean) * rstd;
      xv[i] = xh;
      a += gv[i];
      b += (lane + 32 * i < C) ? gv[i] * xh : 0.f;
    }
  a = warp_sum(a) / (float)C;
  b = warp_sum(b) / (float)C;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < n) {
      const int c = lane + 32 * i;
      if (c < C) {
        float v = rstd * (gv[i] - a - xv[i] * b);
        if (resid) v += resid[row * ldr + c];
        dx[row * lddx + c] = v;
        if (dx_bf) dx_bf[row * ldb + c] = __float2bfloat16(v);
      }
    }
}

// Same, C % 4 == 0 and 16-byte aligned rows: a lane moves float4 (8 bytes of bf16 x) per step -- the scalar kernel ran
// at 1 TB/s on the [802816, 404] query stream of the LoftUp backward.
template <bool X_BF16, int kIt>  // kIt float4 per lane: C <= 128 * kIt
__global__ void __launch_bounds__(256) layernorm_bwd_vec_kernel(const float* __restrict__ dy, long long lddy,
                                                                const void* __restrict__ xv_, long long ldx,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ resid, long long ldr,
                                                                float* __restrict__ dx, long long lddx,
                                                                __nv_bfloat16* __restrict__ dx_bf, long long ldb,
                                                                long long M, int C, float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int C4 = C >> 2;
  const float4* d4 = reinterpret_cast<const float4*>(dy + row * lddy);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 xv[kIt], gv[kIt], rv[kIt];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kIt; ++i) {
    const int c = lane + 32 * i;
    xv[i] = gv[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C4) {
      // the residual gradient is only needed at the end, but its load goes out with the others: one round trip to HBM
      // per row instead of two (the kernel is latency-bound otherwise: one warp per row, four dependent warp reductions)
      if (resid) rv[i] = *reinterpret_cast<const float4*>(resid + row * ldr + 4 * c);
      if (X_BF16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(xv_) + row * ldx + 4 * c);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
        xv[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        xv[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv_) + row * ldx + 4 * c);
      }
      const float4 d = d4[c], g = __ldg(g4 + c);
      gv[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < kIt; ++i)
    if (lane + 32 * i < C4) {
      const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(var) / (float)C + eps);
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < kIt; ++i)
    if (lane + 32 * i < C4) {
      xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd, (xv[i].w - mean) * rstd);
      sa += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      sb += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
  sa = warp_sum(sa) / (float)C;
  sb = warp_sum(sb) / (float)C;
#pragma unroll
  for (int i = 0; i < kIt; ++i) {
    const int c = lane + 32 * i;
    if (c < C4) {
      float4 v = make_float4(rstd * (gv[i].x - sa - xv[i].x * sb), rstd * (gv[i].y - sa - xv[i].y * sb),
                             rstd * (gv[i].z - sa - xv[i].z * sb), rstd * (gv[i].w - sa - xv[i].w * sb));
      v.x += rv[i].x; v.y += rv[i].y; v.z += rv[i].z; v.w += rv[i].w;
      *reinterpret_cast<float4*>(dx + row * lddx + 4 * c) = v;
      if (dx_bf) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<unsigned*>(&lo);
        u.y = *reinterpret_cast<unsigned*>(&hi);
        *reinterpret_cast<uint2*>(dx_bf + row * ldb + 4 * c) = u;
      }
    }
  }
}

// Affine-parameter gradients of LayerNorm (trainable LayerNorms of the simple_vit click embedding):
// dgamma[c] += sum_rows dy[r,c] * xhat[r,c], dbeta[c] += sum_rows dy[r,c].  A block walks a chunk of rows (one warp per
// row at a time, lane = columns lane, lane+32, ...), keeps per-column partial sums in registers, combines its 8 warps
// through shared memory and issues one atomicAdd per column.
template <bool X_BF16>
__global__ void __launch_bounds__(256) layernorm_affine_bwd_kernel(const float* __restrict__ dy, long long lddy,
                                                                   const void* __restrict__ xv_, long long ldx,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   long long M, int C, float eps, int rows_per_block) {
  __shared__ float red[2][8][32 * 8];  // [gamma|beta][warp][column slot] for C <= 1024 handled in 4 passes of 256 columns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  const int n = (C + 31) / 32;
  float ga[kMaxPerLane], gb[kMaxPerLane];
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) ga[i] = gb[i] = 0.f;
  for (long long row = r0 + warp; row < r1; row += 8) {
    const float* xr = reinterpret_cast<const float*>(xv_) + row * ldx;
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv_) + row * ldx;
    const float* dr = dy + row * lddy;
    float xv[kMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i)
      if (i < n) {
        const int c = lane + 32 * i;
        xv[i] = c < C ? (X_BF16 ? __bfloat162float(xb[c]) : xr[c]) : 0.f;
        s += xv[i];
      }
    const float mean = warp_sum(s) / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i)
      if (i < n) {
        const float d = (lane + 32 * i < C) ? xv[i] - mean : 0.f;
        var += d * d;
      }
    const float rstd = rsqrtf(warp_sum(var) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i)
      if (i < n) {
        const int c = lane + 32 * i;
        if (c < C) {
          const float d = dr[c];
          ga[i] = fmaf(d, (xv[i] - mean) * rstd, ga[i]);
          gb[i] += d;
        }
      }
  }
  for (int base = 0; base < n; base += 8) {  // 8 column slots (256 columns) per pass through shared memory
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[0][warp][i * 32 + lane] = (base + i < n) ? ga[base + i] : 0.f;
      red[1][warp][i * 32 + lane] = (base + i < n) ? gb[base + i] : 0.f;
    }
    __syncthreads();
    {
      const int slot = threadIdx.x;  // 256 threads = 256 column slots
      const int c = (base + slot / 32) * 32 + (slot & 31);
      if (c < C) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[0][w][slot]; b += red[1][w][slot]; }
        atomicAdd(dgamma + c, a);
        atomicAdd(dbeta + c, b);
      }
    }
    __syncthreads();
  }
}

template <int KIND>  // 0: nn.GELU (erf form), 1: QuickGELU x * sigmoid(1.702 x) (maskclip/model.py:166-168)
__device__ __forceinline__ float gelu_grad(float v) {
  if (KIND == 1) {  // d/dv [ v * s(av) ] = s + a v s (1 - s)
    const float sg = 1.f / (1.f + __expf(-1.702f * v));
    return sg * (1.f + 1.702f * v * (1.f - sg));
  }
  // d/dv [ v * Phi(v) ] = Phi(v) + v * phi(v)
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
  return cdf + v * pdf;
}

template <int KIND>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat162* __restrict__ dh,
                                                       const __nv_bfloat162* __restrict__ pre,
                                                       __nv_bfloat162* __restrict__ out, long long n2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const float2 d = __bfloat1622float2(dh[i]), x = __bfloat1622float2(pre[i]);
  out[i] = __floats2bfloat162_rn(d.x * gelu_grad<KIND>(x.x), d.y * gelu_grad<KIND>(x.y));
}

// 16 bytes (8 elements) per thread and step, grid-stride: the 2-element kernel ran at 45 % of the HBM copy rate on the
// [802816, 384] hidden activations of the LoftUp FeedForward backward
template <int KIND>
__global__ void __launch_bounds__(256) gelu_bwd_vec_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ pre,
                                                           uint4* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = dh[i], b = pre[i];
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 d = __bfloat1622float2(pa[e]), x = __bfloat1622float2(pb[e]);
      pr[e] = __floats2bfloat162_rn(d.x * gelu_grad<KIND>(x.x), d.y * gelu_grad<KIND>(x.y));
    }
    out[i] = r;
  }
}

// one warp per row of scores; ncols valid entries, row pitch lds / ldp; padded P entries are written as zero
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long lds,
                                                           __nv_bfloat16* __restrict__ P, long long ldp, long long R,
                                                           int ncols, int ncols_pad) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* s = S + row * lds;
  float mx = -INFINITY;
  for (int c = lane; c < ncols; c += 32) mx = fmaxf(mx, s[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < ncols; c += 32) sum += __expf(s[c] - mx);
  const float inv = 1.f / warp_sum(sum);
  __nv_bfloat16* p = P + row * ldp;
  for (int c = lane; c < ncols_pad; c += 32) p[c] = __float2bfloat16(c < ncols ? __expf(s[c] - mx) * inv : 0.f);
}

template <typename DPT>
__global__ void __launch_bounds__(256) attn_ds_kernel(const __nv_bfloat16* __restrict__ P, long long ldp,
                                                      const DPT* __restrict__ dP, long long lddp,
                                                      __nv_bfloat16* __restrict__ dS, long long ldds, long long R,
                                                      int ncols, int ncols_pad) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const __nv_bfloat16* p = P + row * ldp;
  const DPT* dp = dP + row * lddp;
  float dot = 0.f;
  for (int c = lane; c < ncols; c += 32) dot += __bfloat162float(p[c]) * (float)dp[c];
  dot = warp_sum(dot);
  __nv_bfloat16* o = dS + row * ldds;
  for (int c = lane; c < ncols_pad; c += 32)
    o[c] = __float2bfloat16(c < ncols ? __bfloat162float(p[c]) * ((float)dp[c] - dot) : 0.f);
}

// [Z][R][lds] (C valid columns) -> [Z][C][ldd] (R valid columns); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const __nv_bfloat16* __restrict__ src, long long lds,
                                                        long long src_z, __nv_bfloat16* __restrict__ dst, long long ldd,
                                                        long long dst_z, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const __nv_bfloat16* s = src + (long long)blockIdx.z * src_z;
  __nv_bfloat16* d = dst + (long long)blockIdx.z * dst_z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? s[(long long)r * lds + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) d[(long long)c * ldd + r] = tile[tx][i];
  }
}

// 64 x 64 tiles, 32-bit global accesses on both sides (two bf16 along the contiguous dimension); needs even R, C and
// pitches.  The 32 x 32 / 16-bit kernel above moved the 1.6 GB score matrices of the LoftUp backward at 1.7 TB/s.
__global__ void __launch_bounds__(256) transpose64_kernel(const __nv_bfloat16* __restrict__ src, long long lds,
                                                          long long src_z, __nv_bfloat16* __restrict__ dst, long long ldd,
                                                          long long dst_z, int R, int C) {
  __shared__ unsigned short tile[64][65];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const unsigned short* s = reinterpret_cast<const unsigned short*>(src) + (long long)blockIdx.z * src_z;
  unsigned short* d = reinterpret_cast<unsigned short*>(dst) + (long long)blockIdx.z * dst_z;
  const int w = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = ty + 8 * k;
    unsigned v = 0;
    if (r0 + r < R && c0 + 2 * w < C) v = *reinterpret_cast<const unsigned*>(s + (long long)(r0 + r) * lds + c0 + 2 * w);
    tile[r][2 * w] = (unsigned short)(v & 0xffffu);
    tile[r][2 * w + 1] = (unsigned short)(v >> 16);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = ty + 8 * k;
    if (c0 + c < C && r0 + 2 * w < R) {
      const unsigned v = (unsigned)tile[2 * w][c] | ((unsigned)tile[2 * w + 1][c] << 16);
      *reinterpret_cast<unsigned*>(d + (long long)(c0 + c) * ldd + r0 + 2 * w) = v;
    }
  }
}

}  // namespace vb
}  // namespace isp

using namespace isp;

// Backward of LayerNorm over the last dimension of x [M, C] (fp32 | bf16, row pitch ldx), affine weight gamma:
// dx = LN'(x)^T (gamma * dy) (+ resid, the gradient arriving through the residual connection).  dx_bf16 (optional)
// receives a bf16 copy for the next GEMM.  C <= 1024.
extern "C" int isp_layernorm_rows_bwd(const float* dy, long long lddy, const void* x, int x_bf16, long long ldx, const float* gamma,
                                      const float* resid, long long ldr, float* dx, long long lddx, void* dx_bf16,
                                      long long ldb, long long M, int C, float eps, isp_stream_t stream) {
  ISP_REQUIRE(dy && x && gamma && dx, ISP_ERR_BAD_SHAPE, "layernorm_rows_bwd: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && C <= 32 * vb::kMaxPerLane, ISP_ERR_BAD_SHAPE, "layernorm_rows_bwd: bad shape (C <= 1024)");
  const int xs = x_bf16 ? 2 : 4;
  const bool vec = C % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && (ldx * xs) % (x_bf16 ? 8 : 16) == 0 && aligned16(dy) &&
                   aligned16(dx) && aligned16(gamma) && ((uintptr_t)x % 16 == 0) && (!resid || (ldr % 4 == 0 && aligned16(resid))) &&
                   (!dx_bf16 || (ldb % 4 == 0 && (uintptr_t)dx_bf16 % 8 == 0));
  if (vec) {
#define ISP_LN_BWD_VEC(XB, KIT)                                                                                       \
  vb::layernorm_bwd_vec_kernel<XB, KIT><<<cdiv(M, 8), 256, 0, as_stream(stream)>>>(                                     \
      dy, lddy, x, ldx, gamma, resid, ldr, dx, lddx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), ldb, M, C, eps)
    if (C <= 512) {  // half the registers of the general instantiation: more rows in flight per SM
      if (x_bf16) ISP_LN_BWD_VEC(true, 4); else ISP_LN_BWD_VEC(false, 4);
    } else {
      if (x_bf16) ISP_LN_BWD_VEC(true, 8); else ISP_LN_BWD_VEC(false, 8);
    }
#undef ISP_LN_BWD_VEC
    ISP_CHECK_LAUNCH("layernorm_bwd_vec_kernel");
    return ISP_OK;
  }
  if (x_bf16)
    vb::layernorm_bwd_kernel<true><<<cdiv(M, 8), 256, 0, as_stream(stream)>>>(
        dy, lddy, x, ldx, gamma, resid, ldr, dx, lddx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), ldb, M, C, eps);
  else
    vb::layernorm_bwd_kernel<false><<<cdiv(M, 8), 256, 0, as_stream(stream)>>>(
        dy, lddy, x, ldx, gamma, resid, ldr, dx, lddx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), ldb, M, C, eps);
  ISP_CHECK_LAUNCH("layernorm_bwd_kernel");
  return ISP_OK;
}

// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy for LayerNorm over the last dimension of x [M, C] (fp32 | bf16);
// both outputs are ACCUMULATED (zero them first).  C <= 1024.
extern "C" int isp_layernorm_affine_bwd(const float* dy, long long lddy, const void* x, int x_bf16, long long ldx,
                                        float* dgamma, float* dbeta, long long M, int C, float eps, isp_stream_t stream) {
  ISP_REQUIRE(dy && x && dgamma && dbeta && M > 0 && C > 0 && C <= 32 * vb::kMaxPerLane, ISP_ERR_BAD_SHAPE,
              "layernorm_affine_bwd: bad arguments (C <= 1024)");
  const int rows = 64;
  if (x_bf16)
    vb::layernorm_affine_bwd_kernel<true><<<cdiv(M, rows), 256, 0, as_stream(stream)>>>(dy, lddy, x, ldx, dgamma, dbeta, M, C,
                                                                                      eps, rows);
  else
    vb::layernorm_affine_bwd_kernel<false><<<cdiv(M, rows), 256, 0, as_stream(stream)>>>(dy, lddy, x, ldx, dgamma, dbeta, M,
                                                                                       C, eps, rows);
  ISP_CHECK_LAUNCH("layernorm_affine_bwd_kernel");
  return ISP_OK;
}

// dpre = dh * act'(pre) on dense bf16 arrays of n elements (n even); quick == 0: nn.GELU erf form
// (dinov2/layers/mlp.py:34-40), quick == 1: CLIP's QuickGELU.
extern "C" int isp_gelu_bwd_bf16(const void* dh, const void* pre, void* out, long long n, int quick, isp_stream_t stream) {
  ISP_REQUIRE(dh && pre && out && n > 0 && n % 2 == 0, ISP_ERR_BAD_SHAPE, "gelu_bwd_bf16: bad arguments");
  if (n % 8 == 0 && aligned16(dh) && aligned16(pre) && aligned16(out)) {
    const long long n8 = n / 8;
    const unsigned grid = (unsigned)(cdiv(n8, 256) < 148 * 16 ? cdiv(n8, 256) : 148 * 16);
    if (quick)
      vb::gelu_bwd_vec_kernel<1><<<grid, 256, 0, as_stream(stream)>>>((const uint4*)dh, (const uint4*)pre, (uint4*)out, n8);
    else
      vb::gelu_bwd_vec_kernel<0><<<grid, 256, 0, as_stream(stream)>>>((const uint4*)dh, (const uint4*)pre, (uint4*)out, n8);
    ISP_CHECK_LAUNCH("gelu_bwd_vec_kernel");
    return ISP_OK;
  }
  if (quick)
    vb::gelu_bwd_kernel<1><<<cdiv(n / 2, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat162*>(dh), reinterpret_cast<const __nv_bfloat162*>(pre),
        reinterpret_cast<__nv_bfloat162*>(out), n / 2);
  else
    vb::gelu_bwd_kernel<0><<<cdiv(n / 2, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat162*>(dh), reinterpret_cast<const __nv_bfloat162*>(pre),
        reinterpret_cast<__nv_bfloat162*>(out), n / 2);
  ISP_CHECK_LAUNCH("gelu_bwd_kernel");
  return ISP_OK;
}

// P[r, :ncols] = softmax(S[r, :ncols]) (fp32 scores -> bf16 probabilities), P[r, ncols:ncols_pad] = 0.
extern "C" int isp_softmax_rows(const float* S, long long lds, void* P_bf16, long long ldp, long long R, int ncols,
                                int ncols_pad, isp_stream_t stream) {
  ISP_REQUIRE(S && P_bf16 && R > 0 && ncols > 0 && ncols_pad >= ncols && lds >= ncols && ldp >= ncols_pad, ISP_ERR_BAD_SHAPE,
              "softmax_rows: bad arguments");
  vb::softmax_rows_kernel<<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(S, lds, reinterpret_cast<__nv_bfloat16*>(P_bf16), ldp,
                                                                    R, ncols, ncols_pad);
  ISP_CHECK_LAUNCH("softmax_rows_kernel");
  return ISP_OK;
}

// dS[r, c] = P[r, c] * (dP[r, c] - sum_j P[r, j] dP[r, j]) for c < ncols, 0 up to ncols_pad.
extern "C" int isp_attn_ds_rows(const void* P_bf16, long long ldp, const void* dP, int dp_bf16, long long lddp,
                                void* dS_bf16, long long ldds, long long R, int ncols, int ncols_pad, isp_stream_t stream) {
  ISP_REQUIRE(P_bf16 && dP && dS_bf16 && R > 0 && ncols > 0 && ncols_pad >= ncols, ISP_ERR_BAD_SHAPE,
              "attn_ds_rows: bad arguments");
  if (dp_bf16)
    vb::attn_ds_kernel<__nv_bfloat16><<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(P_bf16), ldp, reinterpret_cast<const __nv_bfloat16*>(dP), lddp,
        reinterpret_cast<__nv_bfloat16*>(dS_bf16), ldds, R, ncols, ncols_pad);
  else
    vb::attn_ds_kernel<float><<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(P_bf16), ldp, reinterpret_cast<const float*>(dP), lddp,
        reinterpret_cast<__nv_bfloat16*>(dS_bf16), ldds, R, ncols, ncols_pad);
  ISP_CHECK_LAUNCH("attn_ds_kernel");
  return ISP_OK;
}

// dst[z][c][r] = src[z][r][c] for Z matrices of R x C bf16 (row pitches lds / ldd, matrix pitches src_z / dst_z).
extern "C" int isp_transpose_bf16_batched(const void* src, long long lds, long long src_z, void* dst, long long ldd,
                                          long long dst_z, int Z, int R, int C, isp_stream_t stream) {
  ISP_REQUIRE(src && dst && Z > 0 && R > 0 && C > 0 && lds >= C && ldd >= R && Z <= 65535, ISP_ERR_BAD_SHAPE,
              "transpose_bf16_batched: bad arguments");
  ISP_REQUIRE(cdiv(R, 32) <= 65535, ISP_ERR_UNSUPPORTED, "transpose_bf16_batched: too many rows");
  if (R % 2 == 0 && C % 2 == 0 && lds % 2 == 0 && ldd % 2 == 0 && src_z % 2 == 0 && dst_z % 2 == 0 &&
      (uintptr_t)src % 4 == 0 && (uintptr_t)dst % 4 == 0) {
    const dim3 grid64((unsigned)cdiv(C, 64), (unsigned)cdiv(R, 64), (unsigned)Z);
    vb::transpose64_kernel<<<grid64, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds, src_z,
                                                               reinterpret_cast<__nv_bfloat16*>(dst), ldd, dst_z, R, C);
    ISP_CHECK_LAUNCH("transpose64_kernel");
    return ISP_OK;
  }
  const dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, 32), (unsigned)Z);
  vb::transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), lds, src_z,
                                                           reinterpret_cast<__nv_bfloat16*>(dst), ldd, dst_z, R, C);
  ISP_CHECK_LAUNCH("transpose_kernel");
  return ISP_OK;
}
