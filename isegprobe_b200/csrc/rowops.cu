// Row-wise / elementwise producers and consumers around the tensor-core kernels:
// LayerNorm over channels, LoftUp's Fourier-feature + ChannelNorm producer, batch-global
// min/max, low-res key/value preparation, attention operand repacking, ViT patchify.
// All bandwidth-bound; one warp per row with 128-bit accesses where the layout allows.
#include "common.cuh"

namespace isp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ld_any(const void* p, long long i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, long long i, float v, int is_bf16) {
  if (is_bf16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// ---------------------------------------------------------------------------
// LayerNorm over the last dim of a row-major matrix (nn.LayerNorm / loftup layers.py:26-58):
// y = (x - mean) / sqrt(biased_var + eps) * gamma + beta.  One warp per row, C <= 1024.
// Columns C..ldo-1 of the output row are zero-filled (K padding for the next GEMM).
constexpr int kLnMaxPerLane = 32;
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const void* __restrict__ in, int in_bf16, long long ldi,
                                                             void* __restrict__ out, int out_bf16, long long ldo,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long M, int C,
                                                             float eps) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float x[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + i * 32;
    x[i] = (c < C) ? ld_any(in, row * ldi + c, in_bf16) : 0.f;
    s += x[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + i * 32;
    const float d = (c < C) ? x[i] - mean : 0.f;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + i * 32;
    if (c < C) st_any(out, row * ldo + c, (x[i] - mean) * rstd * gamma[c] + beta[c], out_bf16);
    else if (c < ldo) st_any(out, row * ldo + c, 0.f, out_bf16);
  }
}

// Vectorised variant: each lane owns V groups of 8 consecutive columns (128-bit loads/stores for
// bf16, 2 x 128-bit for f32); needs 16-byte aligned rows.  This is the one the hot path uses.
template <bool IN_BF16>
__device__ __forceinline__ void ld8(const void* p, long long idx, float (&v)[8]) {
  if constexpr (IN_BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + idx);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + idx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <bool OUT_BF16>
__device__ __forceinline__ void st8(void* p, long long idx, const float (&v)[8]) {
  if constexpr (OUT_BF16) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Persistent: each warp walks rows with a grid stride, R rows per iteration (all loads of the R
// rows are issued before the first reduction), gamma / beta staged once per block in shared
// memory (the per-element __ldg of the first version made the kernel LSU-bound at 1.8 TB/s).
// bf16 rows stay PACKED in registers (4 per 8 columns) and are unpacked in each of the three
// passes, so R = 8 rows are in flight per warp: the kernel is latency-bound
// otherwise (a bf16 row of 416 columns is only 832 bytes); 2 blocks / SM.
template <bool IN_BF16>
struct LnRaw {
  uint4 a, b;  // bf16: a only
};
template <bool IN_BF16>
__device__ __forceinline__ void ln_unpack(LnRaw<IN_BF16>& r, float (&v)[8]) {
  if constexpr (IN_BF16) {
    // opaque to the optimiser: otherwise the three passes share one unpack and all R*V*8 floats stay live (spills)
    asm volatile("" : "+r"(r.a.x), "+r"(r.a.y), "+r"(r.a.z), "+r"(r.a.w));
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
    v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
  }
}

template <bool IN_BF16, bool OUT_BF16, int V>
__global__ void __launch_bounds__(256, 2) layernorm_rows_vec_kernel(const void* __restrict__ in, long long ldi,
                                                                 void* __restrict__ out, long long ldo,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, long long M, int C,
                                                                 float eps) {
  constexpr int R = V <= 2 ? 4 : 2;
  __shared__ __align__(16) float g_s[V * 256], b_s[V * 256];
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8;
  for (int c = threadIdx.x; c < V * 256; c += 256) {
    g_s[c] = c < C ? __ldg(gamma + c) : 0.f;
    b_s[c] = c < C ? __ldg(beta + c) : 0.f;
  }
  __syncthreads();
  const float invC = 1.f / (float)C;
  // columns c0 .. c0+7 of vector i; a vector is live if it starts inside the row (the caller guarantees
  // 8-element alignment of ldi, so a live vector can always be loaded whole; columns >= C are masked)
  for (long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R; row0 < M; row0 += nwarps * R) {
    LnRaw<IN_BF16> raw[R][V];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        raw[r][i].a = make_uint4(0u, 0u, 0u, 0u);
        raw[r][i].b = make_uint4(0u, 0u, 0u, 0u);
        if (row < M && c0 < C) {
          if constexpr (IN_BF16) {
            raw[r][i].a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in) + row * ldi + c0);
          } else {
            raw[r][i].a = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(in) + row * ldi + c0);
            raw[r][i].b = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(in) + row * ldi + c0 + 4);
          }
        }
      }
    }
    float s[R], v[R], mean[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s[r] = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        float x[8];
        ln_unpack<IN_BF16>(raw[r][i], x);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[r] += (c0 + e < C) ? x[e] : 0.f;  // pad columns of the input row are ignored
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] = s[r] * invC;
      v[r] = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        float x[8];
        ln_unpack<IN_BF16>(raw[r][i], x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = (c0 + e < C) ? x[e] - mean[r] : 0.f;
          v[r] = fmaf(d, d, v[r]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] += __shfl_xor_sync(0xffffffffu, v[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      if (row >= M) break;
      const float rstd = rsqrtf(v[r] * invC + eps);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c0 = (lane + 32 * i) * 8;
        if (c0 >= ldo) continue;
        float x[8], y[8], g[8], bt[8];
        ln_unpack<IN_BF16>(raw[r][i], x);
        *reinterpret_cast<float4*>(&g[0]) = *reinterpret_cast<const float4*>(g_s + c0);
        *reinterpret_cast<float4*>(&g[4]) = *reinterpret_cast<const float4*>(g_s + c0 + 4);
        *reinterpret_cast<float4*>(&bt[0]) = *reinterpret_cast<const float4*>(b_s + c0);
        *reinterpret_cast<float4*>(&bt[4]) = *reinterpret_cast<const float4*>(b_s + c0 + 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = fmaf((x[e] - mean[r]) * rstd, g[e], bt[e]);  // pads: g = b = 0
        st8<OUT_BF16>(out, row * ldo + c0, y);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Per-channel min / max of a [B,3,H,W] image over batch + space (MinMaxScaler,
// loftup/layers.py:66-71).  mm[0..2] = min, mm[3..5] = max; order-preserving int atomics.
__device__ __forceinline__ int float_ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = float_ord(INFINITY);
  else if (threadIdx.x < 6) mm[threadIdx.x] = float_ord(-INFINITY);
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ img, int* __restrict__ mm, int B,
                                                     long long HW, long long sb, long long sc) {
  const int c = blockIdx.y;
  float mn = INFINITY, mx = -INFINITY;
  for (int b = 0; b < B; ++b) {
    const float* p = img + b * sb + c * sc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
      const float v = p[i];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[c], float_ord(mn));
    atomicMax(&mm[3 + c], float_ord(mx));
  }
}

// ---------------------------------------------------------------------------
// LoftUp query producer: MinMaxScaler -> ImplicitFeaturizer(color, 20 freqs, learned bias)
// -> ChannelNorm(203)  (loftup/layers.py:61-158, loftup/loftup.py:50-56), one warp per pixel,
// output bf16 NHWC [B,H,W,ldo] (203 real channels, rest zero).  fp32 throughout; the
// argument u*f + b is formed with explicit mul/add (no FMA contraction) and accurate
// sinf/cosf because |u*f| reaches 2.2e4.  gridr/gridc/freqs are the host's
// torch.linspace / torch.exp values so the reference's CPU rounding is reproduced exactly.
// sin / cos of a large fp32 argument (the Fourier features reach |arg| ~ 2e4: frequencies up to e^10 on coordinates in
// [-1, 1]): reduce in turns with an FMA pair on a two-word 1 / 2pi (error ~1e-7 turns), then one MUFU on |r| <= pi.
// ~8 instructions instead of sinf's ~25; the result is within 2e-6 of the correctly rounded value of the SAME fp32 argument
// (the argument itself is formed with the reference's operation order), far below the bf16 rounding of the output.
__device__ __forceinline__ float sincos_turns(float a, float quarter) {  // quarter = 0.25 for cos (cos x = sin(x + pi/2)), 0 for sin
  const float kInv2PiHi = 0.15915494f, kInv2PiLo = 6.4206383e-9f;
  const float k = rintf(a * kInv2PiHi);
  float r = fmaf(a, kInv2PiHi, -k);
  r = fmaf(a, kInv2PiLo, r);
  return __sinf((r + quarter) * 6.28318530717958647692f);  // |r| <= 0.5 turns: the MUFU argument stays inside [-pi/2, 3pi/2]
}

constexpr int kFourierPix = 16;  // pixels per warp: the lane's channel constants are loaded once for all of them
__global__ void __launch_bounds__(256) fourier_chnorm_kernel(
    const float* __restrict__ img, long long sb, long long sc, long long sh, long long sw, const int* __restrict__ mm,
    const float* __restrict__ gridr, const float* __restrict__ gridc, const float* __restrict__ freqs,
    const float* __restrict__ bias_sin, const float* __restrict__ bias_cos, const float* __restrict__ gamma,
    const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int B, int H, int W, int ldo, float eps) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * H * W;
  const long long pix0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * kFourierPix;
  if (pix0 >= total) return;
  // lane owns the channel pairs (2 lane, 2 lane + 1) + 64 i: bf16x2 stores, 128 bytes per warp and instruction.  Per channel:
  // which of the five inputs it reads (kind 0..4, 5 = raw colour pass-through, 6 = padding), frequency, phase, affine
  float cf[8], cb[8], cg[8], ce[8];
  int kind[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = 2 * lane + (i >> 1) * 64 + (i & 1);
    cf[i] = 0.f; cb[i] = 0.f; kind[i] = 6;
    if (c < 200) {
      const int cc = c < 100 ? c : c - 100;
      const int f = cc / 5;
      kind[i] = cc - f * 5;
      cf[i] = freqs[f];
      cb[i] = c < 100 ? bias_sin[cc] : bias_cos[cc];
    } else if (c < 203) {
      kind[i] = 5;
    }
    cg[i] = c < 203 ? gamma[c] : 0.f;
    ce[i] = c < 203 ? beta[c] : 0.f;
  }
  float mn[3], sc_[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mn[c] = ord_float(mm[c]);
    sc_[c] = fmaxf(__fsub_rn(ord_float(mm[3 + c]), mn[c]), 1e-4f);
  }
  const long long pend = pix0 + kFourierPix < total ? pix0 + kFourierPix : total;
  // (b, h, w) of the first pixel by division once, then carried along (64-bit div / mod per pixel was a third of the kernel)
  int w = (int)(pix0 % W), h = (int)((pix0 / W) % H), b = (int)(pix0 / ((long long)W * H));
  float quarter[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) quarter[i] = (2 * lane + (i >> 1) * 64 + (i & 1)) >= 100 ? 0.25f : 0.f;
  // lane j < 5 holds input j of the pixel (row coordinate, column coordinate, three scaled colours); channel i of a lane
  // fetches its input with one shuffle from lane src[i] (pass-through channels 200..202: inputs 2..4)
  int src[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = 2 * lane + (i >> 1) * 64 + (i & 1);
    src[i] = kind[i] < 5 ? kind[i] : (kind[i] == 5 ? c - 198 : 0);
  }
  const int cl = (lane >= 2 && lane < 5) ? lane - 2 : 0;
  const float my_mn = cl == 0 ? mn[0] : (cl == 1 ? mn[1] : mn[2]), my_sc = cl == 0 ? sc_[0] : (cl == 1 ? sc_[1] : sc_[2]);
  for (long long pix = pix0; pix < pend; ++pix) {
    const float x = img[b * sb + cl * sc + h * sh + w * sw];
    float uval = __fsub_rn(__fdiv_rn(__fsub_rn(x, my_mn), my_sc), 0.5f);
    uval = lane == 0 ? gridr[h] : (lane == 1 ? gridc[w] : uval);
    if (++w == W) { w = 0; if (++h == H) { h = 0; ++b; } }
    float val[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ud = __shfl_sync(0xffffffffu, uval, src[i]);
      const float arg = __fadd_rn(__fmul_rn(ud, cf[i]), cb[i]);
      float v = sincos_turns(arg, quarter[i]);
      if (i >= 6) v = kind[i] >= 5 ? (kind[i] == 6 ? 0.f : ud) : v;  // only channels >= 192 can be pass-through / padding
      val[i] = v;
      s += v;
    }
    const float mean = warp_sum(s) * (1.f / 203.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = kind[i] < 6 ? val[i] - mean : 0.f;
      var += d * d;
    }
    const float rstd = rsqrtf(warp_sum(var) * (1.f / 203.f) + eps);
    __nv_bfloat16* o = out + pix * ldo;  // ldo is even and the rows 4-byte aligned (checked by the entry point)
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const int c = 2 * lane + (i >> 1) * 64;
      if (c < ldo) {
        const float a0 = kind[i] < 6 ? (val[i] - mean) * rstd * cg[i] + ce[i] : 0.f;
        const float a1 = kind[i + 1] < 6 ? (val[i + 1] - mean) * rstd * cg[i + 1] + ce[i + 1] : 0.f;
        *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(a0, a1);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// LoftUp key/value source: ChannelNorm(C) on the LR features (wrapper, loftup.py:141-149),
// then cat with the 20-channel sine PE of the LR grid (loftup.py:115-118) -> fp32 [B*h*w, C+20].
// One warp per LR pixel; lr is [B,C,h,w] with arbitrary strides.
__global__ void __launch_bounds__(256) lr_prepare_kernel(const float* __restrict__ lr, long long sb, long long sc,
                                                         long long sh, long long sw, const float* __restrict__ cn_w,
                                                         const float* __restrict__ cn_b, int use_cn,
                                                         const float* __restrict__ gridr, const float* __restrict__ gridc,
                                                         const float* __restrict__ freqs5,
                                                         const float* __restrict__ bias_sin,
                                                         const float* __restrict__ bias_cos, float* __restrict__ out,
                                                         int B, int C, int h, int w, float eps) {
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= (long long)B * h * w) return;
  const int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / ((long long)w * h));
  const float* p = lr + b * sb + y * sh + x * sw;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += p[c * sc];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = p[c * sc] - mean; v += d * d; }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
  float* o = out + pix * (C + 20);
  for (int c = lane; c < C; c += 32) {
    const float val = p[c * sc];
    o[c] = use_cn ? (val - mean) * rstd * cn_w[c] + cn_b[c] : val;
  }
  if (lane < 20) {  // channel f*2+d: sin (0..9) then cos (10..19)
    const int cc = lane < 10 ? lane : lane - 10;
    const int f = cc >> 1, d = cc & 1;
    const float arg = __fadd_rn(__fmul_rn(d ? gridc[x] : gridr[y], freqs5[f]), lane < 10 ? bias_sin[cc] : bias_cos[cc]);
    o[C + lane] = lane < 10 ? sinf(arg) : cosf(arg);
  }
}

// ---------------------------------------------------------------------------
// Attention operand repack.  src: [B*T, ld] (f32|bf16) holding per-head slices at column
// col0 + head*hd.  Writes K-style [B, heads, Tpad, DKC] (row = token, zero padded) when
// transpose == 0, or V^T-style [B, heads, DV, Tpad] when transpose == 1 (bf16).
// One block = 64 tokens of one (batch, head): the [64][hd] slice is read with d fastest (coalesced rows),
// staged in shared memory, and written either as padded rows (transpose == 0) or transposed with the
// token index fastest (transpose == 1), so both sides move in full 128-byte segments.
constexpr int kRepackT = 64;
__global__ void __launch_bounds__(256) repack_heads_kernel(const void* __restrict__ src, int src_bf16, long long ld,
                                                           int col0, int hd, __nv_bfloat16* __restrict__ dst, int B,
                                                           int T, int Tpad, int heads, int D, int transpose) {
  extern __shared__ __align__(4) __nv_bfloat16 rp_tile[];  // [kRepackT][pitch], pitch = hd + 2 - (hd & 1): even, 33 (mod 32) words
  const int t0 = blockIdx.x * kRepackT, hh = blockIdx.y, b = blockIdx.z;
  const int pitch = hd + 2 - (hd & 1);
  const int nt = min(kRepackT, Tpad - t0);
  // pairs of columns where the source allows it (bf16, even head width / offsets): 4-byte loads, stores and smem accesses
  const bool pairs = src_bf16 && !(hd & 1) && !((col0 + hh * hd) & 1) && !(ld & 1) &&
                     !(reinterpret_cast<uintptr_t>(src) & 3);
  if (pairs) {
    const int hd2 = hd >> 1;
    const __nv_bfloat16* s16 = static_cast<const __nv_bfloat16*>(src);
    for (int e = threadIdx.x; e < kRepackT * hd2; e += blockDim.x) {
      const int tt = e / hd2, d = (e - tt * hd2) * 2, t = t0 + tt;
      uint32_t v = 0u;
      if (t < T) v = *reinterpret_cast<const uint32_t*>(s16 + ((long long)b * T + t) * ld + col0 + hh * hd + d);
      *reinterpret_cast<uint32_t*>(rp_tile + tt * pitch + d) = v;
    }
  } else {
    for (int e = threadIdx.x; e < kRepackT * hd; e += blockDim.x) {
      const int tt = e / hd, d = e - tt * hd, t = t0 + tt;
      float v = 0.f;
      if (t < T) v = ld_any(src, ((long long)b * T + t) * ld + col0 + hh * hd + d, src_bf16);
      rp_tile[tt * pitch + d] = __float2bfloat16(v);
    }
  }
  __syncthreads();
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  __nv_bfloat16* out = dst + ((long long)b * heads + hh) * (long long)Tpad * D;
  if (!transpose) {
    if (pairs && !(D & 1)) {
      const int D2 = D >> 1;
      for (int e = threadIdx.x; e < nt * D2; e += blockDim.x) {
        const int tt = e / D2, d = (e - tt * D2) * 2;
        const uint32_t v = d < hd ? *reinterpret_cast<const uint32_t*>(rp_tile + tt * pitch + d) : 0u;
        *reinterpret_cast<uint32_t*>(out + (long long)(t0 + tt) * D + d) = v;
      }
    } else {
      for (int e = threadIdx.x; e < nt * D; e += blockDim.x) {
        const int tt = e / D, d = e - tt * D;
        out[(long long)(t0 + tt) * D + d] = d < hd ? rp_tile[tt * pitch + d] : zero;
      }
    }
  } else if (!(Tpad & 1)) {  // two tokens per 4-byte store (t0 is a multiple of 64)
    for (int e = threadIdx.x; e < D * (kRepackT / 2); e += blockDim.x) {
      const int d = e / (kRepackT / 2), tt = (e - d * (kRepackT / 2)) * 2;
      if (tt < nt) {  // nt is even here: Tpad and t0 are
        __nv_bfloat162 v;
        v.x = d < hd ? rp_tile[tt * pitch + d] : zero;
        v.y = d < hd ? rp_tile[(tt + 1) * pitch + d] : zero;
        *reinterpret_cast<__nv_bfloat162*>(out + (long long)d * Tpad + t0 + tt) = v;
      }
    }
  } else {
    for (int e = threadIdx.x; e < D * kRepackT; e += blockDim.x) {
      const int d = e / kRepackT, tt = e - d * kRepackT;
      if (tt < nt) out[(long long)d * Tpad + t0 + tt] = d < hd ? rp_tile[tt * pitch + d] : zero;
    }
  }
}

// ---------------------------------------------------------------------------
// ViT patchify: [B,Cin,H,W] f32 (strided) -> bf16 [B*nh*nw, ldo], column = c*p*p + i*p + j
// (the flattening order of Conv2d(k=s=p) weights), zero padded to ldo.
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, long long sb, long long sc,
                                                       long long sh, long long sw, __nv_bfloat16* __restrict__ out,
                                                       int B, int Cin, int H, int W, int P, int ldo) {
  const int nh = H / P, nw = W / P, K = Cin * P * P;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * nh * nw * ldo;
  if (idx >= total) return;
  const int k = (int)(idx % ldo);
  const long long tok = idx / ldo;
  float v = 0.f;
  if (k < K) {
    const int j = k % P, i = (k / P) % P, c = k / (P * P);
    const int px = (int)(tok % nw), py = (int)((tok / nw) % nh), b = (int)(tok / ((long long)nw * nh));
    v = img[b * sb + c * sc + (py * P + i) * sh + (px * P + j) * sw];
  }
  out[idx] = __float2bfloat16(v);
}

// ViT token assembly: x[b,0,:] = cls + pos[0]; x[b,1+n,:] = patch[b,n,:] (+ extra[b,n,:]) + pos[1+n]
// (DINOv2.py:518-529).  fp32 out [B, T, C].
__global__ void __launch_bounds__(256) vit_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ extra,
                                                         const float* __restrict__ cls, const float* __restrict__ pos,
                                                         float* __restrict__ out, int B, int N, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * (N + 1) * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int t = (int)((idx / C) % (N + 1));
  const int b = (int)(idx / ((long long)C * (N + 1)));
  float v;
  if (t == 0) v = cls[c];
  else {
    const long long src = ((long long)b * N + (t - 1)) * C + c;
    v = patch[src] + (extra ? extra[src] : 0.f);
  }
  out[idx] = v + pos[(long long)t * C + c];
}

// ---------------------------------------------------------------------------
// Classifier of the IS head: Conv2d(C -> K, 1x1) on an NHWC activation
// (heads/base_head.py:16, conv_heads.py:72): out[m, k] = sum_c x[m, c] * w[k, c] + b[k].
// One warp per pixel, K <= 8 classes; output [M, K] fp32 (== NCHW when K == 1).
__global__ void __launch_bounds__(256) rowdot_kernel(const void* __restrict__ x, int x_bf16, long long ld,
                                                     const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ out, long long M, int C, int K) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < C; c += 32) {
    const float v = ld_any(x, row * ld + c, x_bf16);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < K) acc[k] = fmaf(v, w[k * C + c], acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < K) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) out[row * K + k] = s + (b ? b[k] : 0.f);
    }
  }
}

// K == 1, bf16 rows of C <= 512 channels (C % 8 == 0, 16-byte aligned): 16-byte loads, the lane's weights in registers for
// all of the warp's rows, four rows in flight per warp.  HBM-bound (the scalar kernel above moved 1.2 TB/s: 0.52 ms per
// 802816 x 384 activation, 4 ms of every LoftUp e2e step at batch 32).
__global__ void __launch_bounds__(256) rowdot1_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ out, long long M, int C) {
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;  // 16-byte vectors per row (<= 64)
  float wr[2][8];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) wr[j][e] = (lane + 32 * j < nvec) ? w[(lane + 32 * j) * 8 + e] : 0.f;
  const float bias = b ? b[0] : 0.f;
  const long long nwarps = (long long)gridDim.x * 8, wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (long long r0 = wid * 4; r0 < M; r0 += nwarps * 4) {
    float acc[4];
    uint4 v[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        v[i][j] = (r0 + i < M && lane + 32 * j < nvec)
                      ? *reinterpret_cast<const uint4*>(x + (r0 + i) * ld + (lane + 32 * j) * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t u[4] = {v[i][j].x, v[i][j].y, v[i][j].z, v[i][j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a = fmaf(__uint_as_float(u[k] << 16), wr[j][2 * k], a);
          a = fmaf(__uint_as_float(u[k] & 0xffff0000u), wr[j][2 * k + 1], a);
        }
      }
      acc[i] = warp_sum(a);
    }
    if (lane < 4 && r0 + lane < M) out[r0 + lane] = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + bias;
  }
}

}  // namespace isp

using namespace isp;

extern "C" int isp_layernorm_rows(const void* in, int in_bf16, long long ldi, void* out, int out_bf16, long long ldo,
                                  const float* gamma, const float* beta, long long M, int C, float eps,
                                  isp_stream_t stream) {
  ISP_REQUIRE(in && out && gamma && beta, ISP_ERR_BAD_SHAPE, "layernorm_rows: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && C <= 32 * kLnMaxPerLane && ldi >= C && ldo >= C && ldo <= 32 * kLnMaxPerLane,
              ISP_ERR_BAD_SHAPE, "layernorm_rows: bad shape M=%lld C=%d ldi=%lld ldo=%lld", M, C, ldi, ldo);
  if (ldi % 8 == 0 && ldo % 8 == 0 && aligned16(in) && aligned16(out) && ldo <= 1024) {
    const int nvec = (int)((ldo + 7) / 8);
    const int V = (nvec + 31) / 32;  // 8-element vectors per lane
    dim3 grid((unsigned)min((long long)cdiv(M, 16), (long long)148 * 8));
#define ISP_LN_LAUNCH(IB, OB, VV)                                                                        \
  layernorm_rows_vec_kernel<IB, OB, VV><<<grid, 256, 0, as_stream(stream)>>>(in, ldi, out, ldo, gamma, beta, M, C, eps)
#define ISP_LN_V(IB, OB)                                                           \
  do {                                                                             \
    if (V == 1) ISP_LN_LAUNCH(IB, OB, 1); else if (V == 2) ISP_LN_LAUNCH(IB, OB, 2); \
    else if (V == 3) ISP_LN_LAUNCH(IB, OB, 3); else ISP_LN_LAUNCH(IB, OB, 4);       \
  } while (0)
    if (in_bf16 && out_bf16) ISP_LN_V(true, true);
    else if (in_bf16) ISP_LN_V(true, false);
    else if (out_bf16) ISP_LN_V(false, true);
    else ISP_LN_V(false, false);
#undef ISP_LN_V
#undef ISP_LN_LAUNCH
    ISP_CHECK_LAUNCH("layernorm_rows_vec_kernel");
    return ISP_OK;
  }
  layernorm_rows_kernel<<<cdiv(M, 8), 256, 0, as_stream(stream)>>>(in, in_bf16, ldi, out, out_bf16, ldo, gamma, beta, M,
                                                                 C, eps);
  ISP_CHECK_LAUNCH("layernorm_rows_kernel");
  return ISP_OK;
}

extern "C" int isp_minmax_per_channel(const float* img, int* mm6, int B, int H, int W, long long sb, long long sc,
                                      isp_stream_t stream) {
  ISP_REQUIRE(img && mm6 && B > 0 && H > 0 && W > 0, ISP_ERR_BAD_SHAPE, "minmax_per_channel: bad arguments");
  minmax_init_kernel<<<1, 32, 0, as_stream(stream)>>>(mm6);
  ISP_CHECK_LAUNCH("minmax_init_kernel");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)min((long long)296, (HW + 255) / 256), 3);
  minmax_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, mm6, B, HW, sb, sc);
  ISP_CHECK_LAUNCH("minmax_kernel");
  return ISP_OK;
}

extern "C" int isp_loftup_fourier_chnorm(const float* img, long long sb, long long sc, long long sh, long long sw,
                                         const int* mm6, const float* gridr, const float* gridc, const float* freqs20,
                                         const float* bias_sin, const float* bias_cos, const float* gamma,
                                         const float* beta, void* out_bf16, int B, int H, int W, int ldo, float eps,
                                         isp_stream_t stream) {
  ISP_REQUIRE(img && mm6 && gridr && gridc && freqs20 && bias_sin && bias_cos && gamma && beta && out_bf16,
              ISP_ERR_BAD_SHAPE, "loftup_fourier_chnorm: null pointer");
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && ldo >= 203 && ldo <= 224, ISP_ERR_BAD_SHAPE,
              "loftup_fourier_chnorm: bad shape (ldo must be in [203,224])");
  ISP_REQUIRE(ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 3) == 0, ISP_ERR_MISALIGNED,
              "loftup_fourier_chnorm: ldo must be even and the output 4-byte aligned");
  const long long total = (long long)B * H * W;
  fourier_chnorm_kernel<<<cdiv(total, 8 * kFourierPix), 256, 0, as_stream(stream)>>>(
      img, sb, sc, sh, sw, mm6, gridr, gridc, freqs20, bias_sin, bias_cos, gamma, beta,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), B, H, W, ldo, eps);
  ISP_CHECK_LAUNCH("fourier_chnorm_kernel");
  return ISP_OK;
}

extern "C" int isp_loftup_lr_prepare(const float* lr, long long sb, long long sc, long long sh, long long sw,
                                     const float* cn_w, const float* cn_b, const float* gridr, const float* gridc,
                                     const float* freqs5, const float* bias_sin, const float* bias_cos, float* out,
                                     int B, int C, int h, int w, float eps, isp_stream_t stream) {
  ISP_REQUIRE(lr && gridr && gridc && freqs5 && bias_sin && bias_cos && out, ISP_ERR_BAD_SHAPE,
              "loftup_lr_prepare: null pointer");
  ISP_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, ISP_ERR_BAD_SHAPE, "loftup_lr_prepare: bad shape");
  const long long total = (long long)B * h * w;
  lr_prepare_kernel<<<cdiv(total, 8), 256, 0, as_stream(stream)>>>(lr, sb, sc, sh, sw, cn_w, cn_b, cn_w != nullptr, gridr,
                                                                 gridc, freqs5, bias_sin, bias_cos, out, B, C, h, w, eps);
  ISP_CHECK_LAUNCH("lr_prepare_kernel");
  return ISP_OK;
}

extern "C" int isp_repack_heads(const void* src, int src_bf16, long long ld, int col0, int head_dim, void* dst_bf16,
                                int B, int T, int Tpad, int heads, int D, int transpose, isp_stream_t stream) {
  ISP_REQUIRE(src && dst_bf16, ISP_ERR_BAD_SHAPE, "repack_heads: null pointer");
  ISP_REQUIRE(B > 0 && T > 0 && Tpad >= T && heads > 0 && D >= head_dim && head_dim > 0, ISP_ERR_BAD_SHAPE,
              "repack_heads: bad shape");
  ISP_REQUIRE(B <= 65535 && heads <= 65535 && head_dim <= 512, ISP_ERR_UNSUPPORTED, "repack_heads: grid too large");
  const dim3 grid((unsigned)cdiv(Tpad, kRepackT), (unsigned)heads, (unsigned)B);
  const size_t smem = (size_t)kRepackT * (head_dim + 2 - (head_dim & 1)) * sizeof(__nv_bfloat16);
  repack_heads_kernel<<<grid, 256, smem, as_stream(stream)>>>(src, src_bf16, ld, col0, head_dim,
                                                              reinterpret_cast<__nv_bfloat16*>(dst_bf16), B, T, Tpad,
                                                              heads, D, transpose);
  ISP_CHECK_LAUNCH("repack_heads_kernel");
  return ISP_OK;
}

extern "C" int isp_vit_patchify(const float* img, long long sb, long long sc, long long sh, long long sw, void* out_bf16,
                                int B, int Cin, int H, int W, int P, int ldo, isp_stream_t stream) {
  ISP_REQUIRE(img && out_bf16, ISP_ERR_BAD_SHAPE, "vit_patchify: null pointer");
  ISP_REQUIRE(B > 0 && Cin > 0 && P > 0 && H >= P && W >= P && ldo >= Cin * P * P, ISP_ERR_BAD_SHAPE,
              "vit_patchify: bad shape");
  const long long total = (long long)B * (H / P) * (W / P) * ldo;
  patchify_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(img, sb, sc, sh, sw,
                                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16), B, Cin, H, W,
                                                                 P, ldo);
  ISP_CHECK_LAUNCH("patchify_kernel");
  return ISP_OK;
}

extern "C" int isp_vit_assemble_tokens(const float* patch, const float* extra, const float* cls, const float* pos,
                                       float* out, int B, int N, int C, isp_stream_t stream) {
  ISP_REQUIRE(patch && cls && pos && out && B > 0 && N > 0 && C > 0, ISP_ERR_BAD_SHAPE, "vit_assemble_tokens: bad arguments");
  const long long total = (long long)B * (N + 1) * C;
  vit_tokens_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(patch, extra, cls, pos, out, B, N, C);
  ISP_CHECK_LAUNCH("vit_tokens_kernel");
  return ISP_OK;
}

extern "C" int isp_rowdot(const void* x, int x_bf16, long long ld, const float* w, const float* b, float* out,
                          long long M, int C, int K, isp_stream_t stream) {
  ISP_REQUIRE(x && w && out, ISP_ERR_BAD_SHAPE, "rowdot: null pointer");
  ISP_REQUIRE(M > 0 && C > 0 && K > 0 && K <= 8 && ld >= C, ISP_ERR_BAD_SHAPE, "rowdot: bad shape (K <= 8)");
  if (K == 1 && x_bf16 && C % 8 == 0 && C <= 512 && ld % 8 == 0 && aligned16(x)) {
    int num_sms = 0;
    if (int e = device_sm_count(&num_sms)) return e;
    const long long want = cdiv(M, 32);  // 8 warps x 4 rows per block and pass
    const int grid = (int)(want < (long long)num_sms * 8 ? want : (long long)num_sms * 8);
    rowdot1_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, w, b, out, M, C);
    ISP_CHECK_LAUNCH("rowdot1_bf16_kernel");
    return ISP_OK;
  }
  rowdot_kernel<<<cdiv(M, 8), 256, 0, as_stream(stream)>>>(x, x_bf16, ld, w, b, out, M, C, K);
  ISP_CHECK_LAUNCH("rowdot_kernel");
  return ISP_OK;
}
