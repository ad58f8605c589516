// JBU combined-kernel producer, second generation (isp_jbu_filters dispatches here).
//   filters[b,y,x,:] = k + 0.1 * fixup([k, g]),  k = renorm(softmax_49(temp <proj nbr, proj>) * spatial)
// (upstream featup/upsamplers.py JBULearnedRange.forward; restated in oracle/jbu.py.)
//
// v1 (jbu.cu) read every neighbour's 32-float projection straight from global memory with
// one 16-byte load per lane and 32 different cache lines per warp instruction (7.9 ms for
// B=16 at 512^2) and kept everything of a pixel in one thread.  v2 is two kernels:
//   A  jbu_range_kernel : 4 x 32 pixel tile per block; the projection halo tile (10 x 38 pixels x
//      32) is staged in shared memory k-pair-major ([16][pix] float2, odd pitch) so a warp reads
//      32 consecutive pixels conflict-free with LDS.64 and the 32-long dot product is 16 packed
//      FFMA2; softmax / spatial / renormalise in registers; k is written to `filters` coalesced
//      (dense [.,49] or row-padded [.,7,8], the layout AdaptiveConv fetches with one TMA box).
//   B  jbu_fixup_kernel : 256 consecutive pixels per block, TWO pixels per thread: the
//      52->49->49 MLP in outer-product form on the (pixel A, pixel B) pair, one broadcast weight
//      (scalar FFMA2 operand) feeding both pixels; result accumulated onto k in place
//      (filters += 0.1 * o, coalesced read-modify-write that stays in L2).
// Splitting keeps both kernels far below the register limit (v2 as one kernel spilled) and lets
// A run at 4 blocks/SM and B at 3 blocks/SM.
#include "common.cuh"

namespace isp {
namespace jf2 {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// slot of the padded [7][8] layout -> tap index (or -1 for the zero 8th entry)
__device__ __forceinline__ int slot_to_tap(int slot, int out_ld) {
  if (out_ld == 49) return slot;
  return (slot & 7) < 7 ? (slot >> 3) * 7 + (slot & 7) : -1;
}

// ------------------------------------------------------------------ kernel A
constexpr int A_TH = 4, A_TW = 32, A_THREADS = 128;
constexpr int A_PH = A_TH + 6, A_PW = A_TW + 6, A_NPIX = A_PH * A_PW;  // 10 x 38 = 380
constexpr int A_PITCH = A_NPIX + 1;                                    // 381 (odd)
constexpr int A_KS = A_THREADS + 1;                                    // pitch of the k staging rows

struct SmemA {
  union {
    float2 proj[16 * A_PITCH];  // 48768 B
    float k[49 * A_KS];         // 25284 B
  } u;
  float spatial[52];
};

template <int OUT_LD>  // 49 (dense) or 56 (row-padded): compile-time, so the write-out's index arithmetic has no divisions
__global__ void __launch_bounds__(A_THREADS, 4)
jbu_range_kernel(const float* __restrict__ proj, float* __restrict__ filters, int H, int W, float temp, float inv2s2) {
  constexpr int out_ld = OUT_LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemA& S = *reinterpret_cast<SmemA*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, ty0 = blockIdx.y * A_TH, tx0 = blockIdx.x * A_TW;
  if (tid < 52) {
    float sp = 0.f;
    if (tid < 49) {  // linspace(-1,1,7): -1 + k/3
      const float dy = -1.f + (float)(tid / 7) * (1.f / 3.f), dx = -1.f + (float)(tid % 7) * (1.f / 3.f);
      sp = expf(-(dy * dy + dx * dx) * inv2s2);
    }
    S.spatial[tid] = sp;
  }
  {  // projection halo tile -> smem, transposed to k-pair-major (reflect padding = index reflection)
    const float4* p4 = reinterpret_cast<const float4*>(proj) + (size_t)b * H * W * 8;
    for (int idx = tid; idx < A_NPIX * 8; idx += A_THREADS) {
      const int pix = idx >> 3, k4 = idx & 7;
      const int py = pix / A_PW, px = pix - py * A_PW;
      const int gy = clampi(reflect_idx(ty0 + py - 3, H), 0, H - 1);
      const int gx = clampi(reflect_idx(tx0 + px - 3, W), 0, W - 1);
      const float4 v = __ldg(p4 + ((size_t)gy * W + gx) * 8 + k4);
      S.u.proj[(2 * k4) * A_PITCH + pix] = make_float2(v.x, v.y);
      S.u.proj[(2 * k4 + 1) * A_PITCH + pix] = make_float2(v.z, v.w);
    }
  }
  __syncthreads();
  float k[49];
  {
    const float2* ctrp = S.u.proj + (warp + 3) * A_PW + lane + 3;
    float2 ctr[16];
#pragma unroll
    for (int kp = 0; kp < 16; ++kp) ctr[kp] = ctrp[kp * A_PITCH];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float2* np = S.u.proj + (warp + i) * A_PW + lane + j;
        float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int kp = 0; kp < 16; ++kp) d2 = __ffma2_rn(np[kp * A_PITCH], ctr[kp], d2);
        const float d = (d2.x + d2.y) * temp;
        k[i * 7 + j] = d;
        mx = fmaxf(mx, d);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 49; ++t) { k[t] = __expf(k[t] - mx); sum += k[t]; }  // MUFU.EX2: 2 instructions instead of expf's ~10
    const float inv = 1.f / sum;
    float sum2 = 0.f;
#pragma unroll
    for (int t = 0; t < 49; ++t) { k[t] = (k[t] * inv) * S.spatial[t]; sum2 += k[t]; }
    const float inv2 = 1.f / fmaxf(sum2, 1e-7f);  // one division per pixel instead of 49
#pragma unroll
    for (int t = 0; t < 49; ++t) k[t] = k[t] * inv2;
  }
  __syncthreads();  // projection tile dead; its storage becomes the k staging buffer [tap][pixel]
#pragma unroll
  for (int t = 0; t < 49; ++t) S.u.k[t * A_KS + tid] = k[t];
  __syncthreads();
  for (int rr = 0; rr < A_TH; ++rr) {  // coalesced: one tile row = 32 pixels x out_ld contiguous floats
    const int y = ty0 + rr;
    if (y >= H) break;
    const int npx = min(A_TW, W - tx0);
    float* dst = filters + (((size_t)b * H + y) * W + tx0) * out_ld;
    if (OUT_LD == 56) {  // 14 float4 per pixel: taps 7i .. 7i+3, then 7i+4 .. 7i+6 and the zero pad
      float4* dst4 = reinterpret_cast<float4*>(dst);
      for (int e4 = tid; e4 < npx * 14; e4 += A_THREADS) {
        const int px = e4 / 14, s14 = e4 - px * 14;
        const float* kp = &S.u.k[((s14 >> 1) * 7 + (s14 & 1) * 4) * A_KS + rr * 32 + px];
        dst4[e4] = make_float4(kp[0], kp[A_KS], kp[2 * A_KS], (s14 & 1) ? 0.f : kp[3 * A_KS]);
      }
    } else {
      for (int e = tid; e < npx * out_ld; e += A_THREADS) {
        const int px = e / out_ld, tap = e - px * out_ld;
        dst[e] = S.u.k[tap * A_KS + rr * 32 + px];
      }
    }
  }
}

// ------------------------------------------------------------------ kernel B
constexpr int B_THREADS = 128, B_PIX = 2 * B_THREADS;
constexpr int B_VS = B_PIX + 2;  // even pitch: float2 column pairs stay 8-byte aligned
constexpr int kWP = 52;

struct SmemB {
  float v[52 * B_VS];   // [c][pixel]: k (49) + g (3); then hidden; then output
  float w0t[52 * kWP];  // fixup_proj.0 transposed [c][r]
  float w1t[49 * kWP];  // fixup_proj.3 transposed [c][r]
  float b0[kWP], b1[kWP];
};

__device__ __forceinline__ void mlp_layer(float2 (&acc)[kWP], const float* vcol, const float* wt, const float* bias,
                                          int ncin) {
#pragma unroll
  for (int r = 0; r < kWP; ++r) acc[r] = make_float2(bias[r], bias[r]);
  for (int c = 0; c < ncin; ++c) {
    const float2 vc = *reinterpret_cast<const float2*>(vcol + c * B_VS);
    const float4* wr = reinterpret_cast<const float4*>(wt + c * kWP);
#pragma unroll
    for (int q = 0; q < 13; ++q) {
      const float4 w = wr[q];
      acc[q * 4 + 0] = __ffma2_rn(vc, make_float2(w.x, w.x), acc[q * 4 + 0]);
      acc[q * 4 + 1] = __ffma2_rn(vc, make_float2(w.y, w.y), acc[q * 4 + 1]);
      acc[q * 4 + 2] = __ffma2_rn(vc, make_float2(w.z, w.z), acc[q * 4 + 2]);
      acc[q * 4 + 3] = __ffma2_rn(vc, make_float2(w.w, w.w), acc[q * 4 + 3]);
    }
  }
}

// k of this block's pixels, global <-> smem [tap][pixel].  The padded layout (56 floats = 14 float4 per pixel,
// slots 7, 15, ... are the zero 8th entries) moves as float4 with 7 independent loads in flight per thread;
// the dense layout (49) moves as scalars.  All index arithmetic is by compile-time constants.
template <int OUT_LD, bool STORE>
__device__ __forceinline__ void slab_rw(float* __restrict__ slab, float* v, int nlive, int tid) {
  if (OUT_LD == 56) {
    float4* s4 = reinterpret_cast<float4*>(slab);
    constexpr int N4 = B_PIX * 14, PER = N4 / B_THREADS;  // 28 float4 per thread
    static_assert(N4 % B_THREADS == 0 && PER % 7 == 0, "slab_rw tiling");
#pragma unroll 1
    for (int it = 0; it < PER; it += 7) {
      float4 x[7];
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int e4 = (it + u) * B_THREADS + tid, px = e4 / 14;
        x[u] = px < nlive ? s4[e4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int e4 = (it + u) * B_THREADS + tid, px = e4 / 14, s14 = e4 - px * 14;
        const int tap = (s14 >> 1) * 7 + (s14 & 1) * 4;
        float* vp = v + tap * B_VS + px;
        if (!STORE) {
          vp[0] = x[u].x; vp[B_VS] = x[u].y; vp[2 * B_VS] = x[u].z;
          if (!(s14 & 1)) vp[3 * B_VS] = x[u].w;
        } else if (px < nlive) {
          x[u].x = fmaf(0.1f, vp[0], x[u].x); x[u].y = fmaf(0.1f, vp[B_VS], x[u].y);
          x[u].z = fmaf(0.1f, vp[2 * B_VS], x[u].z);
          if (!(s14 & 1)) x[u].w = fmaf(0.1f, vp[3 * B_VS], x[u].w);
          s4[e4] = x[u];
        }
      }
    }
  } else {
#pragma unroll 4
    for (int e = tid; e < B_PIX * 49; e += B_THREADS) {
      const int px = e / 49, tap = e - px * 49;
      if (!STORE) v[tap * B_VS + px] = px < nlive ? slab[e] : 0.f;
      else if (px < nlive) slab[e] = fmaf(0.1f, v[tap * B_VS + px], slab[e]);
    }
  }
}

template <int OUT_LD>
__global__ void __launch_bounds__(B_THREADS, 3)
jbu_fixup_kernel(const float4* __restrict__ g, float* __restrict__ filters, long long npix,
                 const float* __restrict__ fw0, const float* __restrict__ fb0, const float* __restrict__ fw1,
                 const float* __restrict__ fb1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemB& S = *reinterpret_cast<SmemB*>(smem_raw);
  const int tid = threadIdx.x;
  const long long pix0 = (long long)blockIdx.x * B_PIX;
  const int nlive = (int)min((long long)B_PIX, npix - pix0);
  float* slab = filters + pix0 * OUT_LD;
  // k (from kernel A) and g -> smem [c][pixel]
  slab_rw<OUT_LD, false>(slab, S.v, nlive, tid);
  for (int px = tid; px < B_PIX; px += B_THREADS) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < nlive) t = g[pix0 + px];
    S.v[49 * B_VS + px] = t.x;
    S.v[50 * B_VS + px] = t.y;
    S.v[51 * B_VS + px] = t.z;
  }
  // weights: coalesced global reads, transposed on the way into smem ([c][r], rows 49..51 of r are zero)
  for (int i = tid; i < 52 * kWP; i += B_THREADS) S.w0t[i] = 0.f;
  for (int i = tid; i < 49 * kWP; i += B_THREADS) S.w1t[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < 49 * 52; i += B_THREADS) {
    const int r = i / 52, c = i - r * 52;
    S.w0t[c * kWP + r] = __ldg(fw0 + i);
  }
  for (int i = tid; i < 49 * 49; i += B_THREADS) {
    const int r = i / 49, c = i - r * 49;
    S.w1t[c * kWP + r] = __ldg(fw1 + i);
  }
  for (int i = tid; i < kWP; i += B_THREADS) {
    S.b0[i] = (i < 49) ? fb0[i] : 0.f;
    S.b1[i] = (i < 49) ? fb1[i] : 0.f;
  }
  __syncthreads();
  float* vcol = S.v + 2 * tid;  // this thread's pixel pair
  float2 acc[kWP];
  mlp_layer(acc, vcol, S.w0t, S.b0, 52);  // h = GELU(W0 [k;g] + b0)
#pragma unroll
  for (int r = 0; r < 49; ++r)
    *reinterpret_cast<float2*>(vcol + r * B_VS) = make_float2(gelu_erf(acc[r].x), gelu_erf(acc[r].y));
  mlp_layer(acc, vcol, S.w1t, S.b1, 49);  // o = W1 h + b1
#pragma unroll
  for (int t = 0; t < 49; ++t) *reinterpret_cast<float2*>(vcol + t * B_VS) = acc[t];
  __syncthreads();
  slab_rw<OUT_LD, true>(slab, S.v, nlive, tid);  // filters = k + 0.1 * o
}

}  // namespace jf2
}  // namespace isp

namespace isp {
int jbu_fixup_tc_launch(const float* g, float* filters, long long npix, const float* fw0, const float* fb0, const float* fw1,
                        const float* fb1, cudaStream_t stream);  // jbu_fixup_tc.cu
}
using namespace isp;

// filters layout: out_ld == 49 -> dense [B,H,W,49]; out_ld == 56 -> row-padded [B,H,W,7,8]
// simt != 0: the fix-up MLP on the fp32 pipe (jbu_fixup_kernel) also for the padded layout, which otherwise runs on the
// tensor cores (jbu_fixup_tc.cu)
static int jbu_filters_impl(const float* proj, const float* g, float* filters, int B, int H, int W, float temp,
                            float sigma_spatial, const float* fw0, const float* fb0, const float* fw1,
                            const float* fb1, int out_ld, int simt, isp_stream_t stream) {
  ISP_REQUIRE(proj && g && filters && fw0 && fb0 && fw1 && fb1, ISP_ERR_BAD_SHAPE, "jbu_filters: null pointer");
  ISP_REQUIRE(B > 0 && H >= 4 && W >= 4, ISP_ERR_BAD_SHAPE, "jbu_filters: need H,W >= 4 (reflect pad 3), got %dx%d", H, W);
  ISP_REQUIRE(out_ld == 49 || out_ld == 56, ISP_ERR_BAD_SHAPE, "jbu_filters: out_ld must be 49 or 56 (got %d)", out_ld);
  ISP_REQUIRE(aligned16(proj) && aligned16(g), ISP_ERR_MISALIGNED, "jbu_filters: pointers must be 16-byte aligned");
  ISP_REQUIRE(out_ld != 56 || aligned16(filters), ISP_ERR_MISALIGNED, "jbu_filters: padded filters must be 16-byte aligned");
  ISP_REQUIRE(B <= 65535 && cdiv(H, jf2::A_TH) <= 65535, ISP_ERR_UNSUPPORTED, "jbu_filters: grid too large");
  const int smemA = (int)sizeof(jf2::SmemA), smemB = (int)sizeof(jf2::SmemB);
  if (int e = ensure_dynamic_smem((const void*)jf2::jbu_range_kernel<49>, smemA)) return e;
  if (int e = ensure_dynamic_smem((const void*)jf2::jbu_range_kernel<56>, smemA)) return e;
  if (int e = ensure_dynamic_smem((const void*)jf2::jbu_fixup_kernel<49>, smemB)) return e;
  if (int e = ensure_dynamic_smem((const void*)jf2::jbu_fixup_kernel<56>, smemB)) return e;
  const float inv2s2 = 1.f / (2.f * sigma_spatial * sigma_spatial);
  dim3 gridA(cdiv(W, jf2::A_TW), cdiv(H, jf2::A_TH), B);
  if (out_ld == 56)
    jf2::jbu_range_kernel<56><<<gridA, jf2::A_THREADS, smemA, as_stream(stream)>>>(proj, filters, H, W, temp, inv2s2);
  else
    jf2::jbu_range_kernel<49><<<gridA, jf2::A_THREADS, smemA, as_stream(stream)>>>(proj, filters, H, W, temp, inv2s2);
  ISP_CHECK_LAUNCH("jbu_range_kernel");
  const long long npix = (long long)B * H * W;
  if (out_ld == 56 && !simt) {
    ISP_REQUIRE(aligned16(filters), ISP_ERR_MISALIGNED, "jbu_filters: padded filters must be 16-byte aligned");
    return jbu_fixup_tc_launch(g, filters, npix, fw0, fb0, fw1, fb1, as_stream(stream));
  }
  if (out_ld == 56) {
    ISP_REQUIRE(aligned16(filters), ISP_ERR_MISALIGNED, "jbu_filters: padded filters must be 16-byte aligned");
    jf2::jbu_fixup_kernel<56><<<cdiv(npix, jf2::B_PIX), jf2::B_THREADS, smemB, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(g), filters, npix, fw0, fb0, fw1, fb1);
  } else {
    jf2::jbu_fixup_kernel<49><<<cdiv(npix, jf2::B_PIX), jf2::B_THREADS, smemB, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(g), filters, npix, fw0, fb0, fw1, fb1);
  }
  ISP_CHECK_LAUNCH("jbu_fixup_kernel");
  return ISP_OK;
}

extern "C" int isp_jbu_filters(const float* proj, const float* g, float* filters, int B, int H, int W, float temp,
                               float sigma_spatial, const float* fw0, const float* fb0, const float* fw1,
                               const float* fb1, int out_ld, isp_stream_t stream) {
  return jbu_filters_impl(proj, g, filters, B, H, W, temp, sigma_spatial, fw0, fb0, fw1, fb1, out_ld, 0, stream);
}
// the same with the fix-up MLP on the fp32 pipe for every layout (cross-check of the tensor-core kernel, A/B timing)
extern "C" int isp_jbu_filters_simt(const float* proj, const float* g, float* filters, int B, int H, int W, float temp,
                                    float sigma_spatial, const float* fw0, const float* fb0, const float* fw1,
                                    const float* fb1, int out_ld, isp_stream_t stream) {
  return jbu_filters_impl(proj, g, filters, B, H, W, temp, sigma_spatial, fw0, fb0, fw1, fb1, out_ld, 1, stream);
}
