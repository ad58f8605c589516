// AdaptiveConv (FeatUp's per-pixel 7x7 filter; a12 of SURVEY.md section 8a).
//   out[b,y,x,c] = sum_{i,j<7} in[b,y+i,x+j,c] * filt[b,y,x,i*7+j]
//
// Balanced kernel: 49 FMA per 8 output bytes + 8 input bytes (11.4 flop/B at op
// level), so both HBM and the fp32 pipe sit near their roofs.
// Design (NHWC fast path):
//   * lanes = channel PAIRS, so every FMA is a packed FFMA2 (__ffma2_rn) on (c, c+1);
//   * the 49 per-pixel weights are the same for all lanes -> broadcast float4 reads
//     from shared memory (one wavefront feeds 4 taps x 64 channels);
//   * each warp sweeps one output row of the tile keeping a 7x7 float2 register
//     window, so each step loads only the 7 new inputs (a column) from smem;
//   * the input tile ((8+6) x (16+6) pixels x 64 channels) is double-buffered with
//     cp.async across the channel groups of the CTA, filters staged once per tile.
// HBM sees each input/output byte about once (halo re-reads are L2 hits).
#include "common.cuh"

namespace isp {

constexpr int AC_TH = 8, AC_TW = 16, AC_CG = 64;
constexpr int AC_PH = AC_TH + 6, AC_PW = AC_TW + 6, AC_NPIX = AC_PH * AC_PW;
constexpr int AC_WP = 52;  // padded taps per pixel in smem (13 float4)
constexpr int AC_THREADS = 32 * AC_TH;
constexpr size_t AC_SMEM = (size_t)AC_TH * AC_TW * AC_WP * 4 + 2ull * AC_NPIX * AC_CG * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ac_fill(float* dst, const float* __restrict__ in, int b, int Hp, int Wp, int C,
                                        int ty0, int tx0, int cg) {
  for (int idx = threadIdx.x; idx < AC_NPIX * (AC_CG / 4); idx += AC_THREADS) {
    const int pix = idx >> 4, chunk = idx & 15;
    const int py = pix / AC_PW, px = pix - py * AC_PW;
    const int gy = min(ty0 + py, Hp - 1), gx = min(tx0 + px, Wp - 1);  // overhang: clamp (never stored)
    const float* src = in + (((size_t)b * Hp + gy) * Wp + gx) * C + cg * AC_CG + chunk * 4;
    cp_async16(dst + pix * AC_CG + chunk * 4, src);
  }
}

__global__ void __launch_bounds__(AC_THREADS, 1)
adaptive_conv_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ filt, float* __restrict__ out,
                          int H, int W, int C) {
  extern __shared__ __align__(16) float ac_smem[];
  float* w_s = ac_smem;                                   // [TH*TW][52]
  float* in_s0 = ac_smem + AC_TH * AC_TW * AC_WP;         // [2][NPIX][64]
  const int b = blockIdx.z, ty0 = blockIdx.y * AC_TH, tx0 = blockIdx.x * AC_TW;
  const int Hp = H + 6, Wp = W + 6, ncg = C / AC_CG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  ac_fill(in_s0, in, b, Hp, Wp, C, ty0, tx0, 0);
  cp_async_commit();
  for (int idx = threadIdx.x; idx < AC_TH * AC_TW * AC_WP; idx += AC_THREADS) {
    const int pix = idx / AC_WP, t = idx - pix * AC_WP;
    const int y = ty0 + pix / AC_TW, x = tx0 + pix % AC_TW;
    w_s[idx] = (t < 49 && y < H && x < W) ? __ldg(filt + (((size_t)b * H + y) * W + x) * 49 + t) : 0.f;
  }

  const int y = ty0 + warp;
  for (int cg = 0; cg < ncg; ++cg) {
    float* cur = in_s0 + (size_t)(cg & 1) * AC_NPIX * AC_CG;
    if (cg + 1 < ncg) {
      ac_fill(in_s0 + (size_t)((cg + 1) & 1) * AC_NPIX * AC_CG, in, b, Hp, Wp, C, ty0, tx0, cg + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const float2* tile = reinterpret_cast<const float2*>(cur) + lane;  // + pix*32 per pixel
    float2 win[7][7];
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int i = 0; i < 7; ++i) win[i][j] = tile[((warp + i) * AC_PW + j) * 32];
    float* orow = out + (((size_t)b * H + y) * W + tx0) * C + cg * AC_CG + 2 * lane;
#pragma unroll
    for (int x = 0; x < AC_TW; ++x) {
#pragma unroll
      for (int i = 0; i < 7; ++i) win[i][(x + 6) % 7] = tile[((warp + i) * AC_PW + x + 6) * 32];
      const float4* wq = reinterpret_cast<const float4*>(w_s + (warp * AC_TW + x) * AC_WP);
      float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      float wv[AC_WP];
#pragma unroll
      for (int q = 0; q < 13; ++q) {
        const float4 w4 = wq[q];
        wv[q * 4 + 0] = w4.x; wv[q * 4 + 1] = w4.y; wv[q * 4 + 2] = w4.z; wv[q * 4 + 3] = w4.w;
      }
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const int t = i * 7 + j;
          acc[t & 3] = __ffma2_rn(win[i][(x + j) % 7], make_float2(wv[t], wv[t]), acc[t & 3]);
        }
      const float2 r = make_float2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x),
                                   (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
      if (y < H && tx0 + x < W) *reinterpret_cast<float2*>(orow + (size_t)x * C) = r;
    }
    __syncthreads();  // everyone done with `cur` before it is refilled two iterations later
  }
}

// FeatUp-layout (NCHW) version for any C: thread = one output element; 49 strided taps
// through L1/L2.  Kept for interface parity with AdaptiveConv.apply and as an
// independent cross-check of the tiled kernel; not on the fast path.
__global__ void __launch_bounds__(256) adaptive_conv_nchw_kernel(const float* __restrict__ in,
                                                                 const float* __restrict__ filt,
                                                                 float* __restrict__ out, int B, int H, int W, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const int c = (int)((idx / ((long long)W * H)) % C), b = (int)(idx / ((long long)W * H * C));
  const int Wp = W + 6;
  const float* ip = in + (((size_t)b * C + c) * (H + 6) + y) * Wp + x;
  const float* fp = filt + (((size_t)b * H + y) * W + x) * 49;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc = fmaf(__ldg(ip + i * Wp + j), __ldg(fp + i * 7 + j), acc);
  out[idx] = acc;
}

// grad wrt the padded input, NHWC:  gi[b,Y,X,c] = sum_{i,j} go[b,Y-i,X-j,c] * f[b,Y-i,X-j,i*7+j]
// thread = (padded pixel, 4 channels); gathers <= 49 (pixel, tap) pairs.
__global__ void __launch_bounds__(256) adaptive_conv_grad_input_kernel(const float* __restrict__ go,
                                                                       const float* __restrict__ filt,
                                                                       float* __restrict__ gi, int B, int H, int W,
                                                                       int C) {
  const int C4 = C / 4, Hp = H + 6, Wp = W + 6;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * Hp * Wp * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const long long p = idx / C4;
  const int X = (int)(p % Wp), Y = (int)((p / Wp) % Hp), b = (int)(p / ((long long)Wp * Hp));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < 7; ++i) {
    const int yy = Y - i;
    if (yy < 0 || yy >= H) continue;
    for (int j = 0; j < 7; ++j) {
      const int xx = X - j;
      if (xx < 0 || xx >= W) continue;
      const size_t pix = ((size_t)b * H + yy) * W + xx;
      const float f = __ldg(filt + pix * 49 + i * 7 + j);
      const float4 g = __ldg(reinterpret_cast<const float4*>(go) + pix * C4 + c4);
      acc.x = fmaf(g.x, f, acc.x);
      acc.y = fmaf(g.y, f, acc.y);
      acc.z = fmaf(g.z, f, acc.z);
      acc.w = fmaf(g.w, f, acc.w);
    }
  }
  reinterpret_cast<float4*>(gi)[idx] = acc;
}

}  // namespace isp

using namespace isp;

static int ac_check(const void* a, const void* f, const void* o, int B, int H, int W, int C, const char* name) {
  ISP_REQUIRE(a && f && o, ISP_ERR_BAD_SHAPE, "%s: null pointer", name);
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, ISP_ERR_BAD_SHAPE, "%s: bad shape B=%d H=%d W=%d C=%d", name, B, H, W, C);
  ISP_REQUIRE(B <= 65535, ISP_ERR_UNSUPPORTED, "%s: B=%d exceeds grid.z", name, B);
  return ISP_OK;
}

extern "C" int isp_adaptive_conv_fwd_v1(const float* in_padded, const float* filters, float* out, int B, int H, int W,
                                     int C, isp_stream_t stream) {
  if (int e = ac_check(in_padded, filters, out, B, H, W, C, "adaptive_conv_fwd_v1")) return e;
  ISP_REQUIRE(C % AC_CG == 0, ISP_ERR_UNSUPPORTED, "adaptive_conv_fwd: NHWC path needs C %% 64 == 0 (C=%d)", C);
  ISP_REQUIRE(aligned16(in_padded) && aligned16(out), ISP_ERR_MISALIGNED, "adaptive_conv_fwd: 16-byte alignment");
  static bool attr_set = false;
  if (!attr_set) {
    ISP_CUDA(cudaFuncSetAttribute(adaptive_conv_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AC_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(W, AC_TW), cdiv(H, AC_TH), B);
  ISP_REQUIRE(grid.y <= 65535, ISP_ERR_UNSUPPORTED, "adaptive_conv_fwd: H too large");
  adaptive_conv_nhwc_kernel<<<grid, AC_THREADS, AC_SMEM, as_stream(stream)>>>(in_padded, filters, out, H, W, C);
  ISP_CHECK_LAUNCH("adaptive_conv_nhwc_kernel");
  return ISP_OK;
}

extern "C" int isp_adaptive_conv_fwd_nchw(const float* in_padded, const float* filters, float* out, int B, int H, int W,
                                          int C, isp_stream_t stream) {
  if (int e = ac_check(in_padded, filters, out, B, H, W, C, "adaptive_conv_fwd_nchw")) return e;
  const long long total = (long long)B * C * H * W;
  adaptive_conv_nchw_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(in_padded, filters, out, B, H, W, C);
  ISP_CHECK_LAUNCH("adaptive_conv_nchw_kernel");
  return ISP_OK;
}

extern "C" int isp_adaptive_conv_grad_input(const float* grad_out, const float* filters, float* grad_in, int B, int H,
                                            int W, int C, isp_stream_t stream) {
  if (int e = ac_check(grad_out, filters, grad_in, B, H, W, C, "adaptive_conv_grad_input")) return e;
  ISP_REQUIRE(C % 4 == 0, ISP_ERR_UNSUPPORTED, "adaptive_conv_grad_input: C %% 4 == 0 required (C=%d)", C);
  ISP_REQUIRE(aligned16(grad_out) && aligned16(grad_in), ISP_ERR_MISALIGNED, "adaptive_conv_grad_input: alignment");
  const long long total = (long long)B * (H + 6) * (W + 6) * (C / 4);
  adaptive_conv_grad_input_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(grad_out, filters, grad_in, B, H, W, C);
  ISP_CHECK_LAUNCH("adaptive_conv_grad_input_kernel");
  return ISP_OK;
}
