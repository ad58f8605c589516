// FeatUp JBU stage pieces around AdaptiveConv (a12 of SURVEY.md section 8a),
// all fp32, channels-last.  Algorithm: upstream featup/upsamplers.py
// (JBULearnedRange.forward / JBUStack.upsample), restated in oracle/jbu.py.
//
//   guidance --adaptive_avg_pool--> g[B,GH,GW,4]
//   g --range_proj (3->32->32)-----> proj[B,GH,GW,32]
//   proj,g --49-tap range softmax * spatial, renorm, +0.1*fixup MLP--> filters[B,GH,GW,49]
//   source[B,h,w,C] --bicubic x2 + reflect pad 3--> hr_pad[B,GH+6,GW+6,C]
//   (hr_pad, filters) --adaptive_conv.cu--> out[B,GH,GW,C]
#include "common.cuh"

namespace isp {

// ---------------------------------------------------------------------------
// adaptive_avg_pool2d of the 3-channel guidance image, NCHW (strided) -> NHWC4.
// ATen window: [floor(o*I/O), ceil((o+1)*I/O)); sum rows-then-cols in fp32, divide by count.
__global__ void __launch_bounds__(256) pool_guidance_kernel(const float* __restrict__ gd, float4* __restrict__ out,
                                                            int B, int H, int W, int OH, int OW, long long sb,
                                                            long long sc, long long sh, long long sw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW;
  if (idx >= total) return;
  const int ox = (int)(idx % OW);
  const int oy = (int)((idx / OW) % OH);
  const int b = (int)(idx / ((long long)OW * OH));
  const int y0 = (int)(((long long)oy * H) / OH), y1 = (int)((((long long)oy + 1) * H + OH - 1) / OH);
  const int x0 = (int)(((long long)ox * W) / OW), x1 = (int)((((long long)ox + 1) * W + OW - 1) / OW);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < 3; ++c) {
    const float* p = gd + b * sb + c * sc;
    float s = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) s += p[y * sh + x * sw];
    acc[c] = s / (float)((y1 - y0) * (x1 - x0));
  }
  out[idx] = make_float4(acc[0], acc[1], acc[2], 0.f);
}

// ---------------------------------------------------------------------------
// range_proj: per-pixel MLP 3 -> 32 (GELU) -> 32.  A thread owns TWO pixels (lane and lane + 32 of its warp's 64), so every
// broadcast weight load feeds a packed FFMA2; weights in smem.  The 32 outputs of a pixel are 128 contiguous bytes: written
// straight from the registers a warp store would touch 32 different lines, so each warp transposes its 64 x 32 block
// through shared memory and stores 512 contiguous bytes per instruction.
__global__ void __launch_bounds__(128) range_proj_kernel(const float4* __restrict__ g, float4* __restrict__ proj,
                                                         long long npix, const float* __restrict__ w0,
                                                         const float* __restrict__ b0, const float* __restrict__ w1,
                                                         const float* __restrict__ b1) {
  __shared__ float s_w0[32 * 3], s_b0[32], s_w1t[32 * 32], s_b1[32];  // s_w1t[k][o]
  __shared__ float s_out[4][64][33];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) s_w0[i] = w0[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) { s_b0[i] = b0[i]; s_b1[i] = b1[i]; }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_w1t[(i % 32) * 32 + i / 32] = w1[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wpix0 = ((long long)blockIdx.x * 4 + warp) * 64;  // first pixel of this warp
  const long long pa = wpix0 + lane, pb = pa + 32;
  const float4 ga = pa < npix ? g[pa] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 gb = pb < npix ? g[pb] : make_float4(0.f, 0.f, 0.f, 0.f);
  float2 o[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) o[j] = make_float2(s_b1[j], s_b1[j]);
#pragma unroll 4
  for (int k = 0; k < 32; ++k) {
    const float wx = s_w0[k * 3 + 0], wy = s_w0[k * 3 + 1], wz = s_w0[k * 3 + 2], bk = s_b0[k];
    const float ha = gelu_erf_fast(fmaf(wz, ga.z, fmaf(wy, ga.y, fmaf(wx, ga.x, bk))));
    const float hb = gelu_erf_fast(fmaf(wz, gb.z, fmaf(wy, gb.y, fmaf(wx, gb.x, bk))));
    const float2 h2 = make_float2(ha, hb);
    const float4* wr = reinterpret_cast<const float4*>(&s_w1t[k * 32]);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 w = wr[j4];
      o[j4 * 4 + 0] = __ffma2_rn(make_float2(w.x, w.x), h2, o[j4 * 4 + 0]);
      o[j4 * 4 + 1] = __ffma2_rn(make_float2(w.y, w.y), h2, o[j4 * 4 + 1]);
      o[j4 * 4 + 2] = __ffma2_rn(make_float2(w.z, w.z), h2, o[j4 * 4 + 2]);
      o[j4 * 4 + 3] = __ffma2_rn(make_float2(w.w, w.w), h2, o[j4 * 4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    s_out[warp][lane][j] = o[j].x;
    s_out[warp][lane + 32][j] = o[j].y;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) {  // 4 pixels x 8 float4 per instruction
    const int px = r * 4 + (lane >> 3), u = lane & 7;
    if (wpix0 + px < npix)
      proj[(wpix0 + px) * 8 + u] = make_float4(s_out[warp][px][4 * u], s_out[warp][px][4 * u + 1], s_out[warp][px][4 * u + 2],
                                               s_out[warp][px][4 * u + 3]);
  }
}

__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  float x = t + 1.f;
  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

__global__ void __launch_bounds__(256) bicubic2x_pad_kernel(const float* __restrict__ src, float* __restrict__ out,
                                                            int B, int h, int w, int C) {
  const int GH = 2 * h, GW = 2 * w, PH = GH + 6, PW = GW + 6, C4 = C / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * PH * PW * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const long long p = idx / C4;
  const int px = (int)(p % PW), py = (int)((p / PW) % PH), b = (int)(p / ((long long)PW * PH));
  const int oy = reflect_idx(py - 3, GH), ox = reflect_idx(px - 3, GW);
  const float sy = ((float)oy + 0.5f) * 0.5f - 0.5f, sx = ((float)ox + 0.5f) * 0.5f - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float wy[4], wx[4];
  cubic_coeffs(sy - fy, wy);
  cubic_coeffs(sx - fx, wx);
  const int iy = (int)fy, ix = (int)fx;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* s4 = reinterpret_cast<const float4*>(src) + (long long)b * h * w * C4 + c4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int yy = min(max(iy - 1 + i, 0), h - 1);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = min(max(ix - 1 + j, 0), w - 1);
      const float4 v = __ldg(s4 + ((long long)yy * w + xx) * C4);
      r.x = fmaf(v.x, wx[j], r.x);
      r.y = fmaf(v.y, wx[j], r.y);
      r.z = fmaf(v.z, wx[j], r.z);
      r.w = fmaf(v.w, wx[j], r.w);
    }
    acc.x = fmaf(r.x, wy[i], acc.x);
    acc.y = fmaf(r.y, wy[i], acc.y);
    acc.z = fmaf(r.z, wy[i], acc.z);
    acc.w = fmaf(r.w, wy[i], acc.w);
  }
  reinterpret_cast<float4*>(out)[idx] = acc;
}

// Interior of the bicubic x2 output: the thread marches down a strip of kMarchR low-res rows for its 4 output columns x 4 channels.
// Every low-res row is loaded once per strip (6 float4) and filtered horizontally once; the last five horizontal
// results stay in registers and produce two output rows per step (1.1 loads per output float4 instead of 3.75;
// ncu on the 2x4-blocked kernel: 44 % of the stalls were the loads).  Same summation order as the blocked kernel.
constexpr int kMarchR = 8;
__global__ void __launch_bounds__(128, 3) bicubic2x_march_kernel(const float* __restrict__ src, float* __restrict__ out,
                                                                 int B, int h, int w, int C) {
  const int C4 = C / 4, wp = (w + 1) / 2, nstrips = (h + kMarchR - 1) / kMarchR;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * nstrips * wp * C4;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long long p = idx / C4;
  const int lp = (int)(p % wp);
  p /= wp;
  const int ks = (int)(p % nstrips) * kMarchR, b = (int)(p / nstrips);
  const int l0 = 2 * lp;
  float cE[4], cO[4];
  cubic_coeffs(0.75f, cE);
  cubic_coeffs(0.25f, cO);
  const float4* s4 = reinterpret_cast<const float4*>(src) + (long long)b * h * w * C4 + c4;
  long long xoff[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) xoff[j] = (long long)min(max(l0 - 2 + j, 0), w - 1) * C4;
  const int PW = 2 * w + 6, PH = 2 * h + 6;
  float4* o4 = reinterpret_cast<float4*>(out) + (long long)b * PH * PW * C4 + c4;
  float4 hh[5][4];  // horizontal results of the five most recent rows (oldest first)
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = __ldg(s4 + (long long)min(max(ks - 2, 0), h - 1) * w * C4 + xoff[j]);
#pragma unroll
  for (int m = 0; m < kMarchR + 4; ++m) {  // input row ks - 2 + m
    float4 vn[6];
    if (m + 1 < kMarchR + 4) {             // next row in flight while this one is filtered
      const long long yoff = (long long)min(max(ks - 1 + m, 0), h - 1) * w * C4;
#pragma unroll
      for (int j = 0; j < 6; ++j) vn[j] = __ldg(s4 + yoff + xoff[j]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) hh[r][q] = hh[r + 1][q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float* cc = (q & 1) ? cO : cE;
      const int o0 = (q >> 1) + (q & 1);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a.x = fmaf(v[o0 + j].x, cc[j], a.x);
        a.y = fmaf(v[o0 + j].y, cc[j], a.y);
        a.z = fmaf(v[o0 + j].z, cc[j], a.z);
        a.w = fmaf(v[o0 + j].w, cc[j], a.w);
      }
      hh[4][q] = a;
    }
    if (m >= 4) {  // rows k-2 .. k+2 are in the window: k = ks + m - 4
      const int k = ks + m - 4;
      if (k < h) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float cw = a ? cO[r] : cE[r];
              const float4 t = hh[r + a][q];
              o.x = fmaf(t.x, cw, o.x); o.y = fmaf(t.y, cw, o.y); o.z = fmaf(t.z, cw, o.z); o.w = fmaf(t.w, cw, o.w);
            }
            const int ox = 2 * l0 + q;
            if (ox < 2 * w) o4[((long long)(2 * k + a + 3) * PW + ox + 3) * C4] = o;
          }
      }
    }
    if (m + 1 < kMarchR + 4) {
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] = vn[j];
    }
  }
}

__global__ void __launch_bounds__(256) reflect_border_kernel(float* __restrict__ out, int B, int GH, int GW, int C) {
  const int C4 = C / 4, PW = GW + 6, PH = GH + 6;
  const int nborder = 6 * PW + 6 * GH;  // 3 top + 3 bottom rows, then 3 left + 3 right columns of the middle rows
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * nborder * C4) return;
  const int c4 = (int)(idx % C4);
  long long p = idx / C4;
  const int e = (int)(p % nborder), b = (int)(p / nborder);
  int py, px;
  if (e < 6 * PW) {
    const int r = e / PW;
    px = e % PW;
    py = r < 3 ? r : GH + r;  // rows 0..2 and GH+3..GH+5
  } else {
    const int m = e - 6 * PW;
    py = 3 + m / 6;
    const int cidx = m % 6;
    px = cidx < 3 ? cidx : GW + cidx;  // cols 0..2 and GW+3..GW+5
  }
  const int sy = reflect_idx(py - 3, GH) + 3, sx = reflect_idx(px - 3, GW) + 3;
  float4* o4 = reinterpret_cast<float4*>(out) + (long long)b * PH * PW * C4 + c4;
  o4[((long long)py * PW + px) * C4] = o4[((long long)sy * PW + sx) * C4];
}

// Adjoint of isp_jbu_bicubic2x_reflectpad (gradient w.r.t. the low-res source of one JBU stage): every source pixel
// gathers from the up-sampled pixels it fed (bicubic A = -0.75, align_corners = False, border clamp), each of which
// collects the gradient of its interior position and of the reflect-pad copies of it.
__device__ __forceinline__ float bicubic_adj_weight(int Y, int m, int n, const float* cE, const float* cO) {
  // weight of source index m in up-sampled index Y (2n outputs): Y = 2k uses k-2..k+1 (cE), Y = 2k+1 uses k-1..k+2 (cO)
  const int k = Y >> 1;
  const float* c = (Y & 1) ? cO : cE;
  const int base = (Y & 1) ? k - 1 : k - 2;
  float wgt = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (min(max(base + i, 0), n - 1) == m) wgt += c[i];
  return wgt;
}
// sum of the padded-gradient entries that are copies of up-sampled index Y (size G, pad 3, reflect): positions
// Y + 3 (interior), 3 - Y (left frame, 1 <= Y <= 3), 2G + 1 - Y (right frame, G-4 <= Y <= G-2)
__device__ __forceinline__ int reflect_sources(int Y, int G, int (&pos)[3]) {
  int n = 0;
  pos[n++] = Y + 3;
  if (Y >= 1 && Y <= 3) pos[n++] = 3 - Y;
  if (Y >= G - 4 && Y <= G - 2) pos[n++] = 2 * G + 1 - Y;
  return n;
}
__global__ void __launch_bounds__(256) bicubic2x_pad_bwd_kernel(const float* __restrict__ gpad, float* __restrict__ gsrc,
                                                               int B, int h, int w, int C) {
  const int C4 = C / 4, GH = 2 * h, GW = 2 * w, PW = GW + 6, PH = GH + 6;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * h * w * C4) return;
  const int c4 = (int)(idx % C4);
  long long p = idx / C4;
  const int l = (int)(p % w);
  p /= w;
  const int m = (int)(p % h), b = (int)(p / h);
  float cE[4], cO[4];
  cubic_coeffs(0.75f, cE);
  cubic_coeffs(0.25f, cO);
  const float4* g4 = reinterpret_cast<const float4*>(gpad) + (long long)b * PH * PW * C4 + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int Y = max(0, 2 * m - 4); Y <= min(GH - 1, 2 * m + 5); ++Y) {
    const float wy = bicubic_adj_weight(Y, m, h, cE, cO);
    if (wy == 0.f) continue;
    int py[3];
    const int ny = reflect_sources(Y, GH, py);
    for (int X = max(0, 2 * l - 4); X <= min(GW - 1, 2 * l + 5); ++X) {
      const float wx = bicubic_adj_weight(X, l, w, cE, cO);
      if (wx == 0.f) continue;
      int px[3];
      const int nx = reflect_sources(X, GW, px);
      const float wgt = wy * wx;
      for (int a = 0; a < ny; ++a)
        for (int c = 0; c < nx; ++c) {
          const float4 t = __ldg(g4 + ((long long)py[a] * PW + px[c]) * C4);
          acc.x = fmaf(wgt, t.x, acc.x); acc.y = fmaf(wgt, t.y, acc.y);
          acc.z = fmaf(wgt, t.z, acc.z); acc.w = fmaf(wgt, t.w, acc.w);
        }
    }
  }
  reinterpret_cast<float4*>(gsrc)[idx] = acc;
}

}  // namespace isp

using namespace isp;

extern "C" int isp_jbu_pool_guidance(const float* guidance, float* out, int B, int H, int W, int OH, int OW,
                                     long long sb, long long sc, long long sh, long long sw, isp_stream_t stream) {
  ISP_REQUIRE(guidance && out, ISP_ERR_BAD_SHAPE, "jbu_pool_guidance: null pointer");
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, ISP_ERR_BAD_SHAPE, "jbu_pool_guidance: bad shape");
  ISP_REQUIRE(aligned16(out), ISP_ERR_MISALIGNED, "jbu_pool_guidance: out must be 16-byte aligned");
  const long long total = (long long)B * OH * OW;
  pool_guidance_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(guidance, reinterpret_cast<float4*>(out), B, H,
                                                                       W, OH, OW, sb, sc, sh, sw);
  ISP_CHECK_LAUNCH("pool_guidance_kernel");
  return ISP_OK;
}

extern "C" int isp_jbu_range_proj(const float* g, float* proj, long long npix, const float* w0, const float* b0,
                                  const float* w1, const float* b1, isp_stream_t stream) {
  ISP_REQUIRE(g && proj && w0 && b0 && w1 && b1, ISP_ERR_BAD_SHAPE, "jbu_range_proj: null pointer");
  ISP_REQUIRE(npix > 0, ISP_ERR_BAD_SHAPE, "jbu_range_proj: npix=%lld", npix);
  ISP_REQUIRE(aligned16(g) && aligned16(proj), ISP_ERR_MISALIGNED, "jbu_range_proj: pointers must be 16-byte aligned");
  range_proj_kernel<<<cdiv(npix, 256), 128, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(g),
                                                                    reinterpret_cast<float4*>(proj), npix, w0, b0, w1, b1);
  ISP_CHECK_LAUNCH("range_proj_kernel");
  return ISP_OK;
}

// Gradient of isp_jbu_bicubic2x_reflectpad w.r.t. src: gsrc [B,h,w,C] from gpad [B,2h+6,2w+6,C] (fp32 NHWC, C % 4 == 0).
extern "C" int isp_jbu_bicubic2x_reflectpad_bwd(const float* gpad, float* gsrc, int B, int h, int w, int C,
                                                isp_stream_t stream) {
  ISP_REQUIRE(gpad && gsrc && B > 0 && h >= 2 && w >= 2 && C > 0 && C % 4 == 0, ISP_ERR_BAD_SHAPE,
              "jbu_bicubic2x_reflectpad_bwd: need h,w >= 2 and C %% 4 == 0");
  ISP_REQUIRE(aligned16(gpad) && aligned16(gsrc), ISP_ERR_MISALIGNED, "jbu_bicubic2x_reflectpad_bwd: 16-byte alignment");
  const long long total = (long long)B * h * w * (C / 4);
  bicubic2x_pad_bwd_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(gpad, gsrc, B, h, w, C);
  ISP_CHECK_LAUNCH("bicubic2x_pad_bwd_kernel");
  return ISP_OK;
}

extern "C" int isp_jbu_bicubic2x_reflectpad(const float* src, float* out, int B, int h, int w, int C,
                                            isp_stream_t stream) {
  ISP_REQUIRE(src && out, ISP_ERR_BAD_SHAPE, "jbu_bicubic2x_reflectpad: null pointer");
  ISP_REQUIRE(B > 0 && h >= 2 && w >= 2 && C > 0 && C % 4 == 0, ISP_ERR_BAD_SHAPE,
              "jbu_bicubic2x_reflectpad: need h,w >= 2 and C %% 4 == 0 (h=%d w=%d C=%d)", h, w, C);
  ISP_REQUIRE(aligned16(src) && aligned16(out), ISP_ERR_MISALIGNED, "jbu_bicubic2x_reflectpad: 16-byte alignment");
  if (h >= 4 && w >= 4) {  // blocked interior + reflect frame (frame sources lie in the interior)
    const long long total = (long long)B * cdiv(h, kMarchR) * ((w + 1) / 2) * (C / 4);
    bicubic2x_march_kernel<<<cdiv(total, 128), 128, 0, as_stream(stream)>>>(src, out, B, h, w, C);
    ISP_CHECK_LAUNCH("bicubic2x_march_kernel");
    const long long nb = (long long)B * (6 * (2 * w + 6) + 6 * 2 * h) * (C / 4);
    reflect_border_kernel<<<cdiv(nb, 256), 256, 0, as_stream(stream)>>>(out, B, 2 * h, 2 * w, C);
    ISP_CHECK_LAUNCH("reflect_border_kernel");
    return ISP_OK;
  }
  const long long total = (long long)B * (2 * h + 6) * (2 * w + 6) * (C / 4);
  bicubic2x_pad_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(src, out, B, h, w, C);
  ISP_CHECK_LAUNCH("bicubic2x_pad_kernel");
  return ISP_OK;
}
