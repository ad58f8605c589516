// Small kernels for the LiFT x2 upsampler (a13 of SURVEY.md section 8a;
// core/model/upsamplers/LiFT.py:47-122).  LiFT is ~8 GFLOP/image: the two 3x3 convs at
// 2h x 2w and the 1x1 run on the tcgen05 GEMM core; what is left is tiny-channel,
// stride-2 image convolutions, the adaptive max-pool and channel (de)interleaving, all
// plain SIMT, channels-last.
#include "common.cuh"

namespace isp {

// conv3x3, stride 2, padding 1, Cout == 32, Cin <= 32, folded-BN bias + ReLU.
// in: f32, element strides (sb,sc,sh,sw) -> out NHWC f32 [B,Ho,Wo,32].
// weights [32][Cin][3][3] (torch layout) are staged to smem as [ (ky*3+kx)*Cin + ci ][ co ].
// thread = (output pixel, co): the warp's 32 lanes share the pixel (broadcast input reads).
template <bool RELU>
__global__ void __launch_bounds__(256) conv3x3_s2_c32_kernel(const float* __restrict__ in, long long sb, long long sc,
                                                             long long sh, long long sw, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int B, int Cin, int Hi, int Wi, int Ho, int Wo) {
  __shared__ float ws[9 * 32 * 32];
  for (int i = threadIdx.x; i < 9 * Cin * 32; i += blockDim.x) {
    const int co = i % 32, k = i / 32;
    const int ci = k % Cin, kk = k / Cin;
    ws[i] = w[(co * Cin + ci) * 9 + kk];
  }
  __syncthreads();
  const int co = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= (long long)B * Ho * Wo) return;
  const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
  float acc = bias[co];
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= Hi) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= Wi) continue;
      const float* p = in + b * sb + iy * sh + ix * sw;
      const float* wk = ws + ((ky * 3 + kx) * Cin) * 32 + co;
      for (int ci = 0; ci < Cin; ++ci) acc = fmaf(p[ci * sc], wk[ci * 32], acc);
    }
  }
  out[pix * 32 + co] = RELU ? fmaxf(acc, 0.f) : acc;
}

// F.adaptive_max_pool2d on NHWC f32: window [floor(o*I/O), ceil((o+1)*I/O))
__global__ void __launch_bounds__(256) adaptive_maxpool_nhwc_kernel(const float* __restrict__ in,
                                                                    float* __restrict__ out, int B, int C, int Hi,
                                                                    int Wi, int Ho, int Wo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo * C) return;
  const int c = (int)(idx % C);
  const long long p = idx / C;
  const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), b = (int)(p / ((long long)Wo * Ho));
  const int y0 = (int)(((long long)oy * Hi) / Ho), y1 = (int)((((long long)oy + 1) * Hi + Ho - 1) / Ho);
  const int x0 = (int)(((long long)ox * Wi) / Wo), x1 = (int)((((long long)ox + 1) * Wi + Wo - 1) / Wo);
  float m = -INFINITY;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) m = fmaxf(m, in[(((long long)b * Hi + y) * Wi + x) * C + c]);
  out[idx] = m;
}

// Generic strided channel copy: dst[b, y, x, c] (element strides db,dh,dw; channel stride 1)
// = src[b, c, y, x] (element strides sb,sc,sh,sw), f32|bf16 -> f32|bf16.  Used for concat,
// pixel-shuffle of the transposed-conv GEMM output and layout changes.
__global__ void __launch_bounds__(256) copy_channels_kernel(const void* __restrict__ src, int src_bf16, long long sb,
                                                            long long sc, long long sh, long long sw,
                                                            void* __restrict__ dst, int dst_bf16, long long db,
                                                            long long dh, long long dw, int B, int C, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * C) return;
  const int c = (int)(idx % C);
  const long long p = idx / C;
  const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
  const long long si = b * sb + c * sc + y * sh + x * sw;
  const float v = src_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[si])
                           : reinterpret_cast<const float*>(src)[si];
  const long long di = b * db + y * dh + x * dw + c;
  if (dst_bf16) reinterpret_cast<__nv_bfloat16*>(dst)[di] = __float2bfloat16(v);
  else reinterpret_cast<float*>(dst)[di] = v;
}

}  // namespace isp

using namespace isp;

static int conv3x3_s2_c32_common(const float* in, long long sb, long long sc, long long sh, long long sw, const float* w,
                                 const float* bias, float* out, int B, int Cin, int Hi, int Wi, bool relu,
                                 isp_stream_t stream) {
  ISP_REQUIRE(in && w && bias && out, ISP_ERR_BAD_SHAPE, "conv3x3_s2_c32: null pointer");
  ISP_REQUIRE(B > 0 && Cin > 0 && Cin <= 32 && Hi > 0 && Wi > 0, ISP_ERR_BAD_SHAPE, "conv3x3_s2_c32: bad shape (Cin <= 32)");
  const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
  const long long npix = (long long)B * Ho * Wo;
  if (relu)
    conv3x3_s2_c32_kernel<true><<<cdiv(npix, 8), 256, 0, as_stream(stream)>>>(in, sb, sc, sh, sw, w, bias, out, B, Cin, Hi,
                                                                            Wi, Ho, Wo);
  else
    conv3x3_s2_c32_kernel<false><<<cdiv(npix, 8), 256, 0, as_stream(stream)>>>(in, sb, sc, sh, sw, w, bias, out, B, Cin, Hi,
                                                                             Wi, Ho, Wo);
  ISP_CHECK_LAUNCH("conv3x3_s2_c32_kernel");
  return ISP_OK;
}

extern "C" int isp_conv3x3_s2_c32(const float* in, long long sb, long long sc, long long sh, long long sw,
                                  const float* w, const float* bias, float* out, int B, int Cin, int Hi, int Wi,
                                  isp_stream_t stream) {
  return conv3x3_s2_c32_common(in, sb, sc, sh, sw, w, bias, out, B, Cin, Hi, Wi, true, stream);
}

// the same convolution without the ReLU: the raw output a training-mode BatchNorm takes its batch statistics from
extern "C" int isp_conv3x3_s2_c32_raw(const float* in, long long sb, long long sc, long long sh, long long sw,
                                      const float* w, const float* bias, float* out, int B, int Cin, int Hi, int Wi,
                                      isp_stream_t stream) {
  return conv3x3_s2_c32_common(in, sb, sc, sh, sw, w, bias, out, B, Cin, Hi, Wi, false, stream);
}

extern "C" int isp_adaptive_maxpool_nhwc(const float* in, float* out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                         isp_stream_t stream) {
  ISP_REQUIRE(in && out && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, ISP_ERR_BAD_SHAPE,
              "adaptive_maxpool_nhwc: bad arguments");
  const long long total = (long long)B * Ho * Wo * C;
  adaptive_maxpool_nhwc_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(in, out, B, C, Hi, Wi, Ho, Wo);
  ISP_CHECK_LAUNCH("adaptive_maxpool_nhwc_kernel");
  return ISP_OK;
}

extern "C" int isp_copy_channels(const void* src, int src_bf16, long long sb, long long sc, long long sh, long long sw,
                                 void* dst, int dst_bf16, long long db, long long dh, long long dw, int B, int C, int H,
                                 int W, isp_stream_t stream) {
  ISP_REQUIRE(src && dst && B > 0 && C > 0 && H > 0 && W > 0, ISP_ERR_BAD_SHAPE, "copy_channels: bad arguments");
  const long long total = (long long)B * H * W * C;
  copy_channels_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(src, src_bf16, sb, sc, sh, sw, dst, dst_bf16, db,
                                                                      dh, dw, B, C, H, W);
  ISP_CHECK_LAUNCH("copy_channels_kernel");
  return ISP_OK;
}
