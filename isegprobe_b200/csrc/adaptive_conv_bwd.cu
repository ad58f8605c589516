// AdaptiveConv backward with respect to the (padded) input, NHWC fp32 -- FeatUp's AdaptiveConv autograd Function, the part
// iSegProbe needs: the 7x7 kernels depend on the guidance image only, so under the trainer's backward
// (core/training/trainer.py:213-221) only grad_input flows, through the frozen JBU stack to the backbone features.
//   gi[b,Y,X,c] = sum_{i,j<7} go[b,Y-i,X-j,c] * f[b,Y-i,X-j,i*7+j]      (gi has the padded shape [B,H+6,W+6,C])
#include "common.cuh"

namespace isp {

// thread = (padded pixel, 4 channels); gathers <= 49 (pixel, tap) pairs
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

extern "C" int isp_adaptive_conv_grad_input(const float* grad_out, const float* filters, float* grad_in, int B, int H,
                                            int W, int C, isp_stream_t stream) {
  ISP_REQUIRE(grad_out && filters && grad_in, ISP_ERR_BAD_SHAPE, "adaptive_conv_grad_input: null pointer");
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, ISP_ERR_BAD_SHAPE, "adaptive_conv_grad_input: bad shape B=%d H=%d W=%d C=%d",
              B, H, W, C);
  ISP_REQUIRE(C % 4 == 0, ISP_ERR_UNSUPPORTED, "adaptive_conv_grad_input: C %% 4 == 0 required (C=%d)", C);
  ISP_REQUIRE(aligned16(grad_out) && aligned16(grad_in), ISP_ERR_MISALIGNED, "adaptive_conv_grad_input: alignment");
  const long long total = (long long)B * (H + 6) * (W + 6) * (C / 4);
  adaptive_conv_grad_input_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(grad_out, filters, grad_in, B, H, W, C);
  ISP_CHECK_LAUNCH("adaptive_conv_grad_input_kernel");
  return ISP_OK;
}
