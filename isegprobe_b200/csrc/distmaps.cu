// Click-map encoding (a2/a3/a1 of SURVEY.md section 8a).
//
// The reference materialises a [B*2P, 2, H, W] temporary and min-reduces it
// (core/model/ops.py:49-70).  Here each thread owns 4 consecutive pixels of one
// (image, polarity) plane and scans the <= 24 clicks held in shared memory:
// the only HBM traffic is the 4-byte/pixel output (write-bound, 128-bit stores).
// Arithmetic is float32 with explicit __fsub_rn/__fmul_rn/__fadd_rn so no FMA
// contraction can change a bit relative to the reference's separate ops.
#include "common.cuh"

namespace isp {

constexpr int kMaxClicks = 64;  // per polarity

struct ClickSet {
  float r[kMaxClicks];
  float c[kMaxClicks];
  int n;
};

// mode 0: torch path (ops.py:35-77).  mode 1: Cython path (pyx:18-64): rounded
// coordinates, validity on the row only, squared distance / nd^2, no final step.
template <int MODE>
__device__ __forceinline__ void load_clicks(ClickSet& cs, const float* __restrict__ pts, int P, float scale) {
  if (threadIdx.x == 0) {
    int n = 0;
    for (int p = 0; p < P; ++p) {
      float r = pts[p * 3 + 0], c = pts[p * 3 + 1];
      if (MODE == 0) {
        if (fmaxf(r, c) < 0.f) continue;               // ops.py:40
        cs.r[n] = __fmul_rn(r, scale);                  // ops.py:55
        cs.c[n] = __fmul_rn(c, scale);
      } else {
        float rr = rintf(r), rc = rintf(c);             // pyx:31 (python round == half-even)
        if (rr < 0.f) continue;                         // pyx:32
        cs.r[n] = rr;
        cs.c[n] = rc;
      }
      ++n;
    }
    cs.n = n;
  }
  __syncthreads();
}

__device__ __forceinline__ float min_sqdist(const ClickSet& cs, float row, float col, bool divide, float div) {
  float best = 1e6f;  // ops.py:66
  for (int p = 0; p < cs.n; ++p) {
    float dr = __fsub_rn(row, cs.r[p]);
    float dc = __fsub_rn(col, cs.c[p]);
    if (divide) {  // ops.py:59-60 (IEEE division, as ATen)
      dr = __fdiv_rn(dr, div);
      dc = __fdiv_rn(dc, div);
    }
    float d2 = __fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dc, dc));
    best = fminf(best, d2);
  }
  return best;
}

__device__ __forceinline__ float finish(float d2, bool disks, float thr) {
  if (disks) return d2 <= thr ? 1.f : 0.f;                 // ops.py:72-73
  return tanhf(__fmul_rn(__fsqrt_rn(d2), 2.f));            // ops.py:75
}

// grid: (ceil(W/4/128) * H, 2, B)
template <int MODE>
__global__ void __launch_bounds__(128) distmaps_kernel(const float* __restrict__ points, float* __restrict__ out,
                                                       int P, int H, int W, float div, float scale, int disks,
                                                       float thr) {
  __shared__ ClickSet cs;
  const int b = blockIdx.z, s = blockIdx.y;
  load_clicks<MODE>(cs, points + ((size_t)b * 2 * P + (size_t)s * P) * 3, P, scale);
  const int wq = (W + 3) / 4;
  const int bpr = (wq + blockDim.x - 1) / blockDim.x;  // blocks per row
  const int row = blockIdx.x / bpr;
  const int q = (blockIdx.x % bpr) * blockDim.x + threadIdx.x;
  if (q >= wq) return;
  const bool divide = (MODE == 0) ? !disks : true;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float d2 = min_sqdist(cs, (float)row, (float)(q * 4 + k), divide, div);
    v[k] = (MODE == 0) ? finish(d2, disks != 0, thr) : d2;
  }
  float* o = out + (((size_t)b * 2 + s) * H + row) * W + (size_t)q * 4;
  if ((W & 3) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int k = 0; k < 4 && q * 4 + k < W; ++k) o[k] = v[k];
  }
}

// Fused prepare_input: normalise RGB, pass the previous mask through and write the
// click maps behind it (iseg_base_model.py:91-110).  grid: (blocks over H*W/4, 1, B)
__global__ void __launch_bounds__(128) prepare_input_kernel(
    const float* __restrict__ image, const float* __restrict__ points, float* __restrict__ norm_image,
    float* __restrict__ coord, int Cin, int P, int H, int W, float3 mean, float3 stdv, float div, float scale,
    int disks, float thr) {
  __shared__ ClickSet cs[2];
  const int b = blockIdx.z;
  load_clicks<0>(cs[0], points + ((size_t)b * 2 * P) * 3, P, scale);
  load_clicks<0>(cs[1], points + ((size_t)b * 2 * P + P) * 3, P, scale);
  const size_t hw = (size_t)H * W;
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= hw) return;
  const float* img = image + (size_t)b * Cin * hw;
  const int Cc = (Cin == 4) ? 3 : 2;
  float* nimg = norm_image + (size_t)b * 3 * hw;
  float* co = coord + (size_t)b * Cc * hw;
  const float m[3] = {mean.x, mean.y, mean.z}, sd[3] = {stdv.x, stdv.y, stdv.z};
  const int n = (int)min((size_t)4, hw - i4);
  for (int k = 0; k < n; ++k) {
    const size_t i = i4 + k;
    const int row = (int)(i / W), col = (int)(i % W);
#pragma unroll
    for (int c = 0; c < 3; ++c)  // ops.py:104: sub_ then div_
      nimg[c * hw + i] = __fdiv_rn(__fsub_rn(img[c * hw + i], m[c]), sd[c]);
    int o = 0;
    if (Cin == 4) co[(o++) * hw + i] = img[3 * hw + i];
    co[(o++) * hw + i] = finish(min_sqdist(cs[0], (float)row, (float)col, !disks, div), disks != 0, thr);
    co[(o++) * hw + i] = finish(min_sqdist(cs[1], (float)row, (float)col, !disks, div), disks != 0, thr);
  }
}

}  // namespace isp

using namespace isp;

static int check_distmaps(const void* points, const void* out, int B, int P, int H, int W) {
  ISP_REQUIRE((points || P == 0) && out, ISP_ERR_BAD_SHAPE, "distmaps: null pointer");
  ISP_REQUIRE(B > 0 && H > 0 && W > 0 && P >= 0, ISP_ERR_BAD_SHAPE, "distmaps: bad shape B=%d P=%d H=%d W=%d", B, P, H, W);
  ISP_REQUIRE(P <= kMaxClicks, ISP_ERR_UNSUPPORTED, "distmaps: P=%d exceeds %d clicks per polarity", P, kMaxClicks);
  ISP_REQUIRE(B <= 65535, ISP_ERR_UNSUPPORTED, "distmaps: B=%d exceeds grid.z", B);
  return ISP_OK;
}

extern "C" int isp_distmaps_fwd(const float* points, float* out, int B, int P, int H, int W, float norm_radius,
                                float spatial_scale, int use_disks, isp_stream_t stream) {
  if (int e = check_distmaps(points, out, B, P, H, W)) return e;
  ISP_REQUIRE((W % 4) || aligned16(out), ISP_ERR_MISALIGNED, "distmaps: out must be 16-byte aligned");
  const int wq = (W + 3) / 4, bpr = cdiv(wq, 128);
  dim3 grid(bpr * H, 2, B);
  const float div = norm_radius * spatial_scale;           // ops.py:60 (python double product, then f32)
  const float thr = (float)((double)(norm_radius * spatial_scale) * (double)(norm_radius * spatial_scale));
  distmaps_kernel<0><<<grid, 128, 0, as_stream(stream)>>>(points, out, P, H, W, div, spatial_scale, use_disks, thr);
  ISP_CHECK_LAUNCH("distmaps_kernel");
  return ISP_OK;
}

extern "C" int isp_distmaps_rounded_sqdist_fwd(const float* points, float* out, int B, int P, int H, int W,
                                               float norm_delimeter, isp_stream_t stream) {
  if (int e = check_distmaps(points, out, B, P, H, W)) return e;
  ISP_REQUIRE((W % 4) || aligned16(out), ISP_ERR_MISALIGNED, "distmaps: out must be 16-byte aligned");
  const int wq = (W + 3) / 4, bpr = cdiv(wq, 128);
  dim3 grid(bpr * H, 2, B);
  distmaps_kernel<1><<<grid, 128, 0, as_stream(stream)>>>(points, out, P, H, W, norm_delimeter, 1.f, 0, 0.f);
  ISP_CHECK_LAUNCH("distmaps_kernel<bfs>");
  return ISP_OK;
}

extern "C" int isp_prepare_input_fwd(const float* image, const float* points, float* norm_image, float* coord,
                                     int B, int Cin, int P, int H, int W, const float* mean3, const float* std3,
                                     float norm_radius, float spatial_scale, int use_disks, isp_stream_t stream) {
  if (int e = check_distmaps(points, coord, B, P, H, W)) return e;
  ISP_REQUIRE(image && norm_image && mean3 && std3, ISP_ERR_BAD_SHAPE, "prepare_input: null pointer");
  ISP_REQUIRE(Cin == 3 || Cin == 4, ISP_ERR_BAD_SHAPE, "prepare_input: Cin must be 3 or 4, got %d", Cin);
  const size_t hw = (size_t)H * W;
  dim3 grid(cdiv((long long)(hw + 3) / 4, 128), 1, B);
  const float div = norm_radius * spatial_scale;
  const float thr = (float)((double)(norm_radius * spatial_scale) * (double)(norm_radius * spatial_scale));
  prepare_input_kernel<<<grid, 128, 0, as_stream(stream)>>>(
      image, points, norm_image, coord, Cin, P, H, W, make_float3(mean3[0], mean3[1], mean3[2]),
      make_float3(std3[0], std3[1], std3[2]), div, spatial_scale, use_disks, thr);
  ISP_CHECK_LAUNCH("prepare_input_kernel");
  return ISP_OK;
}
