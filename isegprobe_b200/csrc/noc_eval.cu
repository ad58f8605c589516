// Device side of the NoC evaluation loop: the inference transforms either side of the network
// and the click simulator, so that a click costs ONE small device->host read instead of the reference's numpy round trips.
//
//   isp_zoom_in_fwd     ZoomIn._transform + AddHorizontalFlip.transform (core/inference/transforms/zoom_in.py:51-104,216-240,
//                       flip.py:13-29): crop the ROI of (image || previous probability map), bilinear (align_corners=True)
//                       resize to the network size, and write the horizontally flipped copy next to it.
//   isp_unzoom_probs    the inverse chain (flip.py:31-36 averages the LOGITS, base_transform.py:38-39 sigmoid, zoom_in.py:106-130
//                       resizes the probabilities back to the ROI and pastes them into a zero map), fused with what the driver
//                       does next on the host (core/inference/evaluation.py:76-84, utils.py:107-120): threshold, IoU counts
//                       against the ground truth (ignore label excluded), and the bounding box of {p > 0.5} the next ZoomIn
//                       needs (zoom_in.py:59-66).
//   isp_noc_next_click  Clicker._get_next_click (core/inference/clicker.py:58-91): exact Euclidean distance transform
//                       (cv2.distanceTransform(DIST_L2, DIST_MASK_PRECISE) of the zero-padded masks) of the false-negative and
//                       false-positive regions, already-clicked pixels zeroed, first maximum in row-major order.
//
// Interpolation follows torch's upsample_bilinear2d (align_corners=True): src = dst * (in-1)/(out-1) in fp32, i1 = i0 + (i0 < in-1),
// value = (1-ly)((1-lx) v00 + lx v01) + ly((1-lx) v10 + lx v11).
#include <limits.h>

#include "common.cuh"

namespace isp {

__device__ __forceinline__ float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

__device__ __forceinline__ float lerp4(float v00, float v01, float v10, float v11, float lx, float ly) {
  const float hx = 1.f - lx, hy = 1.f - ly;
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// one thread per output pixel; channels 0..2 from the image, 3 from the previous probability map (zeros when null)
__global__ void __launch_bounds__(256) zoom_in_kernel(const float* __restrict__ image, const float* __restrict__ prev, int Hs,
                                                      int Ws, int rmin, int cmin, int rh, int rw, float* __restrict__ out,
                                                      int S0, int S1, int with_flip) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= S1) return;
  const float sy = ac_scale(rh, S0) * (float)oy, sx = ac_scale(rw, S1) * (float)ox;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < rh - 1 ? 1 : 0), x1 = x0 + (x0 < rw - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const size_t plane = (size_t)Hs * Ws;
  const size_t i00 = (size_t)(rmin + y0) * Ws + cmin + x0, i01 = (size_t)(rmin + y0) * Ws + cmin + x1;
  const size_t i10 = (size_t)(rmin + y1) * Ws + cmin + x0, i11 = (size_t)(rmin + y1) * Ws + cmin + x1;
  const size_t oplane = (size_t)S0 * S1;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float* src = c < 3 ? image + c * plane : prev;
    const float v = src ? lerp4(__ldg(src + i00), __ldg(src + i01), __ldg(src + i10), __ldg(src + i11), lx, ly) : 0.f;
    out[c * oplane + (size_t)oy * S1 + ox] = v;
    if (with_flip) out[(4 + c) * oplane + (size_t)oy * S1 + (S1 - 1 - ox)] = v;
  }
}

// stats: [0] intersection, [1] union, [2] pixels with p > box_thr, [3] rmin, [4] rmax, [5] cmin, [6] cmax of those pixels
__global__ void unzoom_init_kernel(int* stats) {
  stats[0] = stats[1] = stats[2] = 0;
  stats[3] = INT_MAX; stats[4] = -1; stats[5] = INT_MAX; stats[6] = -1; stats[7] = 0;
}

__device__ __forceinline__ float prob_at(const float* __restrict__ l0, const float* __restrict__ l1, int S1, int y, int x,
                                         int with_flip) {
  float v = __ldg(l0 + (size_t)y * S1 + x);
  if (with_flip) v = 0.5f * (v + __ldg(l1 + (size_t)y * S1 + (S1 - 1 - x)));
  return 1.f / (1.f + expf(-v));
}

__global__ void __launch_bounds__(256) unzoom_kernel(const float* __restrict__ logits, int S0, int S1, int with_flip, int Hs,
                                                     int Ws, int rmin, int cmin, int rh, int rw, float* __restrict__ prob,
                                                     const int* __restrict__ gt, float pred_thr, float box_thr,
                                                     unsigned char* __restrict__ pred_mask, int* __restrict__ stats) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  int inter = 0, uni = 0, nbox = 0;
  if (x < Ws) {
    float p = 0.f;
    const int ry = y - rmin, rx = x - cmin;
    if (ry >= 0 && ry < rh && rx >= 0 && rx < rw) {
      const float sy = ac_scale(S0, rh) * (float)ry, sx = ac_scale(S1, rw) * (float)rx;
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = y0 + (y0 < S0 - 1 ? 1 : 0), x1 = x0 + (x0 < S1 - 1 ? 1 : 0);
      const float* l1 = logits + (size_t)S0 * S1;
      p = lerp4(prob_at(logits, l1, S1, y0, x0, with_flip), prob_at(logits, l1, S1, y0, x1, with_flip),
                prob_at(logits, l1, S1, y1, x0, with_flip), prob_at(logits, l1, S1, y1, x1, with_flip), sx - (float)x0,
                sy - (float)y0);
    }
    const size_t i = (size_t)y * Ws + x;
    prob[i] = p;
    const bool m = p > pred_thr;
    pred_mask[i] = m ? 1 : 0;
    if (gt) {
      const int g = __ldg(gt + i);
      const bool keep = g != -1, obj = g == 1;
      inter = (m && obj && keep) ? 1 : 0;
      uni = ((m || obj) && keep) ? 1 : 0;
    }
    if (p > box_thr) {
      nbox = 1;
      atomicMin(&stats[5], x);
      atomicMax(&stats[6], x);
    }
  }
  // block-level counts (one row segment per block)
  const int ai = __syncthreads_count(inter), au = __syncthreads_count(uni), ab = __syncthreads_count(nbox);
  if (threadIdx.x == 0) {
    if (ai) atomicAdd(&stats[0], ai);
    if (au) atomicAdd(&stats[1], au);
    if (ab) {
      atomicAdd(&stats[2], ab);
      atomicMin(&stats[3], y);
      atomicMax(&stats[4], y);
    }
  }
}

// ---------------------------------------------------------------------------------------------- click simulator
// vertical pass: g[m][y][x] = distance (in pixels) from (y, x) to the nearest zero of map m in column x, the rows -1 and H of
// the zero padding included.  m = 0: false negatives (gt == 1 and not predicted), m = 1: false positives; ignore pixels are zero.
__global__ void __launch_bounds__(128) edt_cols_kernel(const int* __restrict__ gt, const unsigned char* __restrict__ pred, int H,
                                                       int W, int* __restrict__ g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (x >= W) return;
  int* gm = g + (size_t)m * H * W;
  int d = 0;
  for (int y = 0; y < H; ++y) {
    const int gv = __ldg(gt + (size_t)y * W + x);
    const bool p = pred[(size_t)y * W + x] != 0;
    const bool on = m == 0 ? (gv == 1 && !p) : (gv != 1 && gv != -1 && p);
    d = on ? d + 1 : 0;
    gm[(size_t)y * W + x] = d;
  }
  d = 0;
  for (int y = H - 1; y >= 0; --y) {
    const int up = gm[(size_t)y * W + x];
    d = up ? d + 1 : 0;
    gm[(size_t)y * W + x] = min(up, d);
  }
}

// horizontal pass + argmax: d2(x) = min_x' (x - x')^2 + g(x')^2 with zero columns at -1 and W; dt = sqrtf(d2) (exact: d2 < 2^24),
// zero at clicked pixels; best[m] = max over the map of (float bits << 32 | ~linear index): first maximum in row-major order
__global__ void __launch_bounds__(256) edt_rows_kernel(const int* __restrict__ g, const unsigned char* __restrict__ clicked, int H,
                                                       int W, unsigned long long* __restrict__ best) {
  extern __shared__ int grow[];
  const int y = blockIdx.x, m = blockIdx.y;
  const int* gm = g + ((size_t)m * H + y) * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const int v = gm[x];
    grow[x] = v * v;
  }
  __syncthreads();
  unsigned long long key = 0;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int d2 = min((x + 1) * (x + 1), (W - x) * (W - x));
    if (grow[x] < d2) d2 = grow[x];
    if (d2 > 0) {
      // candidates further away than sqrt(d2) cannot win: walk outwards from x
      for (int k = 1; k * k < d2; ++k) {
        if (x - k >= 0) d2 = min(d2, k * k + grow[x - k]);
        if (x + k < W) d2 = min(d2, k * k + grow[x + k]);
      }
    }
    float dt = sqrtf((float)d2);
    if (clicked[(size_t)y * W + x]) dt = 0.f;
    const unsigned int idx = (unsigned int)((size_t)y * W + x);
    const unsigned long long k2 = ((unsigned long long)__float_as_uint(dt) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
    key = k2 > key ? k2 : key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  __shared__ unsigned long long wbest[8];
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) key = wbest[w] > key ? wbest[w] : key;
    atomicMax(best + m, key);
  }
}

__global__ void click_init_kernel(unsigned long long* best) { best[0] = best[1] = 0ull; }

}  // namespace isp

using namespace isp;

extern "C" int isp_zoom_in_fwd(const float* image, const float* prev, int Hs, int Ws, int rmin, int rmax, int cmin, int cmax,
                               float* out, int S0, int S1, int with_flip, isp_stream_t stream) {
  ISP_REQUIRE(image && out, ISP_ERR_BAD_SHAPE, "zoom_in_fwd: null pointer");
  ISP_REQUIRE(Hs > 0 && Ws > 0 && S0 > 0 && S1 > 0 && rmin >= 0 && cmin >= 0 && rmax >= rmin && cmax >= cmin && rmax < Hs &&
                  cmax < Ws,
              ISP_ERR_BAD_SHAPE, "zoom_in_fwd: ROI [%d,%d]x[%d,%d] outside the %dx%d image", rmin, rmax, cmin, cmax, Hs, Ws);
  dim3 grid((S1 + 255) / 256, S0);
  zoom_in_kernel<<<grid, 256, 0, as_stream(stream)>>>(image, prev, Hs, Ws, rmin, cmin, rmax - rmin + 1, cmax - cmin + 1, out,
                                                      S0, S1, with_flip ? 1 : 0);
  ISP_CHECK_LAUNCH("zoom_in_kernel");
  return ISP_OK;
}

extern "C" int isp_unzoom_probs(const float* logits, int S0, int S1, int with_flip, int Hs, int Ws, int rmin, int rmax,
                                int cmin, int cmax, float* prob, const int* gt, float pred_thr, float box_thr,
                                unsigned char* pred_mask, int* stats, isp_stream_t stream) {
  ISP_REQUIRE(logits && prob && pred_mask && stats, ISP_ERR_BAD_SHAPE, "unzoom_probs: null pointer");
  ISP_REQUIRE(Hs > 0 && Ws > 0 && S0 > 0 && S1 > 0 && rmin >= 0 && cmin >= 0 && rmax >= rmin && cmax >= cmin && rmax < Hs &&
                  cmax < Ws,
              ISP_ERR_BAD_SHAPE, "unzoom_probs: ROI outside the image");
  unzoom_init_kernel<<<1, 1, 0, as_stream(stream)>>>(stats);
  dim3 grid((Ws + 255) / 256, Hs);
  unzoom_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, S0, S1, with_flip ? 1 : 0, Hs, Ws, rmin, cmin, rmax - rmin + 1,
                                                     cmax - cmin + 1, prob, gt, pred_thr, box_thr, pred_mask, stats);
  ISP_CHECK_LAUNCH("unzoom_kernel");
  count_launch(1);
  return ISP_OK;
}

// work: 2*H*W ints; best: 2 x uint64 (false-negative map, false-positive map): high word = float bits of the largest
// distance, low word = 0xFFFFFFFF - linear index of its first occurrence
extern "C" int isp_noc_next_click(const int* gt, const unsigned char* pred_mask, const unsigned char* clicked, int H, int W,
                                  int* work, unsigned long long* best, isp_stream_t stream) {
  ISP_REQUIRE(gt && pred_mask && clicked && work && best, ISP_ERR_BAD_SHAPE, "noc_next_click: null pointer");
  ISP_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 31) && W <= 8192, ISP_ERR_BAD_SHAPE, "noc_next_click: bad shape");
  click_init_kernel<<<1, 1, 0, as_stream(stream)>>>(best);
  edt_cols_kernel<<<dim3((W + 127) / 128, 2), 128, 0, as_stream(stream)>>>(gt, pred_mask, H, W, work);
  edt_rows_kernel<<<dim3(H, 2), 256, W * sizeof(int), as_stream(stream)>>>(work, clicked, H, W, best);
  ISP_CHECK_LAUNCH("edt_rows_kernel");
  count_launch(2);
  return ISP_OK;
}
