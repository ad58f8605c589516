// tcgen05 GEMM / implicit-GEMM 3x3 convolution core (sm_100a).
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )          bf16 operands, fp32 accumulate in TMEM
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected lane: tcgen05.mma 128 x BN x 16, tcgen05.commit)
//   warps 2-5: epilogue       (tcgen05.ld TMEM->registers, bias/activation/residual, global store)
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// GEMM mode   : A is a row-major [M, K] bf16 matrix (2D tensor map).
// CONV3x3 mode: A is an NHWC bf16 activation [Nimg, H, W, Cin]; a tile is TH x TW output
//               pixels (TH*TW = 128); for every tap (r,s) and 64-channel chunk the producer
//               issues one 4D TMA box at (h0+r-1, w0+s-1) -- out-of-bounds rows/cols are
//               zero-filled by TMA, which IS the padding=1 -- so im2col never exists in memory.
//               W is [Cout, 9*Cin_pad] with k = (r*3+s)*Cin_pad + c.
// Replaces: cuDNN/cuBLAS fp32 calls under nn.Conv2d / nn.Linear in LoftUp, ConvSegHead and
// the ViT blocks.
#include "tc_common.cuh"

namespace isp {

tmap_encode_fn get_tmap_encode() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tmap_encode_fn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char* what) {
  tmap_encode_fn enc = get_tmap_encode();
  ISP_REQUIRE(enc, ISP_ERR_CUDA, "%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", what);
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i + 1];  // stride of dim i+1
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ISP_REQUIRE(r == CUDA_SUCCESS, ISP_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return ISP_OK;
}

namespace gemm {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;  // also forces one CTA per SM (TMEM: 512 columns each)

struct Params {
  // problem
  long long M;        // rows (GEMM) or Nimg*H*W (conv)
  int N, K;           // K = padded reduction length actually looped (multiple of 64)
  int BN;             // tile width, multiple of 16, <= 256
  int stages;
  // conv mode (TW == 0 -> plain GEMM)
  int TW, TH, H, W, cin_chunks;  // cin_chunks = Cin_pad / 64
  long long tiles_m, tiles_n;
  int tiles_w, tiles_h;          // conv: tiles per image row / column
  // epilogue
  const float* bias;   // [N] or null
  const void* resid;   // [M, ldr] or null
  int resid_bf16, ldr;
  float alpha;         // out = alpha * act(acc + bias) + resid
  int act;             // 0 none, 1 relu, 2 gelu(erf), 3 quick-gelu
  void* D;
  int ldd, out_bf16;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return gelu_erf(v);
  if (act == 3) return v / (1.f + __expf(-1.702f * v));
  return v;
}

struct TileCoord {
  long long m0;     // first output row of the tile (GEMM) / unused (conv)
  int n0;
  int img, h0, w0;  // conv
};

__device__ __forceinline__ TileCoord tile_coord(const Params& p, long long t) {
  TileCoord c;
  const long long tm = t / p.tiles_n;
  c.n0 = (int)(t % p.tiles_n) * p.BN;
  c.m0 = tm * BM;
  c.img = c.h0 = c.w0 = 0;
  if (p.TW) {
    const int per_img = p.tiles_w * p.tiles_h;
    c.img = (int)(tm / per_img);
    const int r = (int)(tm % per_img);
    c.h0 = (r / p.tiles_w) * p.TH;
    c.w0 = (r % p.tiles_w) * p.TW;
  }
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2, stage_bytes = a_bytes + b_bytes;
  // dynamic smem base is 1024-aligned by declaration; keep tiles 1024-aligned (b_bytes % 1024 == 0 since BN % 8 == 0)
  const long long ntiles = p.tiles_m * p.tiles_n;
  const int kblocks = p.K / BK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull_bar[a], 1); tc::mbar_init(&tempty_bar[a], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const TileCoord tc_ = tile_coord(p, t);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          if (p.TW) {
            const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            const int r = tap / 3, q = tap - r * 3;
            tc::tma_load_4d(sa, &tmA, &full_bar[s], cc * BK, tc_.w0 + q - 1, tc_.h0 + r - 1, tc_.img);
          } else {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, (int)tc_.m0);
          }
          tc::tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, tc_.n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    const uint32_t idesc = tc::idesc_bf16_f32(BM, p.BN);
    uint32_t it = 0, tl = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      tc::mbar_wait(&tempty_bar[acc], aph ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;  // accumulator stages at columns 0 and 256
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        tc::mbar_wait(&full_bar[s], ph);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t adesc = tc::smem_desc_k_sw128(sa), bdesc = tc::smem_desc_k_sw128(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // +32 B per 16-element K step inside the 128B swizzle row
            tc::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          tc::umma_commit(&empty_bar[s]);                       // smem slot free when these MMAs retire
          if (kb == kblocks - 1) tc::umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    uint32_t tl = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      const TileCoord tc_ = tile_coord(p, t);
      long long row;
      bool row_ok;
      if (p.TW) {
        const int hh = tc_.h0 + row_in_tile / p.TW, ww = tc_.w0 + row_in_tile % p.TW;
        row_ok = hh < p.H && ww < p.W;
        row = ((long long)tc_.img * p.H + hh) * p.W + ww;
      } else {
        row = tc_.m0 + row_in_tile;
        row_ok = row < p.M;
      }
      tc::mbar_wait(&tfull_bar[acc], aph);
      tc::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        const int ncol = min(32, p.BN - c0);  // 32 or 16 (BN % 16 == 0)
        if (ncol == 32) {
          tc::tmem_ld32(t_addr + c0, v);
        } else {
          uint32_t h[16];
          tc::tmem_ld16(t_addr + c0, h);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = h[i];
        }
        tc::tmem_ld_wait();
        if (row_ok) {
          const int nbase = tc_.n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j >= ncol || nbase + j >= p.ldd) break;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int n = nbase + j + e;
              float x = __uint_as_float(v[j + e]);
              if (n < p.N) {
                if (p.bias) x += __ldg(p.bias + n);
                x = apply_act(x, p.act) * p.alpha;
                if (p.resid) {
                  x += p.resid_bf16
                           ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.resid)[row * p.ldr + n])
                           : reinterpret_cast<const float*>(p.resid)[row * p.ldr + n];
                }
              } else {
                x = 0.f;
              }
              o[e] = x;
            }
            const int nvalid = min(8, p.ldd - (nbase + j));  // columns up to ldd are ours to write (pad = 0)
            if (p.out_bf16) {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.D) + row * p.ldd + nbase + j;
              if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                uint4 u;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(o[4], o[5]), b3 = __floats2bfloat162_rn(o[6], o[7]);
                u.x = *reinterpret_cast<uint32_t*>(&b0); u.y = *reinterpret_cast<uint32_t*>(&b1);
                u.z = *reinterpret_cast<uint32_t*>(&b2); u.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(dst) = u;
              } else {
                for (int e = 0; e < nvalid; ++e) dst[e] = __float2bfloat16(o[e]);
              }
            } else {
              float* dst = reinterpret_cast<float*>(p.D) + row * p.ldd + nbase + j;
              if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
              } else {
                for (int e = 0; e < nvalid; ++e) dst[e] = o[e];
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, kTmemCols);
}

static int pick_bn(int N) {
  // widest tile <= 256 that wastes the least: N itself if it fits, else an even split
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= 256) return n16;
  for (int parts = 2; parts <= 16; ++parts) {
    const int bn = ((N + parts - 1) / parts + 15) / 16 * 16;
    if (bn <= 256) return bn;
  }
  return 256;
}

static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, Params& p, cudaStream_t stream) {
  static int num_sms = 0;
  static bool attr_set = false;
  if (!num_sms) {
    int dev = 0;
    ISP_CUDA(cudaGetDevice(&dev));
    ISP_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!attr_set) {
    ISP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  const int stage_bytes = BM * BK * 2 + p.BN * BK * 2;
  p.stages = kSmemBudget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  ISP_REQUIRE(p.stages >= 2, ISP_ERR_UNSUPPORTED, "gemm_tc: tile too large for a 2-stage pipeline");
  const long long ntiles = p.tiles_m * p.tiles_n;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  gemm_tc_kernel<<<grid, kThreads, kSmemBudget, stream>>>(tmA, tmB, p);
  ISP_CHECK_LAUNCH("gemm_tc_kernel");
  return ISP_OK;
}

}  // namespace gemm
}  // namespace isp

using namespace isp;

extern "C" int isp_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                const void* resid, int resid_bf16, long long ldr, float alpha, int act, void* D,
                                long long ldd, int out_bf16, long long M, int N, int K, isp_stream_t stream) {
  ISP_REQUIRE(A && W && D, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: null pointer");
  ISP_REQUIRE(M > 0 && N > 0 && K > 0, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: bad shape M=%lld N=%d K=%d", M, N, K);
  ISP_REQUIRE(lda >= K && ldw >= K && ldd >= N, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: leading dimensions too small");
  ISP_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, ISP_ERR_MISALIGNED,
              "gemm_bf16_tc: lda/ldw must be multiples of 8 elements (TMA 16-byte strides), got %lld/%lld", lda, ldw);
  ISP_REQUIRE(aligned16(A) && aligned16(W), ISP_ERR_MISALIGNED, "gemm_bf16_tc: A/W must be 16-byte aligned");
  ISP_REQUIRE(act >= 0 && act <= 3, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: unknown activation %d", act);
  ISP_REQUIRE(M < (1ll << 31), ISP_ERR_UNSUPPORTED, "gemm_bf16_tc: M too large for TMA coordinates");
  gemm::Params p = {};
  p.M = M; p.N = N;
  p.K = (K + gemm::BK - 1) / gemm::BK * gemm::BK;  // TMA zero-fills the K tail (global dim = K)
  p.BN = gemm::pick_bn(N);
  p.TW = 0;
  p.tiles_m = (M + gemm::BM - 1) / gemm::BM;
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.bias = bias; p.resid = resid; p.resid_bf16 = resid_bf16; p.ldr = (int)ldr; p.alpha = alpha; p.act = act;
  p.D = D; p.ldd = (int)ldd; p.out_bf16 = out_bf16;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {2, (uint64_t)lda * 2};
    const uint32_t box[2] = {gemm::BK, gemm::BM};
    if (int e = make_tmap_bf16(&tmA, A, 2, dims, str, box, "gemm_bf16_tc(A)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[2] = {2, (uint64_t)ldw * 2};
    const uint32_t box[2] = {gemm::BK, (uint32_t)p.BN};
    if (int e = make_tmap_bf16(&tmB, W, 2, dims, str, box, "gemm_bf16_tc(W)")) return e;
  }
  return gemm::launch(tmA, tmB, p, as_stream(stream));
}

extern "C" int isp_conv3x3_bf16_tc(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                                   int Nimg, int H, int Wd, int Cin, int ldx, int Cout, int ldy, isp_stream_t stream) {
  ISP_REQUIRE(X && Wp && Y, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: null pointer");
  ISP_REQUIRE(Nimg > 0 && H > 0 && Wd > 0 && Cout > 0 && Cin > 0, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: bad shape");
  ISP_REQUIRE(ldx >= Cin && ldx % 8 == 0, ISP_ERR_MISALIGNED,
              "conv3x3_bf16_tc: channel stride must be >= Cin and a multiple of 8 (got %d)", ldx);
  ISP_REQUIRE(ldy >= Cout, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: ldy < Cout");
  const int Cin_pad = (Cin + 63) / 64 * 64;  // weights are packed [Cout][9][Cin_pad]; TMA zero-fills c >= Cin
  ISP_REQUIRE(aligned16(X) && aligned16(Wp), ISP_ERR_MISALIGNED, "conv3x3_bf16_tc: X/W must be 16-byte aligned");
  ISP_REQUIRE(act >= 0 && act <= 3, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: unknown activation %d", act);
  gemm::Params p = {};
  p.M = (long long)Nimg * H * Wd; p.N = Cout; p.K = 9 * Cin_pad;
  p.BN = gemm::pick_bn(Cout);
  // tile = TH x TW output pixels; prefer wide rows (fewer halo re-reads), fall back to 16x8
  int TW = 16;
  for (int cand : {128, 64, 32, 16, 8}) {
    if (Wd % cand == 0 && H % (128 / cand) == 0) { TW = cand; break; }
  }
  p.TW = TW; p.TH = 128 / TW; p.H = H; p.W = Wd; p.cin_chunks = Cin_pad / 64;
  p.tiles_w = (Wd + p.TW - 1) / p.TW; p.tiles_h = (H + p.TH - 1) / p.TH;
  p.tiles_m = (long long)Nimg * p.tiles_w * p.tiles_h;
  p.tiles_n = (Cout + p.BN - 1) / p.BN;
  p.bias = bias; p.resid = nullptr; p.alpha = 1.f; p.act = act;
  p.D = Y; p.ldd = ldy; p.out_bf16 = out_bf16;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wd, (uint64_t)H, (uint64_t)Nimg};
    const uint64_t str[4] = {2, (uint64_t)ldx * 2, (uint64_t)Wd * ldx * 2, (uint64_t)H * Wd * ldx * 2};
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if (int e = make_tmap_bf16(&tmA, X, 4, dims, str, box, "conv3x3_bf16_tc(X)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)Cout}, str[2] = {2, (uint64_t)p.K * 2};
    const uint32_t box[2] = {gemm::BK, (uint32_t)p.BN};
    if (int e = make_tmap_bf16(&tmB, Wp, 2, dims, str, box, "conv3x3_bf16_tc(W)")) return e;
  }
  return gemm::launch(tmA, tmB, p, as_stream(stream));
}
