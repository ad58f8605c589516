// tcgen05 GEMM / implicit-GEMM 3x3 convolution core (sm_100a).
//
//   D[M,N] = alpha * act( A[M,K] * W[N,K]^T + bias ) + resid     bf16 operands, fp32 accumulate (TMEM)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one elected lane: tcgen05.mma 128 x BN x 16, tcgen05.commit)
//   warps 2-5: epilogue       each warp owns 32 rows: tcgen05.ld -> bias/act -> (+ residual chunk it
//                             TMA-loaded itself) -> 128B-swizzled smem chunk -> TMA store.
//                             All global traffic of the epilogue is bulk/coalesced; OOB rows and
//                             columns are clipped by TMA, so there is no per-element masking.
// TMEM holds two 256-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
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

int make_tmap(CUtensorMap* out, int elem_bytes, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const char* what, bool swizzle128) {
  tmap_encode_fn enc = get_tmap_encode();
  ISP_REQUIRE(enc, ISP_ERR_CUDA, "%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", what);
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i + 1];  // stride of dim i+1
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ISP_REQUIRE(r == CUDA_SUCCESS, ISP_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
  return ISP_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char* what) {
  return make_tmap(out, 2, base, rank, dims, strides_bytes, box, what, true);
}

namespace gemm {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter (they split the chunks)
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kEpiBytesPerWarp = 2 * 4096;         // two output staging chunks of 32 rows x 128 B
constexpr int kEpiBytes = kEpiWarps * kEpiBytesPerWarp;
constexpr int kMainBudget = 160 * 1024;
constexpr int kSmemBytes = kMainBudget + kEpiBytes;  // 224 KB: also forces one CTA per SM (TMEM: 512 columns)

struct Params {
  long long M;        // rows (GEMM) or Nimg*H*W (conv)
  int N, K;           // K = padded reduction length actually looped (multiple of 64)
  int BN;             // tile width, multiple of 16, <= 256
  int stages;
  // conv mode (TW == 0 -> plain GEMM)
  int TW, TH, H, W, cin_chunks;
  long long tiles_m, tiles_n;
  int tiles_w, tiles_h;
  // epilogue
  const float* bias;   // [N] or null
  const void* resid;   // [M, ldr] in the dtype of D, or null (GEMM mode only)
  long long ldr;
  float alpha;         // out = alpha * act(acc + bias) + resid
  int act;             // 0 none, 1 relu, 2 gelu(erf), 3 quick-gelu
  int out_bf16;
};

// erf to ~1.5e-7 absolute (Abramowitz-Stegun 7.1.26): 1 RCP + 1 EX2 + 7 FMA instead of erff's ~25
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
  const float r = fmaf(-poly * t, e, 1.f);
  return copysignf(r, x);
}

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if constexpr (ACT == 1) return fmaxf(v, 0.f);
  if constexpr (ACT == 2) return 0.5f * v * (1.f + erf_fast(v * 0.70710678118654752440f));
  if constexpr (ACT == 3) return v / (1.f + __expf(-1.702f * v));
  if constexpr (ACT == 4) {  // tanh-form GELU, one MUFU: |gelu_tanh - gelu_erf| < 5e-4, below bf16 output rounding
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(v * fmaf(v * v, 0.0356774081f, 0.7978845608f)));
    const float hv = 0.5f * v;
    return fmaf(hv, th, hv);
  }
  return v;
}

struct TileCoord {
  long long m0;
  int n0;
  int img, h0, w0;
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

// What the (non-inlined) epilogue needs, passed BY VALUE: reading a by-reference Params from the
// caller's stack cost a local-memory load per use (ncu: 59 % long-scoreboard stalls in the epilogue).
struct EpiArgs {
  long long M, ldr;
  const float* bias;
  const void* resid;
  float alpha;
  int N, BN, TW;
};

// One epilogue warp, its share of one tile.  CW = columns per 128-byte chunk (64 bf16 / 32 f32).
// The two warps of a lane quarter take alternate chunks.  Per chunk: prefetch the residual slice
// with coalesced 16-byte loads (8 lanes = one 128-byte row segment), tcgen05.ld the accumulator,
// bias/activation in registers, stage into 128B-swizzled smem, add the residual there
// (row-coalesced pass), TMA-store the chunk.  st_seq counts this warp's stores (staging buffer).
template <bool OUT_BF16, int ACT>
__device__ __noinline__ void epilogue_tile(const EpiArgs ea, const CUtensorMap* tmD, const CUtensorMap* tmDt,
                                              const TileCoord tc_, uint32_t t_addr, uint8_t* stage, uint32_t& st_seq,
                                              int q, int half, int lane) {
  constexpr int CW = OUT_BF16 ? 64 : 32;
  constexpr int ESZ = OUT_BF16 ? 2 : 4;
  const int nchunks = (ea.BN + CW - 1) / CW;
  int c1, c2 = 0, c3 = 0;  // where this warp's 32 rows live in the output tensor
  if (ea.TW) {
    const int bw = ea.TW < 32 ? ea.TW : 32;
    const int pix = q * 32;
    c1 = tc_.w0 + (ea.TW >= 32 ? pix % ea.TW : 0);
    c2 = tc_.h0 + (ea.TW >= 32 ? pix / ea.TW : q * (32 / bw));
    c3 = tc_.img;
  } else {
    c1 = (int)tc_.m0 + q * 32;
  }
  const int r7 = lane & 7;
  const int sub_r = lane >> 3, sub_u = lane & 7;  // residual pass: lane -> (row within group of 4, 16-byte unit)
  for (int c = half; c < nchunks; c += 2) {
    const int col0 = c * CW;                 // column inside the tile
    const int ncols = min(CW, ea.BN - col0);  // multiple of 16
    const bool is_tail = ncols < CW;
    const int row_bytes = ncols * ESZ;       // dense row pitch of a tail chunk
    const int nglob = tc_.n0 + col0;         // first global column of the chunk
    const bool full = nglob + ncols <= ea.N;  // no column masking needed
    // ---- residual prefetch (registers), coalesced: 8 lanes cover one row's 128-byte slice
    uint4 rres[8];
    const int units = row_bytes >> 4;
    if (ea.resid) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long grow = tc_.m0 + q * 32 + it * 4 + sub_r;
        rres[it] = make_uint4(0u, 0u, 0u, 0u);
        if (grow < ea.M && sub_u < units && nglob + sub_u * (16 / ESZ) < ea.ldr)
          rres[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(ea.resid) +
                                                          (grow * ea.ldr + nglob) * ESZ + sub_u * 16));
      }
    }
    // ---- staging buffer of this chunk (32 rows x 128 B, 16-byte units XOR-swizzled like TMA's
    //      SWIZZLE_128B; tail chunks dense and un-swizzled)
    if (lane == 0) tc::tma_store_wait_read<1>();  // the store that used this buffer two chunks ago has read it
    __syncwarp();
    uint8_t* obuf = stage + (st_seq & 1) * 4096;
    uint8_t* ob = obuf + lane * (is_tail ? row_bytes : 128);
    // ---- accumulator -> registers -> bias / activation / scale -> staging, 32 columns at a time
    //      (columns >= N give exact zeros: they are K-padding for the next GEMM)
#pragma unroll
    for (int sb = 0; sb < CW / 32; ++sb) {
      const int scol = sb * 32;
      if (scol >= ncols) break;
      uint32_t vr[32];
      if (scol + 32 <= ncols) {
        tc::tmem_ld32(t_addr + col0 + scol, vr);
      } else {  // 16 live columns
        tc::tmem_ld16(t_addr + col0 + scol, *reinterpret_cast<uint32_t(*)[16]>(&vr[0]));
#pragma unroll
        for (int e = 16; e < 32; ++e) vr[e] = 0u;
      }
      tc::tmem_ld_wait();
#pragma unroll
      for (int e4 = 0; e4 < 32; e4 += 4) {
        const int n = nglob + scol + e4;
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (ea.bias) {
          if (n + 3 < ea.N) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(ea.bias + n));
            b[0] = bb.x; b[1] = bb.y; b[2] = bb.z; b[3] = bb.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < ea.N) b[e] = __ldg(ea.bias + n + e);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = act_fn<ACT>(__uint_as_float(vr[e4 + e]) + b[e]) * ea.alpha;
          vr[e4 + e] = __float_as_uint((full || n + e < ea.N) ? x : 0.f);
        }
      }
      constexpr int UPS = OUT_BF16 ? 4 : 8;  // 16-byte units produced per 32-column sub-block
#pragma unroll
      for (int uu = 0; uu < UPS; ++uu) {
        const int u = sb * UPS + uu;
        if (u * 16 >= row_bytes) break;
        uint4 w;
        if constexpr (OUT_BF16) {
          __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(vr[uu * 8 + 0]), __uint_as_float(vr[uu * 8 + 1]));
          __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(vr[uu * 8 + 2]), __uint_as_float(vr[uu * 8 + 3]));
          __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(vr[uu * 8 + 4]), __uint_as_float(vr[uu * 8 + 5]));
          __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(vr[uu * 8 + 6]), __uint_as_float(vr[uu * 8 + 7]));
          w = make_uint4(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1),
                         *reinterpret_cast<uint32_t*>(&b2), *reinterpret_cast<uint32_t*>(&b3));
        } else {
          w = make_uint4(vr[uu * 4 + 0], vr[uu * 4 + 1], vr[uu * 4 + 2], vr[uu * 4 + 3]);
        }
        *reinterpret_cast<uint4*>(ob + (is_tail ? (u << 4) : ((u ^ r7) << 4))) = w;
      }
    }
    // ---- residual: row-coalesced read-modify-write of the staged chunk
    if (ea.resid) {
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (sub_u >= units) continue;
        const int r = it * 4 + sub_r;
        uint8_t* sp = obuf + (is_tail ? r * row_bytes + (sub_u << 4) : r * 128 + ((sub_u ^ (r & 7)) << 4));
        uint4 o = *reinterpret_cast<uint4*>(sp);
        const uint4 rr = rres[it];
        const int nb = nglob + sub_u * (16 / ESZ);
        if constexpr (OUT_BF16) {
          // packed bf16x2 adds (HADD2.BF16): one instruction per two elements
          uint32_t ow[4] = {o.x, o.y, o.z, o.w};
          uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!full) {
              if (nb + 2 * i >= ea.N) rw[i] &= 0xffff0000u;
              if (nb + 2 * i + 1 >= ea.N) rw[i] &= 0x0000ffffu;
            }
            __nv_bfloat162 s2 = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&ow[i]), *reinterpret_cast<__nv_bfloat162*>(&rw[i]));
            ow[i] = *reinterpret_cast<uint32_t*>(&s2);
          }
          o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        } else {
          float f[4] = {__uint_as_float(o.x), __uint_as_float(o.y), __uint_as_float(o.z), __uint_as_float(o.w)};
          const float g[4] = {__uint_as_float(rr.x), __uint_as_float(rr.y), __uint_as_float(rr.z), __uint_as_float(rr.w)};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (full || nb + i < ea.N) f[i] += g[i];
          o = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
        }
        *reinterpret_cast<uint4*>(sp) = o;
      }
    }
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const CUtensorMap* m = is_tail ? tmDt : tmD;
      if (ea.TW) tc::tma_store_4d(m, obuf, nglob, c1, c2, c3);
      else tc::tma_store_2d(m, obuf, nglob, c1);
      tc::tma_store_commit();
    }
    ++st_seq;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmDt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2, stage_bytes = a_bytes + b_bytes;
  const long long ntiles = p.tiles_m * p.tiles_n;
  const int kblocks = p.K / BK;
  uint8_t* epi_smem = smem + kMainBudget;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmD);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull_bar[a], 1); tc::mbar_init(&tempty_bar[a], kEpiWarps); }
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
          tc::umma_commit(&empty_bar[s]);                          // smem slot free when these MMAs retire
          if (kb == kblocks - 1) tc::umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;      // TMEM lane quarter this warp may access == its 32 rows of the tile
    const int half = ew >> 2;    // which of the two warps sharing that quarter (takes chunks half, half+2, ...)
    uint8_t* stage = epi_smem + ew * kEpiBytesPerWarp;
    uint32_t tl = 0, st_seq = 0;
    const EpiArgs ea = {p.M, p.ldr, p.bias, p.resid, p.alpha, p.N, p.BN, p.TW};
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      const TileCoord tc_ = tile_coord(p, t);
      tc::mbar_wait(&tfull_bar[acc], aph);
      tc::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);
#define ISP_EPI(OB, A) epilogue_tile<OB, A>(ea, &tmD, &tmDt, tc_, t_addr, stage, st_seq, q, half, lane)
      if (p.out_bf16) {
        if (p.act == 0) ISP_EPI(true, 0); else if (p.act == 1) ISP_EPI(true, 1);
        else if (p.act == 2) ISP_EPI(true, 2); else if (p.act == 3) ISP_EPI(true, 3); else ISP_EPI(true, 4);
      } else {
        if (p.act == 0) ISP_EPI(false, 0); else if (p.act == 1) ISP_EPI(false, 1);
        else if (p.act == 2) ISP_EPI(false, 2); else if (p.act == 3) ISP_EPI(false, 3); else ISP_EPI(false, 4);
      }
#undef ISP_EPI
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
    }
    if (lane == 0) tc::tma_store_wait_all();  // global writes complete before the CTA retires
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

static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmDt,
                  Params& p, cudaStream_t stream) {
  static int num_sms = 0;
  static bool attr_set = false;
  if (!num_sms) {
    int dev = 0;
    ISP_CUDA(cudaGetDevice(&dev));
    ISP_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!attr_set) {
    ISP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int stage_bytes = BM * BK * 2 + p.BN * BK * 2;
  p.stages = kMainBudget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  ISP_REQUIRE(p.stages >= 2, ISP_ERR_UNSUPPORTED, "gemm_tc: tile too large for a 2-stage pipeline");
  const long long ntiles = p.tiles_m * p.tiles_n;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  gemm_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmA, tmB, tmD, tmDt, p);
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
  const int esz = out_bf16 ? 2 : 4;
  ISP_REQUIRE((ldd * esz) % 16 == 0 && aligned16(D), ISP_ERR_MISALIGNED,
              "gemm_bf16_tc: D rows must be 16-byte aligned (ldd=%lld)", ldd);
  ISP_REQUIRE(aligned16(A) && aligned16(W), ISP_ERR_MISALIGNED, "gemm_bf16_tc: A/W must be 16-byte aligned");
  ISP_REQUIRE(act >= 0 && act <= 4, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: unknown activation %d", act);
  ISP_REQUIRE(!bias || aligned16(bias), ISP_ERR_MISALIGNED, "gemm_bf16_tc: bias must be 16-byte aligned");
  ISP_REQUIRE(M < (1ll << 31), ISP_ERR_UNSUPPORTED, "gemm_bf16_tc: M too large for TMA coordinates");
  if (resid) {
    ISP_REQUIRE((resid_bf16 != 0) == (out_bf16 != 0), ISP_ERR_UNSUPPORTED,
                "gemm_bf16_tc: residual must have the output dtype");
    ISP_REQUIRE(ldr >= N && (ldr * esz) % 16 == 0 && aligned16(resid), ISP_ERR_MISALIGNED,
                "gemm_bf16_tc: residual rows must be 16-byte aligned (ldr=%lld)", ldr);
  }
  gemm::Params p = {};
  p.M = M; p.N = N;
  p.K = (K + gemm::BK - 1) / gemm::BK * gemm::BK;  // TMA zero-fills the K tail (global dim = K)
  p.BN = gemm::pick_bn(N);
  p.TW = 0;
  p.tiles_m = (M + gemm::BM - 1) / gemm::BM;
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.bias = bias; p.resid = resid; p.ldr = ldr; p.alpha = alpha; p.act = act; p.out_bf16 = out_bf16;
  CUtensorMap tmA, tmB, tmD;
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
  const uint32_t cw = out_bf16 ? 64 : 32;
  const uint32_t tailw = (uint32_t)p.BN % cw;  // narrower last chunk of every tile (0 = none)
  CUtensorMap tmDt;
  {
    const uint64_t dims[2] = {(uint64_t)ldd, (uint64_t)M}, str[2] = {(uint64_t)esz, (uint64_t)ldd * esz};
    const uint32_t box[2] = {cw, 32}, boxt[2] = {tailw ? tailw : cw, 32};
    if (int e = make_tmap(&tmD, esz, D, 2, dims, str, box, "gemm_bf16_tc(D)", true)) return e;
    if (int e = make_tmap(&tmDt, esz, D, 2, dims, str, boxt, "gemm_bf16_tc(D tail)", false)) return e;
  }
  return gemm::launch(tmA, tmB, tmD, tmDt, p, as_stream(stream));
}

extern "C" int isp_conv3x3_bf16_tc(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                                   int Nimg, int H, int Wd, int Cin, int ldx, int Cout, int ldy, isp_stream_t stream) {
  ISP_REQUIRE(X && Wp && Y, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: null pointer");
  ISP_REQUIRE(Nimg > 0 && H > 0 && Wd > 0 && Cout > 0 && Cin > 0, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: bad shape");
  ISP_REQUIRE(ldx >= Cin && ldx % 8 == 0, ISP_ERR_MISALIGNED,
              "conv3x3_bf16_tc: channel stride must be >= Cin and a multiple of 8 (got %d)", ldx);
  const int esz = out_bf16 ? 2 : 4;
  ISP_REQUIRE(ldy >= Cout && (ldy * esz) % 16 == 0 && aligned16(Y), ISP_ERR_MISALIGNED,
              "conv3x3_bf16_tc: output pixels must be 16-byte aligned (ldy=%d)", ldy);
  const int Cin_pad = (Cin + 63) / 64 * 64;  // weights are packed [Cout][9][Cin_pad]; TMA zero-fills c >= Cin
  ISP_REQUIRE(aligned16(X) && aligned16(Wp), ISP_ERR_MISALIGNED, "conv3x3_bf16_tc: X/W must be 16-byte aligned");
  ISP_REQUIRE(act >= 0 && act <= 4, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc: unknown activation %d", act);
  ISP_REQUIRE(!bias || aligned16(bias), ISP_ERR_MISALIGNED, "conv3x3_bf16_tc: bias must be 16-byte aligned");
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
  p.bias = bias; p.resid = nullptr; p.ldr = 0; p.alpha = 1.f; p.act = act; p.out_bf16 = out_bf16;
  CUtensorMap tmA, tmB, tmD, tmDt;
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
  {
    const uint32_t bw = TW < 32 ? TW : 32;
    const uint64_t dims[4] = {(uint64_t)ldy, (uint64_t)Wd, (uint64_t)H, (uint64_t)Nimg};
    const uint64_t str[4] = {(uint64_t)esz, (uint64_t)ldy * esz, (uint64_t)Wd * ldy * esz, (uint64_t)H * Wd * ldy * esz};
    const uint32_t cw = out_bf16 ? 64u : 32u, tailw = (uint32_t)p.BN % cw;
    const uint32_t box[4] = {cw, bw, 32 / bw, 1}, boxt[4] = {tailw ? tailw : cw, bw, 32 / bw, 1};
    if (int e = make_tmap(&tmD, esz, Y, 4, dims, str, box, "conv3x3_bf16_tc(Y)", true)) return e;
    if (int e = make_tmap(&tmDt, esz, Y, 4, dims, str, boxt, "conv3x3_bf16_tc(Y tail)", false)) return e;
  }
  return gemm::launch(tmA, tmB, tmD, tmDt, p, as_stream(stream));
}
