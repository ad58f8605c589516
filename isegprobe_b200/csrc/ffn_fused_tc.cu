// LoftUp FeedForward block in ONE kernel (sm_100a, tcgen05):
//
//     out = x + Linear2( GELU( Linear1( LayerNorm(x) ) ) )          loftup/layers.py:161-174 inside CATransformer (:222-228)
//
// for bf16 token rows x [M, ldx] (200 704 pixel queries per image, width 404 -> hidden 384 -> 404).  As two GEMM launches the
// block moved the [M, 384] hidden activations through HBM (616 MB written + read per 4 images) and spent 1.09 ms per 4 images
// at ~35 % of the tensor pipe (both GEMMs bound by their epilogue / store path); here the hidden tile of a 128-row block
// never leaves the SM:
//
//   warp 0 (TMA)  : phase 1 streams x [128 x 64] + W1 [384 x 64] k-chunks, phase 2 W2 [<=416 x 64] k-chunks, 2-deep ring
//   warp 1 (MMA)  : phase 1  acc[128 x 384] (TMEM cols 0..383) = x W1g^T        (two N = 192 MMAs per 16-wide k-step)
//                   phase 2  acc[128 x 416] (TMEM cols 0..415) = H W2^T         (two N = 208 MMAs per k-step, A = H in smem)
//   warps 2..17   : epilogue 1  acc -> fused LayerNorm (rstd * acc - rstd * mean * g[n]) + b1 -> GELU (tanh form, as the
//                   unfused path) -> bf16 -> H [128 x 384] in shared memory, K-major 128B-swizzled (the A operand of phase 2)
//                   epilogue 2  acc + b2 + x (residual re-read from global, L2-resident) -> bf16 -> global, and the row
//                   statistics (sum, sum of squares of the stored values) the next fused LayerNorm needs: 4 slots per row
//
// The LayerNorm is folded exactly as in isp_gemm_bf16_tc_ex: W1g = W1 * gamma (bf16), g[n] = sum_k W1g[n,k],
// b1' = W1 beta + b1, and (sum, sum of squares) of every x row come from the producer's statistics slots.
// TMEM: one 416-column accumulator region reused by both phases (512 columns would not hold both), so the two epilogues
// are not overlapped with MMAs of the same tile; the TMA warp keeps prefetching the next phase's weight chunks meanwhile.
// Measured at 802816 rows (4 images), with parts switched off: MMAs + TMA alone 0.445 ms (bound by the L2->SM fill of the
// weights, 779 KB per 128-row tile), + epilogue 1 0.58 ms, + epilogue 2 1.02 ms.  Epilogue 2 is the weak part: it reads the
// residual and writes the output straight from registers, one row per lane -- 32 different lines per memory instruction,
// which the LSU serves at a fraction of the coalesced rate (the GEMM kernel's staged TMA stores do not fit next to the
// 96 KB hidden tile).  Equal to the two GEMM launches in time, so it stays opt-in (LoftUpUpsampler.fuse_ffn).
#include "tc_common.cuh"

namespace isp {
namespace ffn {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 16;                       // four per TMEM lane quarter: each takes a quarter of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;       // 576
constexpr int kStages = 2;
constexpr uint32_t kXBytes = BM * BK * 2;           // 16 KB
constexpr uint32_t kWBytesMax = 416 * BK * 2;       // 52 KB
constexpr uint32_t kStageBytes = 64 * 1024;         // x chunk + W chunk
constexpr int kHidMax = 384;
constexpr uint32_t kHBytes = BM * kHidMax * 2;      // 96 KB
constexpr uint32_t kSmem = kStages * kStageBytes + kHBytes;  // 224 KB

struct Params {
  long long M;
  int K1;        // real width of x (404)
  int k1chunks;  // ceil(ldx / 64)
  int NH;        // hidden width (<= 384, multiple of 64)
  int N2;        // output width (<= 416)
  int n2half;    // MMA N of phase 2: two of them cover round_up(N2, 16) (multiple of 16, <= 208)
  long long ldx, ldo;
  const __nv_bfloat16* x;   // residual (same tensor as the A operand)
  __nv_bfloat16* out;
  const float* ln_stats;    // [M][ln_slots][2]
  int ln_slots;
  float ln_invC, ln_eps;
  const float* g1;          // [NH]
  const float* b1;          // [NH]
  const float* b2;          // [N2]
  float* stats_out;         // [M][4][2] or null
  long long ntiles;
};

__device__ __forceinline__ float gelu_tanh(float v) {  // act == 4 of isp_gemm_bf16_tc
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(v * fmaf(v * v, 0.0356774081f, 0.7978845608f)));
  const float hv = 0.5f * v;
  return fmaf(hv, th, hv);
}

__global__ void __launch_bounds__(kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc1_full, h_ready, acc2_full, acc_free;
  __shared__ uint32_t tmem_base_s;
  uint8_t* ring = smem;
  uint8_t* sH = smem + kStages * kStageBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k2chunks = p.NH / BK;
  const uint32_t w1_bytes = (uint32_t)p.NH * BK * 2;
  const uint32_t w2_bytes = 2u * (uint32_t)p.n2half * BK * 2;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmX); tc::prefetch_tmap(&tmW1); tc::prefetch_tmap(&tmW2);
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(&acc1_full, 1); tc::mbar_init(&acc2_full, 1);
    tc::mbar_init(&h_ready, kEpiWarps); tc::mbar_init(&acc_free, kEpiWarps);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t n = 0;
      for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int m0 = (int)(t * BM);
        for (int kc = 0; kc < p.k1chunks; ++kc, ++n) {
          const int s = n % kStages;
          tc::mbar_wait(&empty_bar[s], ((n / kStages) & 1) ^ 1);
          uint8_t* st = ring + s * kStageBytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], kXBytes + w1_bytes);
          tc::tma_load_2d(st, &tmX, &full_bar[s], kc * BK, m0);
          tc::tma_load_2d(st + kXBytes, &tmW1, &full_bar[s], kc * BK, 0);
          tc::tma_load_2d(st + kXBytes + (p.NH / 2) * BK * 2, &tmW1, &full_bar[s], kc * BK, p.NH / 2);
        }
        for (int kc = 0; kc < k2chunks; ++kc, ++n) {
          const int s = n % kStages;
          tc::mbar_wait(&empty_bar[s], ((n / kStages) & 1) ^ 1);
          uint8_t* st = ring + s * kStageBytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], w2_bytes);
          tc::tma_load_2d(st, &tmW2, &full_bar[s], kc * BK, 0);
          tc::tma_load_2d(st + p.n2half * BK * 2, &tmW2, &full_bar[s], kc * BK, p.n2half);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tc::idesc_bf16_f32(BM, p.NH / 2);
    const uint32_t idesc2 = tc::idesc_bf16_f32(BM, p.n2half);
    const bool leader = tc::elect_one();
    const uint32_t ring_lo = tc::smem_u32(ring), h_lo = tc::smem_u32(sH);
    const uint32_t w1_half = (uint32_t)(p.NH / 2) * BK * 2, w2_half = (uint32_t)p.n2half * BK * 2;
    const int k1steps = (p.K1 + 15) / 16;  // 16-wide MMA k-steps that hold real columns of x
    uint32_t n = 0, it = 0;
    for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
      tc::mbar_wait(&acc_free, (it & 1) ^ 1);  // the previous tile's epilogue 2 has drained the accumulator
      tc::tc_fence_after();
      for (int kc = 0; kc < p.k1chunks; ++kc, ++n) {
        const int s = n % kStages;
        tc::mbar_wait(&full_bar[s], (n / kStages) & 1);
        tc::tc_fence_after();
        const uint32_t xa = ring_lo + s * kStageBytes, wa = xa + kXBytes;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (kc * (BK / 16) + k < k1steps) {
            const uint64_t da = tc::smem_desc_k_sw128(xa + k * 32);
            const uint64_t db0 = tc::smem_desc_k_sw128(wa + k * 32), db1 = tc::smem_desc_k_sw128(wa + w1_half + k * 32);
            if (leader) {
              tc::umma_bf16(tmem, da, db0, idesc1, (kc | k) ? 1u : 0u);
              tc::umma_bf16(tmem + p.NH / 2, da, db1, idesc1, (kc | k) ? 1u : 0u);
            }
          }
        }
        if (leader) tc::umma_commit(&empty_bar[s]);
      }
      if (leader) tc::umma_commit(&acc1_full);
      tc::mbar_wait(&h_ready, it & 1);  // H is in shared memory and the accumulator has been read
      tc::tc_fence_after();
      for (int kc = 0; kc < k2chunks; ++kc, ++n) {
        const int s = n % kStages;
        tc::mbar_wait(&full_bar[s], (n / kStages) & 1);
        tc::tc_fence_after();
        const uint32_t wa = ring_lo + s * kStageBytes, ha = h_lo + kc * (BM * BK * 2);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = tc::smem_desc_k_sw128(ha + k * 32);
          const uint64_t db0 = tc::smem_desc_k_sw128(wa + k * 32), db1 = tc::smem_desc_k_sw128(wa + w2_half + k * 32);
          if (leader) {
            tc::umma_bf16(tmem, da, db0, idesc2, (kc | k) ? 1u : 0u);
            tc::umma_bf16(tmem + p.n2half, da, db1, idesc2, (kc | k) ? 1u : 0u);
          }
        }
        if (leader) tc::umma_commit(&empty_bar[s]);
      }
      if (leader) tc::umma_commit(&acc2_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter
    const int part = ew >> 2;        // which quarter of the columns
    const int row = q * 32 + lane;   // row of the tile == TMEM lane
    const uint32_t t_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int r7 = row & 7;
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
      const long long grow = t * BM + row;
      const bool row_ok = grow < p.M;
      // LayerNorm statistics of this x row (fixed summation order: deterministic)
      float ln_rstd = 1.f, ln_nrm = 0.f;
      if (row_ok) {
        float su = 0.f, sq = 0.f;
        const float2* sp2 = reinterpret_cast<const float2*>(p.ln_stats) + grow * p.ln_slots;
        for (int k = 0; k < p.ln_slots; ++k) {
          const float2 v = __ldg(sp2 + k);
          su += v.x;
          sq += v.y;
        }
        const float mean = su * p.ln_invC;
        ln_rstd = rsqrtf(fmaxf(sq * p.ln_invC - mean * mean, 0.f) + p.ln_eps);
        ln_nrm = -ln_rstd * mean;
      }
      // ---- epilogue 1: hidden tile -> shared memory
      tc::mbar_wait(&acc1_full, it & 1);
      tc::tc_fence_after();
      const int nh4 = p.NH / 4;  // columns per part (multiple of 16)
      for (int c0 = part * nh4; c0 < (part + 1) * nh4; c0 += 16) {
        uint32_t vr[16];
        tc::tmem_ld16(t_addr + c0, vr);
        tc::tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(p.g1 + c0 + e));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + c0 + e));
          const float a0 = fmaf(ln_rstd, __uint_as_float(vr[e]), fmaf(ln_nrm, gg.x, bb.x));
          const float a1 = fmaf(ln_rstd, __uint_as_float(vr[e + 1]), fmaf(ln_nrm, gg.y, bb.y));
          const float a2 = fmaf(ln_rstd, __uint_as_float(vr[e + 2]), fmaf(ln_nrm, gg.z, bb.z));
          const float a3 = fmaf(ln_rstd, __uint_as_float(vr[e + 3]), fmaf(ln_nrm, gg.w, bb.w));
          __nv_bfloat162 h0 = __floats2bfloat162_rn(gelu_tanh(a0), gelu_tanh(a1));
          __nv_bfloat162 h1 = __floats2bfloat162_rn(gelu_tanh(a2), gelu_tanh(a3));
          w[e >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          w[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        // H chunk (64 hidden columns) = [128 rows x 128 B], 16-byte units XOR-swizzled by the row (what TMA would write)
        uint8_t* hrow = sH + (c0 >> 6) * (BM * BK * 2) + row * 128;
        const int u0 = (c0 & 63) >> 3;
        *reinterpret_cast<uint4*>(hrow + (((u0) ^ r7) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(hrow + (((u0 + 1) ^ r7) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&h_ready);
      // ---- epilogue 2: + bias + residual -> global, row statistics.  The residual / bias of a 16-column group are
      // loaded one group ahead (the first one before the accumulator is even complete): their L2 latency would otherwise
      // be exposed seven times per row (ncu: 70 % of the stall samples were long-scoreboard waits in this loop).
      const int n2p = 2 * p.n2half;                 // accumulator columns (multiple of 32)
      const int per = ((n2p / 4 + 15) / 16) * 16;    // columns per part
      const int cb = part * per, ce = min(n2p, cb + per);
      float st_sum = 0.f, st_sq = 0.f;
      const __nv_bfloat16* xrow = p.x + grow * p.ldx;
      __nv_bfloat16* orow = p.out + grow * p.ldo;
      uint4 nr0 = make_uint4(0, 0, 0, 0), nr1 = nr0;
      float4 nb[4];
      auto prefetch = [&](int c0) {
        if (row_ok && c0 < p.ldo) {
          nr0 = __ldg(reinterpret_cast<const uint4*>(xrow + c0));
          nr1 = __ldg(reinterpret_cast<const uint4*>(xrow + c0 + 8));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)  // b2 is padded to a multiple of 16 floats by the caller's packer
          nb[k] = c0 + 4 * k < p.N2 ? __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + 4 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (cb < ce) prefetch(cb);
      tc::mbar_wait(&acc2_full, it & 1);
      tc::tc_fence_after();
      for (int c0 = cb; c0 < ce; c0 += 16) {
        uint32_t vr[16];
        tc::tmem_ld16(t_addr + c0, vr);
        const uint32_t rw[8] = {nr0.x, nr0.y, nr0.z, nr0.w, nr1.x, nr1.y, nr1.z, nr1.w};
        const float bv[16] = {nb[0].x, nb[0].y, nb[0].z, nb[0].w, nb[1].x, nb[1].y, nb[1].z, nb[1].w,
                              nb[2].x, nb[2].y, nb[2].z, nb[2].w, nb[3].x, nb[3].y, nb[3].z, nb[3].w};
        if (c0 + 16 < ce) prefetch(c0 + 16);
        tc::tmem_ld_wait();
        if (row_ok && c0 < p.ldo) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const int n = c0 + e;
            float v0 = 0.f, v1 = 0.f;
            if (n < p.N2) v0 = __uint_as_float(vr[e]) + bv[e] + __uint_as_float(rw[e >> 1] << 16);
            if (n + 1 < p.N2) v1 = __uint_as_float(vr[e + 1]) + bv[e + 1] + __uint_as_float(rw[e >> 1] & 0xffff0000u);
            __nv_bfloat162 ob = __floats2bfloat162_rn(v0, v1);
            w[e >> 1] = *reinterpret_cast<uint32_t*>(&ob);
            const float lo = __uint_as_float(w[e >> 1] << 16), hi = __uint_as_float(w[e >> 1] & 0xffff0000u);
            st_sum += lo + hi;
            st_sq = fmaf(lo, lo, fmaf(hi, hi, st_sq));
          }
          *reinterpret_cast<uint4*>(orow + c0) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      if (p.stats_out && row_ok) reinterpret_cast<float2*>(p.stats_out)[grow * 4 + part] = make_float2(st_sum, st_sq);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_free);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace ffn
}  // namespace isp

using namespace isp;

// out[M, ldo] = x + (GELU(LN(x) W1^T + b1)) W2^T + b2 with the LayerNorm folded (W1g, g1, b1 from the host packer
// tc.pack_ln_linear).  x bf16 [M, ldx] with K1 real columns (padding columns zero); W1g bf16 [NH, ldw1 >= K1];
// W2 bf16 [N2, ldw2 >= NH]; b2 fp32 with round_up(N2, 4) readable floats, 16-byte aligned; ln_stats fp32 [M][ln_slots][2];
// stats_out fp32 [M][4][2] or NULL.
extern "C" int isp_ffn_fused_bf16_tc(const void* x, long long ldx, int K1, const void* W1g, long long ldw1, const float* g1,
                                     const float* b1, int NH, const void* W2, long long ldw2, const float* b2, int N2,
                                     void* out, long long ldo, long long M, const float* ln_stats, int ln_slots,
                                     float ln_eps, float* stats_out, isp_stream_t stream) {
  ISP_REQUIRE(x && W1g && g1 && b1 && W2 && b2 && out && ln_stats, ISP_ERR_BAD_SHAPE, "ffn_fused_bf16_tc: null pointer");
  ISP_REQUIRE(M > 0 && M < (1ll << 31) && K1 > 0 && ln_slots > 0, ISP_ERR_BAD_SHAPE, "ffn_fused_bf16_tc: bad shape");
  ISP_REQUIRE(NH > 0 && NH <= ffn::kHidMax && NH % 64 == 0, ISP_ERR_UNSUPPORTED,
              "ffn_fused_bf16_tc: hidden width %d must be a multiple of 64, <= %d (TMEM / shared-memory tile)", NH, ffn::kHidMax);
  ISP_REQUIRE(N2 > 0 && N2 <= 416 && N2 % 4 == 0, ISP_ERR_UNSUPPORTED,
              "ffn_fused_bf16_tc: output width %d must be a multiple of 4, <= 416 (TMEM accumulator)", N2);
  ISP_REQUIRE(ldx >= K1 && ldx % 8 == 0 && ldx <= 448 && ldw1 >= K1 && ldw1 % 8 == 0 && ldw2 >= NH && ldw2 % 8 == 0,
              ISP_ERR_BAD_SHAPE, "ffn_fused_bf16_tc: leading dimensions (multiples of 8; ldx <= 448)");
  ISP_REQUIRE(ldo >= N2 && ldo % 8 == 0 && ldo <= (N2 + 15) / 16 * 16 && ldx >= ldo, ISP_ERR_BAD_SHAPE,
              "ffn_fused_bf16_tc: N2 <= ldo <= round_up(N2, 16) (padding columns are zero-filled) and ldx >= ldo (residual)");
  ISP_REQUIRE(aligned16(x) && aligned16(W1g) && aligned16(W2) && aligned16(out) && aligned16(g1) && aligned16(b1) && aligned16(b2),
              ISP_ERR_MISALIGNED,
              "ffn_fused_bf16_tc: 16-byte alignment");
  ffn::Params p = {};
  p.M = M; p.K1 = K1; p.k1chunks = (int)((ldx + 63) / 64); p.NH = NH; p.N2 = N2;
  p.n2half = ((N2 + 31) / 32) * 16;  // two halves, each a multiple of 16
  p.ldx = ldx; p.ldo = ldo;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ln_stats = ln_stats; p.ln_slots = ln_slots; p.ln_invC = 1.f / (float)K1; p.ln_eps = ln_eps;
  p.g1 = g1; p.b1 = b1; p.b2 = b2; p.stats_out = stats_out;
  p.ntiles = (M + ffn::BM - 1) / ffn::BM;
  ISP_REQUIRE(2 * p.n2half <= 416 && (uint32_t)(2 * p.n2half * 128) <= ffn::kStageBytes &&
                  ffn::kXBytes + (uint32_t)NH * 128 <= ffn::kStageBytes,
              ISP_ERR_UNSUPPORTED, "ffn_fused_bf16_tc: tile does not fit the shared-memory ring");
  CUtensorMap tmX, tmW1, tmW2;
  {
    const uint64_t dims[2] = {(uint64_t)ldx, (uint64_t)M}, str[2] = {2, (uint64_t)ldx * 2};
    const uint32_t box[2] = {ffn::BK, ffn::BM};
    if (int e = make_tmap_bf16(&tmX, x, 2, dims, str, box, "ffn_fused(x)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)ldw1, (uint64_t)NH}, str[2] = {2, (uint64_t)ldw1 * 2};
    const uint32_t box[2] = {ffn::BK, (uint32_t)NH / 2};
    if (int e = make_tmap_bf16(&tmW1, W1g, 2, dims, str, box, "ffn_fused(W1)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)ldw2, (uint64_t)N2}, str[2] = {2, (uint64_t)ldw2 * 2};
    const uint32_t box[2] = {ffn::BK, (uint32_t)p.n2half};
    if (int e = make_tmap_bf16(&tmW2, W2, 2, dims, str, box, "ffn_fused(W2)")) return e;
  }
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  if (int e = ensure_dynamic_smem((const void*)ffn::ffn_fused_kernel, (int)ffn::kSmem)) return e;
  const int grid = (int)(p.ntiles < num_sms ? p.ntiles : num_sms);
  ffn::ffn_fused_kernel<<<grid, ffn::kThreads, ffn::kSmem, as_stream(stream)>>>(tmX, tmW1, tmW2, p);
  ISP_CHECK_LAUNCH("ffn_fused_kernel");
  return ISP_OK;
}
