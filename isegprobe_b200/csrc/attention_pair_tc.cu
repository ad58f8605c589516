// Flash-style attention on tcgen05, TWO query tiles per CTA (sm_100a): the LoftUp cross-attention kernel
// (200 704 pixel queries x 1024 low-res keys, 4 heads x 101; nn.MultiheadAttention slow path in the reference,
// loftup/layers.py:186-202).
//
// Why two tiles: the one-tile kernel (attention_tc.cu) streams every K / V^T block (56 KB) from L2 once per 128-query tile;
// with all softmax arithmetic removed it still takes 1.15 of its 1.5 ms (ISP_ATTN_DEBUG=2 probe, round 2) -- 13.2 GB of
// operand ingest at 11.5 TB/s, the L2->SM fabric ceiling the conv / GEMM kernels also sit on (12.3 TB/s).
// Here a K / V^T block is loaded once for 256 queries, which halves that traffic.  One thread owns a whole query row (no
// row-max exchange, no split accumulators), and the scores are produced in SUB-blocks of 64 keys so that every tile has TWO
// 64-column score buffers: S_X(u+2) is issued as soon as P_X(u) V has been, i.e. the softmax warps always find the next
// sub-block's scores waiting and never stall on the tensor pipe (a first version with one 128-key buffer per tile spent 38 %
// of the softmax warps' time in that wait: 1.35 ms against 1.56 ms for the one-tile kernel).
//
// TMEM columns: tile A scores @0 / @64, tile B @128 / @192 (fp32; the bf16 probabilities are written back over the first 32
// columns of the same buffer and fed to tcgen05.mma as the A operand), O_A @256, O_B @256+DV.
// Work item = (pair of adjacent 128-query tiles, head).  Q is expected pre-scaled by 1/sqrt(head_dim).
#include "tc_common.cuh"

namespace isp {
namespace attn2 {

constexpr int kSoftmaxWarps = 8;  // 4 per query tile, one thread per query row
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr int BQ = 128, BKEY = 128, SUB = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThresh = 8.0f;  // log2 units: P stays <= 2^8

template <int DK_CHUNKS, int DV>
struct Plan {
  static constexpr uint32_t kQTile = DK_CHUNKS * 16384;
  static constexpr uint32_t kKStage = DK_CHUNKS * 16384;
  static constexpr uint32_t kVChunk = (DV * 128 + 1023) / 1024 * 1024;  // chunk bases stay 1024-byte aligned (swizzle)
  static constexpr uint32_t kVStage = 2 * kVChunk;
  static constexpr uint32_t kOBytes = BQ * DV * 2;  // ONE staging tile, used by the two query tiles in turn
  static constexpr uint32_t kBytes = 2 * kQTile + 2 * kKStage + 2 * kVStage + kOBytes;
};

struct Params {
  int nkeys, nblocks;      // real keys, blocks of 128 (K / V^T padded with zeros)
  int heads;
  long long rows_per_img;  // queries per image
  int pairs_per_img;       // ceil(tiles / 2)
  long long nitems;        // B * pairs_per_img * heads
  int q_head_stride, o_head_stride;
  float* lse;              // optional [B][heads][rows_per_img] (log2 units)
  int lsum_col;            // LSUM: O column whose V^T row is all ones
#ifdef ISP_ATTN_PROBE
  int debug_mode;          // timing probe (ISP_ATTN_DEBUG=2): TMEM ld / st and the barrier chain without the softmax arithmetic
#endif
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes: Cody-Waite split + degree-3 minimax polynomial (max relative error 1.0e-4, P is rounded to
// bf16 right after); the power of two goes straight into the exponent field.  x is clamped at -126 (masked keys are -inf).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = __fadd_rn(x, 12582912.f);
  const float f = __fsub_rn(x, __fsub_rn(t, 12582912.f));  // in [-0.5, 0.5]
  float p = fmaf(f, 0.055008713f, 0.24221069f);
  p = fmaf(p, f, 0.6932829f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Per item, key blocks j = 0 .. nb-1 (128 keys, one K / V^T smem stage each), sub-blocks u = 0 .. 2 nb - 1 (64 keys),
// tiles X in {A, B}:
//   warp 0 (TMA)  : both Q tiles once per item; K block [128 keys x DK], V^T block [DV x 128 keys] -> 2-deep rings,
//                   each block feeding BOTH tiles.
//   warp 1 (MMA)  : S_A(0), S_B(0), S_A(1), S_B(1); then per sub-block and tile  [P_X(u) ready]  O_X += P_X(u) V(u),
//                   S_X(u+2) = Q_X K(u+2)^T into the buffer P_X(u) occupied.  tcgen05.mma executes in issue order, which
//                   orders the read of P_X(u) before S_X(u+2) overwrites it.
//   warps 2-5 / 6-9 : softmax of tile A / B, one thread per query row: tcgen05.ld the 64 scores, online softmax in fp32
//                   (exp2; O in TMEM is rescaled only when the row max grew by more than 2^8), P -> bf16 -> tcgen05.st;
//                   after the last sub-block O / l -> bf16 -> smem -> one TMA store per tile.
// LSUM: V^T row p.lsum_col is all ones, so the row sum of the bf16 probabilities accumulates in that O column on the tensor
// pipe (no FADDs, and the normalisation uses exactly the rounded probabilities).  POLY of every 8 exponentials run on the
// FMA pipe instead of MUFU (16 ex2 / clk / SM: 1024 clk per 128x128 block against 896 clk of tensor work).
template <int DK_CHUNKS, int DK_STEPS, int DV, bool LSUM, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t q_full, q_empty, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2][2],
      p_full[2][2], pv_done[2][2], o_free[2];  // s_full / p_full / pv_done: [tile][score buffer = sub-block parity]
  __shared__ uint32_t tmem_base_s;
  static_assert(256 + 2 * DV <= 512, "TMEM: two score buffers and two accumulators");

  using PL = Plan<DK_CHUNKS, DV>;
  constexpr uint32_t kQTile = PL::kQTile, kKStage = PL::kKStage, kVStage = PL::kVStage, kVChunk = PL::kVChunk;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * kQTile;
  uint8_t* sV = sK + 2 * kKStage;
  uint8_t* sO = sV + 2 * kVStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nblocks;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV); tc::prefetch_tmap(&tmO);
    tc::mbar_init(&q_full, 1); tc::mbar_init(&q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&k_full[i], 1); tc::mbar_init(&k_empty[i], 1);
      tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1);
      for (int b2 = 0; b2 < 2; ++b2) {
        tc::mbar_init(&s_full[i][b2], 1); tc::mbar_init(&p_full[i][b2], kSoftmaxWarps / 2);
        tc::mbar_init(&pv_done[i][b2], 1);
      }
      tc::mbar_init(&o_free[i], 1);
    }
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
      uint32_t kv_it = 0, item_it = 0;
      for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
        const int h = (int)(it % p.heads);
        const long long tq = it / p.heads;
        const int b = (int)(tq / p.pairs_per_img);
        const int pr = (int)(tq % p.pairs_per_img);
        const long long row0 = (long long)b * p.rows_per_img + (long long)pr * (2 * BQ);
        tc::mbar_wait(&q_empty, (item_it & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&q_full, 2 * kQTile);
        for (int x = 0; x < 2; ++x)
          for (int c = 0; c < DK_CHUNKS; ++c)
            tc::tma_load_2d(sQ + x * kQTile + c * 16384, &tmQ, &q_full, h * p.q_head_stride + c * 64,
                            (int)(row0 + x * BQ));
        const long long bh = (long long)b * p.heads + h;
        for (int j = 0; j < nb; ++j, ++kv_it) {
          const int s = kv_it & 1;
          const uint32_t ph = (kv_it >> 1) & 1;
          tc::mbar_wait(&k_empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&k_full[s], DK_CHUNKS * 16384);
          for (int c = 0; c < DK_CHUNKS; ++c)
            tc::tma_load_2d(sK + s * kKStage + c * 16384, &tmK, &k_full[s], c * 64,
                            (int)(bh * (long long)(nb * BKEY) + j * BKEY));
          tc::mbar_wait(&v_empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&v_full[s], 2 * DV * 128);
          for (int c = 0; c < 2; ++c)
            tc::tma_load_2d(sV + s * kVStage + c * kVChunk, &tmV, &v_full[s], j * BKEY + c * 64, (int)(bh * DV));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_qk = tc::idesc_bf16_f32(BQ, SUB);
    const uint32_t idesc_pv = tc::idesc_bf16_f32(BQ, DV);
    const bool leader = tc::elect_one();
    const uint32_t q_lo = tc::smem_u32(sQ), k_lo = tc::smem_u32(sK), v_lo = tc::smem_u32(sV);
    uint32_t kv_it = 0, item_it = 0, gu = 0;  // gu: sub-blocks processed so far (all items)
    const int nu = 2 * nb;
    // S_X(u) = Q_X K(u)^T (64 keys: rows (u&1)*64.. of the K stage of block u/2), then signal tile X's softmax warps;
    // the last of a block's four products frees the K stage
    auto issue_qk = [&](int x, int u) {
      const uint32_t kvi = kv_it + (u >> 1);
      const int s = kvi & 1, hf = u & 1;
      if (x == 0 && hf == 0) {
        tc::mbar_wait(&k_full[s], (kvi >> 1) & 1);
        tc::tc_fence_after();
      }
      const uint32_t d = tmem + x * 128 + hf * SUB;
      const uint32_t qa = q_lo + x * kQTile, ka = k_lo + s * kKStage + hf * (SUB * 128);
#pragma unroll
      for (int k = 0; k < DK_STEPS; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        const uint64_t da = tc::smem_desc_k_sw128(qa + off), db = tc::smem_desc_k_sw128(ka + off);
        if (leader) tc::umma_bf16(d, da, db, idesc_qk, k ? 1u : 0u);
      }
      if (leader) {
        tc::umma_commit(&s_full[x][hf]);
        if (x == 1 && hf == 1) tc::umma_commit(&k_empty[s]);
        if (x == 1 && u == nu - 1) tc::umma_commit(&q_empty);  // last QK of the item: Q smem free when it retires
      }
    };
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      tc::mbar_wait(&q_full, item_it & 1);
      tc::tc_fence_after();
      issue_qk(0, 0); issue_qk(1, 0); issue_qk(0, 1); issue_qk(1, 1);
      for (int u = 0; u < nu; ++u, ++gu) {
        const uint32_t kvi = kv_it + (u >> 1);
        const int s = kvi & 1, hf = u & 1;
        const uint32_t va = v_lo + s * kVStage + hf * kVChunk;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          // O_X += P_X(u) V(u): P from TMEM (16 keys = 8 packed columns per MMA), V^T chunk (64 keys) from smem
          tc::mbar_wait(&p_full[x][hf], (gu >> 1) & 1);
          if (x == 0 && hf == 0) tc::mbar_wait(&v_full[s], (kvi >> 1) & 1);
          tc::tc_fence_after();
          const uint32_t dO = tmem + 256 + x * DV, aP = tmem + x * 128 + hf * SUB;
#pragma unroll
          for (int k = 0; k < SUB / 16; ++k) {
            const uint64_t db = tc::smem_desc_k_sw128(va + k * 32);
            if (leader) tc::umma_bf16_ts(dO, aP + k * 8, db, idesc_pv, (u | k) ? 1u : 0u);
          }
          if (leader) {
            tc::umma_commit(&pv_done[x][hf]);
            if (x == 1 && hf == 1) tc::umma_commit(&v_empty[s]);
          }
          if (u + 2 < nu) issue_qk(x, u + 2);
        }
      }
      kv_it += nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int sw = warp - 2;
    const int x = sw >> 2;       // query tile of the pair
    const int q = warp & 3;      // TMEM lane quarter this warp may access
    const int r = q * 32 + lane; // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tS0 = tmem + x * 128 + lane_addr;
    const uint32_t tO = tmem + 256 + x * DV + lane_addr;
    uint32_t gu = 0, item_it = 0;  // gu: sub-blocks processed so far (all items)
    const int nu = 2 * nb;
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      const int h = (int)(it % p.heads);
      const long long tq = it / p.heads;
      const int b = (int)(tq / p.pairs_per_img);
      const int qt = (int)(tq % p.pairs_per_img) * 2 + x;
      float m_ref = -INFINITY;  // exponent reference, log2 units
      float l = 0.f;
      for (int u = 0; u < nu; ++u, ++gu) {
        const int hf = u & 1;
        const uint32_t tS = tS0 + hf * SUB;
        tc::mbar_wait(&s_full[x][hf], (gu >> 1) & 1);
        tc::tc_fence_after();
        uint32_t sv[SUB];
        tc::tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tc::tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tc::tmem_ld_wait();
#ifdef ISP_ATTN_PROBE  // -DISP_ATTN_PROBE builds only: profiles/r02_attention_probes.txt
        if (p.debug_mode == 2) {
#pragma unroll
          for (int c = 0; c < SUB / 16; ++c) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pk[e] = sv[c * 16 + 2 * e] ^ sv[c * 16 + 2 * e + 1];
            tc::tmem_st8(tS + c * 8, pk);
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&p_full[x][hf]);
          continue;
        }
#endif
        const int kbase = u * SUB;
        if (kbase + SUB > p.nkeys) {  // only the last key block can hold padded keys (warp-uniform)
#pragma unroll
          for (int e = 0; e < SUB; ++e)
            if (kbase + e >= p.nkeys) sv[e] = __float_as_uint(-INFINITY);
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < SUB; e += 4) {
          mx[0] = fmaxf(mx[0], __uint_as_float(sv[e]));
          mx[1] = fmaxf(mx[1], __uint_as_float(sv[e + 1]));
          mx[2] = fmaxf(mx[2], __uint_as_float(sv[e + 2]));
          mx[3] = fmaxf(mx[3], __uint_as_float(sv[e + 3]));
        }
        const float mj = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * kLog2e;
        // lazy max update: only move the reference when the row max grew by more than 2^8
        float scale = 1.f;
        bool need = false;
        if (mj > m_ref + kRescaleThresh) {
          scale = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - mj);
          m_ref = mj;
          need = (u > 0);
        }
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale this row of O once P_X(u-1) V has retired.  S_X(u) being full proves that P_X(u-2) V has (it was
          // issued before S_X(u)), hence also P_X(u-3) V -- the previous phase of the barrier waited on -- so the parity
          // wait cannot alias.
          tc::mbar_wait(&pv_done[x][hf ^ 1], ((gu - 1) >> 1) & 1);
          tc::tc_fence_after();
          const float f = need ? scale : 1.f;
          for (int c = 0; c < DV; c += 16) {
            uint32_t o[16];
            tc::tmem_ld16(tO + c, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tc::tmem_st16(tO + c, o);
          }
        }
        const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < SUB / 16; ++c) {  // 16 keys -> 8 packed bf16x2 columns of P
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float x0 = fmaf(__uint_as_float(sv[c * 16 + e]), kLog2e, neg_m);
            const float x1 = fmaf(__uint_as_float(sv[c * 16 + e + 1]), kLog2e, neg_m);
            const float p0 = ((e & 7) < POLY) ? ex2_poly(x0) : ex2_approx(x0);
            const float p1 = (((e + 1) & 7) < POLY) ? ex2_poly(x1) : ex2_approx(x1);
            if constexpr (!LSUM) rs[(e >> 1) & 3] += p0 + p1;
            __nv_bfloat162 bb = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          tc::tmem_st8(tS + c * 8, pk);  // over the scores this thread has already read
        }
        if constexpr (!LSUM) l = l * scale + (rs[0] + rs[1]) + (rs[2] + rs[3]);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        // no double arrival in one phase: S_X(u+2) (same buffer, same barrier) is only issued after P_X(u) V, which waited
        // for all four arrivals of sub-block u
        if (lane == 0) tc::mbar_arrive(&p_full[x][hf]);
      }
      // epilogue: O / l -> bf16 -> staging tile (shared by the two query tiles in turn) -> one TMA store per tile
      // the last product P_X V (odd sub-block nu-1; products retire in order, so the even one before it has too).  The
      // previous phase of this barrier (sub-block nu-3) is complete -- S_X(nu-1) was issued after it -- so no aliasing.
      tc::mbar_wait(&pv_done[x][1], ((gu - 1) >> 1) & 1);
      tc::tc_fence_after();
      if constexpr (LSUM) {
        uint32_t lv[1];
        tc::tmem_ld1(tO + p.lsum_col, lv);
        tc::tmem_ld_wait();
        l = __uint_as_float(lv[0]);
      }
      const float inv = 1.f / l;
      const long long row_local = (long long)qt * BQ + r;
      if (p.lse && row_local < p.rows_per_img)
        p.lse[((long long)b * p.heads + h) * p.rows_per_img + row_local] = m_ref + log2f(l);
      // tile A waits until tile B's store of the previous item has read the staging tile, tile B until tile A's of this item
      if (x == 0) tc::mbar_wait(&o_free[1], (item_it & 1) ^ 1);
      else tc::mbar_wait(&o_free[0], item_it & 1);
      uint8_t* orow = sO + r * (DV * 2);
      for (int c = 0; c < DV; c += 16) {
        uint32_t o[16];
        tc::tmem_ld16(tO + c, o);
        tc::tmem_ld_wait();
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(o[2 * e]) * inv, __uint_as_float(o[2 * e + 1]) * inv);
          w[e] = *reinterpret_cast<uint32_t*>(&bb);
        }
        *reinterpret_cast<uint4*>(orow + c * 2) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(orow + c * 2 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
      }
      // the next item's first P_X V overwrites O_X only after this tile's next p_full arrivals, which follow these reads
      tc::tc_fence_before();
      tc::fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(10 + x) : "memory");  // the four warps of this tile
      if ((sw & 3) == 0 && lane == 0) {
        if ((long long)qt * BQ < p.rows_per_img) {  // clipped at the image's last row by the 3-D map
          tc::tma_store_3d(&tmO, sO, h * p.o_head_stride, qt * BQ, b);
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
        }
        tc::mbar_arrive(&o_free[x]);
      }
    }
    if ((sw & 3) == 0 && lane == 0) tc::tma_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace attn2

// Launch for the DV = 112 / DK = 128 geometry (variant 1 of isp_attention_bf16_tc*); same operand layouts.
int attention_pair_launch(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt, void* out,
                          long long ldo, int o_head_stride, int B, long long rows_per_img, int heads, int nkeys, float* lse,
                          int lsum_col, int poly, isp_stream_t stream) {
  constexpr int DV = 112, DKC = 128;
  attn2::Params p = {};
  p.nkeys = nkeys;
  p.nblocks = (nkeys + attn2::BKEY - 1) / attn2::BKEY;
  p.heads = heads;
  p.rows_per_img = rows_per_img;
  const int tiles = (int)((rows_per_img + attn2::BQ - 1) / attn2::BQ);
  p.pairs_per_img = (tiles + 1) / 2;
  p.nitems = (long long)B * p.pairs_per_img * heads;
  p.q_head_stride = q_head_stride;
  p.o_head_stride = o_head_stride;
  p.lse = lse;
  p.lsum_col = lsum_col;
#ifdef ISP_ATTN_PROBE
  {
    const char* dbg = getenv("ISP_ATTN_DEBUG");
    p.debug_mode = dbg ? atoi(dbg) : 0;
  }
#endif
  const long long nkp = (long long)p.nblocks * attn2::BKEY;
  CUtensorMap tmQ, tmK, tmV, tmO;
  {  // out viewed as [B][rows_per_img][ldo]: a tile's store is clipped at its own image's last row
    const uint64_t dims[3] = {(uint64_t)ldo, (uint64_t)rows_per_img, (uint64_t)B};
    const uint64_t str[3] = {2, (uint64_t)ldo * 2, (uint64_t)rows_per_img * ldo * 2};
    const uint32_t box[3] = {(uint32_t)DV, attn2::BQ, 1};
    if (int e = make_tmap(&tmO, 2, out, 3, dims, str, box, "attention_pair(out)", false)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)ldq, (uint64_t)B * rows_per_img}, str[2] = {2, (uint64_t)ldq * 2};
    const uint32_t box[2] = {64, attn2::BQ};
    if (int e = make_tmap_bf16(&tmQ, Q, 2, dims, str, box, "attention_pair(Q)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)DKC, (uint64_t)B * heads * nkp}, str[2] = {2, (uint64_t)DKC * 2};
    const uint32_t box[2] = {64, attn2::BKEY};
    if (int e = make_tmap_bf16(&tmK, K, 2, dims, str, box, "attention_pair(K)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)nkp, (uint64_t)B * heads * DV}, str[2] = {2, (uint64_t)nkp * 2};
    const uint32_t box[2] = {64, (uint32_t)DV};
    if (int e = make_tmap_bf16(&tmV, Vt, 2, dims, str, box, "attention_pair(Vt)")) return e;
  }
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  const int smem = (int)attn2::Plan<2, DV>::kBytes;
#define ISP_ATTN2_ATTR(L, P_) \
  if (int e = ensure_dynamic_smem((const void*)attn2::attention_pair_kernel<2, 7, DV, L, P_>, smem)) return e
  ISP_ATTN2_ATTR(false, 0); ISP_ATTN2_ATTR(true, 0); ISP_ATTN2_ATTR(true, 2); ISP_ATTN2_ATTR(true, 3); ISP_ATTN2_ATTR(true, 4);
#undef ISP_ATTN2_ATTR
  const int grid = (int)(p.nitems < num_sms ? p.nitems : num_sms);
#define ISP_ATTN2_GO(L, P_) \
  attn2::attention_pair_kernel<2, 7, DV, L, P_><<<grid, attn2::kThreads, smem, as_stream(stream)>>>(tmQ, tmK, tmV, tmO, p)
  if (lsum_col < 0) ISP_ATTN2_GO(false, 0);
  else if (poly == 4) ISP_ATTN2_GO(true, 4);
  else if (poly == 3) ISP_ATTN2_GO(true, 3);
  else if (poly == 2) ISP_ATTN2_GO(true, 2);
  else ISP_ATTN2_GO(true, 0);
#undef ISP_ATTN2_GO
  ISP_CHECK_LAUNCH("attention_pair_kernel");
  return ISP_OK;
}

}  // namespace isp
