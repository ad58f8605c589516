// Flash-style attention on tcgen05 (sm_100a): O = softmax(Q K^T) V without ever
// writing the score matrix to HBM.  Used for
//   * LoftUp cross-attention: 200 704 pixel queries x 1024 low-res keys, 4 heads x 101
//     (nn.MultiheadAttention slow path in the reference, loftup/layers.py:186-202, which
//     materialises [B*4, HW, hw] fp32 = 3.29 GB / image / layer);
//   * DINOv2 self-attention: 1025 tokens, 6 heads x 64 (dinov2/layers/attention.py:54-71).
//
// Work item = (128-query tile, head); see the kernel comment for the pipeline.
// Q is expected pre-scaled by 1/sqrt(head_dim) (folded into the projection weights).
// Head dims are zero-padded on the K / V^T side only (DK_STEPS*16 >= head_dim, DV >= head_dim).
#include "tc_common.cuh"

namespace isp {
namespace attn {

constexpr int kSoftmaxWarps = 8;                 // two threads per query row: 64 keys of every block each
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr int BQ = 128, BKEY = 128;
// Shared-memory plan of one instantiation: Q tile = DK_CHUNKS 64-column chunks [128 rows x 64];
// K block likewise [128 keys x 64] per chunk, 2 stages; V^T block = two 64-key chunks [DV x 64], 2 stages;
// O staging tile [128 rows x DV] bf16 (dense) for the TMA store when it still fits (DV <= 128).
template <int DK_CHUNKS, int DV>
struct Plan {
  static constexpr uint32_t kQBytes = DK_CHUNKS * 16384;
  static constexpr uint32_t kKStage = DK_CHUNKS * 16384;
  static constexpr uint32_t kVChunk = (DV * 128 + 1023) / 1024 * 1024;  // chunk bases stay 1024-byte aligned (swizzle)
  static constexpr uint32_t kVStage = 2 * kVChunk;
  static constexpr bool kStageO = DV <= 128;
  static constexpr uint32_t kOBytes = kStageO ? BQ * DV * 2 : 0;
  static constexpr uint32_t kBytes = kQBytes + 2 * kKStage + 2 * kVStage + kOBytes;
};
constexpr uint32_t kSmemMin = 120 * 1024;        // requested at least: one CTA per SM (each CTA allocates all of TMEM)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThresh = 8.0f;           // log2 units: P stays <= 2^8

struct Params {
  int nkeys, nblocks;          // real keys, blocks of 128 (K / V^T padded with zeros)
  int heads;
  long long rows_per_img;      // queries per image (tile rows never straddle stored rows of 2 images)
  int tiles_per_img;
  long long nitems;            // B * tiles_per_img * heads
  int q_head_stride;           // column offset between heads in Q (elements)
  int o_head_stride;           // column offset between heads in out
  void* out;                   // bf16 [B*rows_per_img, ldo] (direct-store epilogue of the wide-head variant)
  long long ldo;
  int stagger_ns;              // split mode: the second half's warps start every tile this much later
  float* lse;                  // optional [B][heads][rows_per_img]: log2-domain log-sum-exp of every score row (for the backward)
};

__device__ __forceinline__ float ex2_approx(float x) {  // single MUFU.EX2 (ftz); inputs are <= 8
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Work item = (128-query tile, head).  Per item, for each block j of 128 keys:
//   warp 0 (TMA)  : Q once per item; K block [128 keys x DK], V^T block [DV x 128 keys] -> 2-deep rings
//   warp 1 (MMA)  : S(j) = Q K(j)^T into TMEM buffer j&1 -- issued TWO blocks ahead, so the softmax of
//                   block j never waits for the tensor pipe;  O += P(j) V(j) with P read FROM TMEM
//                   (tcgen05.mma A operand): probabilities never touch shared memory.
//                   All lanes run the (warp-uniform) loop, one elected lane issues: descriptors stay in
//                   uniform registers.
//   warps 2-9     : TWO threads per query row (the two warps sharing a TMEM lane quarter take 64 keys of
//                   the block each): tcgen05.ld half the S row once, row max exchanged through smem,
//                   online softmax in fp32 (exp2; O in TMEM is rescaled only when the row max grows by
//                   more than 2^8), P -> bf16 -> tcgen05.st over the first 64 columns of the S buffer
//                   (all S values of the row are in registers by then); after the last block
//                   O / l -> bf16 -> smem -> one TMA store per tile.
// TMEM columns: S buffers @0 and @128, O @256.  tcgen05.mma instructions execute in issue order, which
// orders PV(j) (reads P(j) in buffer j&1) before QK(j+2) (overwrites that buffer).
template <int DK_CHUNKS, int DK_STEPS, int DV>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t q_full, q_empty, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2],
      pv_done;
  // Split mode (TMEM has room for two O accumulators: 256 + 2*DV <= 512): the two threads of a query row are fully
  // independent -- each keeps its own running max / sum for ITS 64 keys of every block and its own accumulator
  // O_h += P[:, half] V[half] (two K=64 MMAs instead of one K=128); the halves are merged once per tile in the
  // epilogue.  No per-block max exchange and no barrier between the two warps that share an SM sub-partition, so
  // one of them can run its exponentials while the other waits for TMEM (measured: the kernel was bound by the
  // latency of that lock-step chain, not by MUFU -- removing every MUFU only took it from 1.60 to 1.42 ms).
  constexpr bool kSplit = 256 + 2 * DV <= 512;
  __shared__ uint32_t tmem_base_s;
  __shared__ float xchg[3][2][BQ];  // [slot][half][row]: slots 0/1 = block max (alternating), 2 = partial row sum

  using PL = Plan<DK_CHUNKS, DV>;
  constexpr uint32_t kKStage = PL::kKStage, kVStage = PL::kVStage, kVChunk = PL::kVChunk;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + PL::kQBytes;
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
      tc::mbar_init(&s_full[i], 1);
    }
    tc::mbar_init(&p_full[0], kSplit ? kSoftmaxWarps / 2 : kSoftmaxWarps);
    tc::mbar_init(&p_full[1], kSoftmaxWarps / 2);
    tc::mbar_init(&pv_done, 1);
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
        const int b = (int)(tq / p.tiles_per_img);
        const int qt = (int)(tq % p.tiles_per_img);
        const long long row0 = (long long)b * p.rows_per_img + (long long)qt * BQ;
        tc::mbar_wait(&q_empty, (item_it & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&q_full, DK_CHUNKS * 16384);
        for (int c = 0; c < DK_CHUNKS; ++c)
          tc::tma_load_2d(sQ + c * 16384, &tmQ, &q_full, h * p.q_head_stride + c * 64, (int)row0);
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
    const uint32_t idesc_qk = tc::idesc_bf16_f32(BQ, BKEY);
    const uint32_t idesc_pv = tc::idesc_bf16_f32(BQ, DV);
    const bool leader = tc::elect_one();
    const uint32_t q_lo = tc::smem_u32(sQ), k_lo = tc::smem_u32(sK), v_lo = tc::smem_u32(sV);
    uint32_t kv_it = 0, item_it = 0, g = 0;  // g: blocks processed so far (all items) == index of the next P / S buffer use
    // S buffer (gj & 1) = Q K(gj)^T, then signal the softmax warps; frees the K stage
    auto issue_qk = [&](uint32_t gj) {
      const int s = gj & 1;
      tc::mbar_wait(&k_full[s], (gj >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t d = tmem + s * 128;
      const uint32_t ka = k_lo + s * kKStage;
#pragma unroll
      for (int k = 0; k < DK_STEPS; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        const uint64_t da = tc::smem_desc_k_sw128(q_lo + off), db = tc::smem_desc_k_sw128(ka + off);
        if (leader) tc::umma_bf16(d, da, db, idesc_qk, k ? 1u : 0u);
      }
      if (leader) {
        tc::umma_commit(&s_full[s]);
        tc::umma_commit(&k_empty[s]);
      }
    };
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      tc::mbar_wait(&q_full, item_it & 1);
      issue_qk(kv_it);
      if (nb > 1) issue_qk(kv_it + 1);
      if (nb <= 2 && leader) tc::umma_commit(&q_empty);
      for (int j = 0; j < nb; ++j, ++g) {
        const uint32_t kvi = kv_it + j;
        const int s = kvi & 1;
        // O += P(j) V(j): P from TMEM (16 keys = 8 packed columns per MMA), V^T from smem
        tc::mbar_wait(&p_full[0], g & 1);
        tc::mbar_wait(&v_full[s], (kvi >> 1) & 1);
        tc::tc_fence_after();
        {
          const uint32_t dO = tmem + 256, aP = tmem + s * 128;
          const uint32_t va = v_lo + s * kVStage;
          if constexpr (kSplit) {
#pragma unroll
            for (int k = 0; k < BKEY / 32; ++k) {  // keys 0..63 -> accumulator 0 (P at S columns 0..31)
              const uint64_t db = tc::smem_desc_k_sw128(va + (k & 3) * 32);
              if (leader) tc::umma_bf16_ts(dO, aP + k * 8, db, idesc_pv, (j | k) ? 1u : 0u);
            }
            tc::mbar_wait(&p_full[1], g & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int k = 0; k < BKEY / 32; ++k) {  // keys 64..127 -> accumulator 1 (P at S columns 64..95)
              const uint64_t db = tc::smem_desc_k_sw128(va + kVChunk + (k & 3) * 32);
              if (leader) tc::umma_bf16_ts(dO + DV, aP + 64 + k * 8, db, idesc_pv, (j | k) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BKEY / 16; ++k) {
              const uint32_t boff = (k >> 2) * kVChunk + (k & 3) * 32;
              const uint64_t db = tc::smem_desc_k_sw128(va + boff);
              if (leader) tc::umma_bf16_ts(dO, aP + k * 8, db, idesc_pv, (j | k) ? 1u : 0u);
            }
          }
          if (leader) {
            tc::umma_commit(&v_empty[s]);
            tc::umma_commit(&pv_done);
          }
        }
        if (j + 2 < nb) {
          issue_qk(kvi + 2);
          if (j + 3 == nb && leader) tc::umma_commit(&q_empty);  // last QK of the item issued: Q smem free when it retires
        }
      }
      kv_it += nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int sw = warp - 2;
    const int half = sw >> 2;        // which 64 keys of every block (and which half of the O columns)
    const int q = warp & 3;          // TMEM lane quarter
    const int r = q * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tO = tmem + 256 + lane_addr;
    const uint32_t tOh = tO + (kSplit ? half * DV : 0);  // split mode: this thread's own accumulator
    const int bar_id = 1 + q;        // named barrier shared by the two warps of this quarter
    float* my_x = &xchg[0][half][r];
    const float* other_x = &xchg[0][half ^ 1][r];
    constexpr int kSlot = 2 * BQ;    // floats per exchange slot
    constexpr int DVH0 = ((DV / 2 + 15) / 16) * 16;  // O columns [0, DVH0) belong to half 0, the rest to half 1
    const int oc0 = half ? DVH0 : 0, oc1 = half ? DV : DVH0;
    uint32_t g = 0, item_it = 0;     // g: blocks processed so far (all items)
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      const int h = (int)(it % p.heads);
      const long long tq = it / p.heads;
      const int b = (int)(tq / p.tiles_per_img);
      const int qt = (int)(tq % p.tiles_per_img);
      if (kSplit && half && p.stagger_ns) __nanosleep(p.stagger_ns);
      float m_ref = -INFINITY;  // exponent reference, in log2 units (s * log2e); identical in both threads of a row
      float l = 0.f;            // this thread's part of the row sum
      for (int j = 0; j < nb; ++j, ++g) {
        const uint32_t tS = tmem + (g & 1) * 128 + lane_addr;
        tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc::tc_fence_after();
        uint32_t sv[64];  // this thread's 64 keys of the S row, read from TMEM once
        tc::tmem_ld32(tS + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tc::tmem_ld32(tS + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tc::tmem_ld_wait();
        const int kbase = j * BKEY + half * 64;
        if (kbase + 64 > p.nkeys) {  // only the last key block can hold padded keys (warp-uniform)
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (kbase + e >= p.nkeys) sv[e] = __float_as_uint(-INFINITY);
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: the row max is latency-bound otherwise
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          mx[0] = fmaxf(mx[0], __uint_as_float(sv[e]));
          mx[1] = fmaxf(mx[1], __uint_as_float(sv[e + 1]));
          mx[2] = fmaxf(mx[2], __uint_as_float(sv[e + 2]));
          mx[3] = fmaxf(mx[3], __uint_as_float(sv[e + 3]));
        }
        const float mine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        // Exchange with the thread that holds the other 64 keys of this row.  The barrier also orders
        // this thread's S loads before the partner's P stores (P aliases the first 64 columns of S).
        // Slots alternate per block: the partner reads slot (j&1) right after barrier j and cannot pass
        // barrier j+1 before that, so the write of block j+2 to the same slot is safe.
        float mj;
        if constexpr (kSplit) {
          mj = mine * kLog2e;  // this half's own reference: nothing to exchange
        } else {
          my_x[(j & 1) * kSlot] = mine;
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          mj = fmaxf(mine, other_x[(j & 1) * kSlot]) * kLog2e;
        }
        // lazy max update: only move the reference when the row max grew by more than 2^8
        float scale = 1.f;
        bool need = false;
        if (mj > m_ref + kRescaleThresh) {
          scale = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - mj);
          m_ref = mj;
          need = (j > 0);
        }
        if (__any_sync(0xffffffffu, need)) {  // rare: rescale this warp's share of O once PV(j-1) has retired
          tc::mbar_wait(&pv_done, (g - 1) & 1);
          tc::tc_fence_after();
          const float f = need ? scale : 1.f;
          for (int c = kSplit ? 0 : oc0; c < (kSplit ? DV : oc1); c += 16) {
            uint32_t o[16];
            tc::tmem_ld16(tOh + c, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tc::tmem_st16(tOh + c, o);
          }
        }
        const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;  // a half with no valid key yet: P = 2^-inf = 0
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 16 keys -> 8 packed bf16x2 columns of P
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c * 16 + e]), kLog2e, neg_m));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c * 16 + e + 1]), kLog2e, neg_m));
            rs[(e >> 1) & 3] += p0 + p1;
            __nv_bfloat162 bb = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          tc::tmem_st8(tS + half * (kSplit ? 64 : 32) + c * 8, pk);  // split mode: inside this thread's own S columns
        }
        l = l * scale + (rs[0] + rs[1]) + (rs[2] + rs[3]);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          // S(j+1) is ready a block ahead, so a warp may finish block j+1 before a sibling has arrived for block j: never
          // arrive twice in one phase (the previous phase has normally completed long ago -- one poll)
          if (g > 0) tc::mbar_wait(&p_full[kSplit ? half : 0], (g - 1) & 1);
          tc::mbar_arrive(&p_full[kSplit ? half : 0]);
        }
      }
      // epilogue: O / l -> bf16 -> dense smem tile -> one TMA store (clipped at the image's last row);
      // the wide-head variant has no smem left for the tile and stores its rows directly
      my_x[2 * kSlot] = l;
      if constexpr (kSplit) my_x[0] = m_ref;
      if (PL::kStageO && sw == 0 && lane == 0) tc::tma_store_wait_read<0>();  // previous tile's store has read the staging tile
      asm volatile("bar.sync 9, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
      float inv, f0 = 1.f, f1 = 0.f;  // O = (O_0 f0 + O_1 f1) * inv
      float m_row = m_ref, l_row;      // row maximum (log2 units) and the row sum relative to it
      if constexpr (kSplit) {
        const float m_o = other_x[0], l_o = other_x[2 * kSlot];
        const float m = fmaxf(m_ref, m_o);
        const float f_me = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m);
        const float f_o = (m_o == -INFINITY) ? 0.f : ex2_approx(m_o - m);
        m_row = m;
        l_row = l * f_me + l_o * f_o;
        f0 = half ? f_o : f_me;
        f1 = half ? f_me : f_o;
      } else {
        l_row = l + other_x[2 * kSlot];
      }
      inv = 1.f / l_row;
      if (p.lse && half == 0 && (long long)qt * BQ + r < p.rows_per_img)
        p.lse[((long long)b * p.heads + h) * p.rows_per_img + (long long)qt * BQ + r] = m_row + log2f(l_row);
      tc::mbar_wait(&pv_done, (g - 1) & 1);
      tc::tc_fence_after();
      const long long row_local = (long long)qt * BQ + r;
      uint8_t* orow = PL::kStageO ? sO + r * (DV * 2)
                                  : reinterpret_cast<uint8_t*>(p.out) +
                                        (((long long)b * p.rows_per_img + row_local) * p.ldo + h * p.o_head_stride) * 2;
      const bool row_ok = PL::kStageO || row_local < p.rows_per_img;
      for (int c = oc0; c < oc1; c += 16) {
        uint32_t o[16];
        tc::tmem_ld16(tO + c, o);
        if constexpr (kSplit) {
          uint32_t o1[16];
          tc::tmem_ld16(tO + DV + c, o1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            o[e] = __float_as_uint(fmaf(__uint_as_float(o[e]), f0, __uint_as_float(o1[e]) * f1));
        } else {
          tc::tmem_ld_wait();
        }
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(o[2 * e]) * inv, __uint_as_float(o[2 * e + 1]) * inv);
          w[e] = *reinterpret_cast<uint32_t*>(&bb);
        }
        if (row_ok) {
          *reinterpret_cast<uint4*>(orow + c * 2) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(orow + c * 2 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      // the next item's PV(0) overwrites O only after all eight warps' next p_full arrivals, which
      // follow these TMEM reads in program order
      tc::tc_fence_before();
      if (PL::kStageO) {
        tc::fence_proxy_async();
        asm volatile("bar.sync 9, %0;" ::"n"(32 * kSoftmaxWarps) : "memory");
        if (sw == 0 && lane == 0) {
          tc::tma_store_3d(&tmO, sO, h * p.o_head_stride, qt * BQ, b);
          tc::tma_store_commit();
        }
      }
    }
    if (sw == 0 && lane == 0) tc::tma_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace attn
}  // namespace isp

using namespace isp;

namespace isp {
int attention_pair_launch(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt, void* out,
                          long long ldo, int o_head_stride, int B, long long rows_per_img, int heads, int nkeys, float* lse,
                          int lsum_col, int poly, isp_stream_t stream);  // attention_pair_tc.cu
}

// Q: bf16 [B*rows_per_img, ldq]; head h reads columns [h*q_head_stride, +DK) (zero K padding
// makes any extra columns harmless).  K: bf16 [B, heads, nblocks*128, DKC] (DKC = 64 or 128,
// zero padded).  Vt: bf16 [B, heads, DV, nblocks*128].  out: bf16 [B*rows_per_img, ldo], head h
// writes DV columns at h*o_head_stride.  variant 0: head_dim <= 64 (DV = 64); 1: <= 112 (DV = 112) on the two-tile kernel of
// attention_pair_tc.cu; 2: <= 144 (DV = 144, K padded to 192 columns); 3: the DV = 112 geometry on this file's one-tile
// kernel (cross-check in the tests, and the "before" of tools/tune_attention.py).
static int attention_fwd(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt, void* out,
                         long long ldo, int o_head_stride, int B, long long rows_per_img, int heads, int nkeys, int variant,
                         float* lse, int lsum_col, int poly, isp_stream_t stream) {
  ISP_REQUIRE(Q && K && Vt && out, ISP_ERR_BAD_SHAPE, "attention_bf16_tc: null pointer");
  ISP_REQUIRE(B > 0 && rows_per_img > 0 && heads > 0 && nkeys > 0, ISP_ERR_BAD_SHAPE, "attention_bf16_tc: bad shape");
  ISP_REQUIRE(variant >= 0 && variant <= 3, ISP_ERR_UNSUPPORTED, "attention_bf16_tc: variant %d", variant);
  const bool pair = variant == 1;
  if (variant == 3) variant = 1;
  const int DV = variant == 0 ? 64 : variant == 1 ? 112 : 144, DKC = variant == 0 ? 64 : variant == 1 ? 128 : 192;
  ISP_REQUIRE(lsum_col < DV, ISP_ERR_BAD_SHAPE, "attention_bf16_tc: lsum_col %d outside the %d value columns", lsum_col, DV);
  ISP_REQUIRE((lsum_col < 0 && poly == 0) || pair, ISP_ERR_UNSUPPORTED,
              "attention_bf16_tc: the ones-column row sum / polynomial exp2 exist for variant 1 only");
  ISP_REQUIRE(poly == 0 || (lsum_col >= 0 && (poly == 2 || poly == 3 || poly == 4)), ISP_ERR_UNSUPPORTED,
              "attention_bf16_tc: poly must be 0, or 2 / 3 / 4 (of every 8 exponentials) together with lsum_col >= 0");
  ISP_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && o_head_stride % 8 == 0 && q_head_stride % 8 == 0, ISP_ERR_MISALIGNED,
              "attention_bf16_tc: ldq/ldo/q_head_stride/o_head_stride must be multiples of 8 (TMA box starts and "
              "vector stores need 16-byte alignment)");
  ISP_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(Vt) && aligned16(out), ISP_ERR_MISALIGNED,
              "attention_bf16_tc: 16-byte alignment");
  ISP_REQUIRE((long long)B * rows_per_img < (1ll << 31), ISP_ERR_UNSUPPORTED, "attention_bf16_tc: too many rows");
  if (pair)
    return attention_pair_launch(Q, ldq, q_head_stride, K, Vt, out, ldo, o_head_stride, B, rows_per_img, heads, nkeys, lse,
                                 lsum_col, poly, stream);
  attn::Params p = {};
  p.nkeys = nkeys;
  p.nblocks = (nkeys + attn::BKEY - 1) / attn::BKEY;
  p.heads = heads;
  p.rows_per_img = rows_per_img;
  p.tiles_per_img = (int)((rows_per_img + attn::BQ - 1) / attn::BQ);
  p.nitems = (long long)B * p.tiles_per_img * heads;
  p.q_head_stride = q_head_stride;
  p.o_head_stride = o_head_stride;
  p.out = out; p.ldo = ldo;
  p.lse = lse;
  p.stagger_ns = 400;  // measured 150..800 ns: 1.53 -> 1.48 ms on the LoftUp shape (the two warps of an SM sub-partition
                       // then alternate between their TMEM-load and exponential phases instead of colliding)
  const long long nkp = (long long)p.nblocks * attn::BKEY;
  CUtensorMap tmQ, tmK, tmV, tmO;
  {  // out viewed as [B][rows_per_img][ldo]: a tile's store is clipped at its own image's last row
    const uint64_t dims[3] = {(uint64_t)ldo, (uint64_t)rows_per_img, (uint64_t)B};
    const uint64_t str[3] = {2, (uint64_t)ldo * 2, (uint64_t)rows_per_img * ldo * 2};
    const uint32_t box[3] = {(uint32_t)DV, attn::BQ, 1};
    if (int e = make_tmap(&tmO, 2, out, 3, dims, str, box, "attention(out)", false)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)ldq, (uint64_t)B * rows_per_img}, str[2] = {2, (uint64_t)ldq * 2};
    const uint32_t box[2] = {64, attn::BQ};
    if (int e = make_tmap_bf16(&tmQ, Q, 2, dims, str, box, "attention(Q)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)DKC, (uint64_t)B * heads * nkp}, str[2] = {2, (uint64_t)DKC * 2};
    const uint32_t box[2] = {64, attn::BKEY};
    if (int e = make_tmap_bf16(&tmK, K, 2, dims, str, box, "attention(K)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)nkp, (uint64_t)B * heads * DV}, str[2] = {2, (uint64_t)nkp * 2};
    const uint32_t box[2] = {64, (uint32_t)DV};
    if (int e = make_tmap_bf16(&tmV, Vt, 2, dims, str, box, "attention(Vt)")) return e;
  }
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  auto smem_of = [](uint32_t plan) { return (int)(plan > attn::kSmemMin ? plan : attn::kSmemMin); };
  const int sm0 = smem_of(attn::Plan<1, 64>::kBytes), sm1 = smem_of(attn::Plan<2, 112>::kBytes),
            sm2 = smem_of(attn::Plan<3, 144>::kBytes);
  if (int e = ensure_dynamic_smem((const void*)attn::attention_kernel<1, 4, 64>, sm0)) return e;
  if (int e = ensure_dynamic_smem((const void*)attn::attention_kernel<2, 7, 112>, sm1)) return e;
  if (int e = ensure_dynamic_smem((const void*)attn::attention_kernel<3, 9, 144>, sm2)) return e;
  const int grid = (int)(p.nitems < num_sms ? p.nitems : num_sms);
  if (variant == 2)
    attn::attention_kernel<3, 9, 144><<<grid, attn::kThreads, sm2, as_stream(stream)>>>(tmQ, tmK, tmV, tmO, p);
  else if (variant == 1)
    attn::attention_kernel<2, 7, 112><<<grid, attn::kThreads, sm1, as_stream(stream)>>>(tmQ, tmK, tmV, tmO, p);
  else
    attn::attention_kernel<1, 4, 64><<<grid, attn::kThreads, sm0, as_stream(stream)>>>(tmQ, tmK, tmV, tmO, p);
  ISP_CHECK_LAUNCH("attention_kernel");
  return ISP_OK;
}

extern "C" int isp_attention_bf16_tc(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                                     void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                                     int heads, int nkeys, int variant, isp_stream_t stream) {
  return attention_fwd(Q, ldq, q_head_stride, K, Vt, out, ldo, o_head_stride, B, rows_per_img, heads, nkeys, variant,
                       nullptr, -1, 0, stream);
}

// Same, and also writes lse[B][heads][rows_per_img] (fp32, log2 units: log2 of the sum over keys of 2^(s*log2e)), the
// row statistic isp_attention_bwd_bf16_tc needs to recompute the probabilities tile by tile.
extern "C" int isp_attention_bf16_tc_lse(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                                         void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                                         int heads, int nkeys, int variant, float* lse, isp_stream_t stream) {
  ISP_REQUIRE(lse, ISP_ERR_BAD_SHAPE, "attention_bf16_tc_lse: null lse");
  return attention_fwd(Q, ldq, q_head_stride, K, Vt, out, ldo, o_head_stride, B, rows_per_img, heads, nkeys, variant, lse,
                       -1, 0, stream);
}

// Tuned entry of the two-tile kernel (variant 1, LoftUp cross-attention): `lsum_col` >= 0 names a V^T row the caller filled
// with ones (a zero-padding row of the head, e.g. 101 for head_dim 101 in 112) -- the row sum of the probabilities then
// accumulates in that O column on the tensor pipe instead of 128 FADDs per thread and key block; `poly` in {0, 2, 3, 4}
// moves that many of every 8 exponentials from the MUFU pipe to an FMA-pipe polynomial (max relative error 1e-4, below the
// bf16 rounding of P).  lse may be NULL.  Column lsum_col of `out` receives 1.0 (callers' next weight matrix has zero rows
// there).
extern "C" int isp_attention_bf16_tc_opt(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                                         void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                                         int heads, int nkeys, int variant, float* lse, int lsum_col, int poly,
                                         isp_stream_t stream) {
  return attention_fwd(Q, ldq, q_head_stride, K, Vt, out, ldo, o_head_stride, B, rows_per_img, heads, nkeys, variant, lse,
                       lsum_col, poly, stream);
}
