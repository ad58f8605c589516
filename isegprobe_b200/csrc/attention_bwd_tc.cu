// Flash-style attention BACKWARD on tcgen05 (sm_100a): dK, dV (and optionally dQ) of O = softmax(Q K^T) V without
// writing any score-shaped tensor to HBM.  Replaces the materialised path (S, P, dP, dS = 12 bytes per score through
// HBM; LoftUp: 822 M scores per image and layer) under the LoftUp cross-attention backward
// (nn.MultiheadAttention autograd in the reference, loftup/layers.py:186-202 under trainer.py:213-221).
//
// One CTA owns (image, head, block of 128 keys, chunk of query rows) and walks its queries 64 at a time.  Everything is
// computed TRANSPOSED so that the 128 keys are the MMA M dimension (full-rate 128-row MMAs) and the TMEM lanes:
//
//   S^T  [128 keys x 64 q] = K  Q^T          A = K tile (K-major),  B = Q tile  (K-major)      -> TMEM (double-buffered)
//   dP^T [128 keys x 64 q] = V dO^T          A = V tile (K-major),  B = dO tile (K-major)      -> TMEM (double-buffered)
//   P^T = 2^(S^T log2e - lse[q]),  dS^T = P^T (dP^T - D[q])      softmax warps: TMEM -> registers -> bf16 -> back into TMEM
//                                                                 over the thread's own S^T columns (keys on the lanes is
//                                                                 exactly the A-operand layout of the next two products)
//   dV [128 keys x d] += P^T  dO             A = P^T  from TMEM,   B = dO tile walked MN-major (the same smem tile)
//   dK [128 keys x d] += dS^T Q              A = dS^T from TMEM,   B = Q tile walked MN-major (the same smem tile)
//   dQ^T [d x 64 q]    = K^T dS^T  (optional) A = K tile walked MN-major, B = dS^T as a 128B-swizzled smem tile written by the
//                                            softmax warps; accumulator in its own 64 TMEM columns
//
// The Q / dO tiles are therefore loaded ONCE per step and used as two different operands; dK / dV stay in TMEM for the
// whole chunk and are added to the fp32 gradients with red.global.add.v4 at the end (chunks of the same key block
// accumulate there); dQ^T is drained every step with red.global.add (the eight key blocks of a query add up).
// lse is the forward's log-sum-exp in log2 units (isp_attention_bf16_tc_lse), D[q] = sum_d dO[q,d] O[q,d]
// (isp_attention_rowdot_heads); both are [B][heads][rows].
//
// Warps: 0 = TMA producer (4-deep Q/dO ring; lse/D by 1-D bulk copies on the same barrier, or copied by the warp when the
// row count is not a multiple of 4), 1 = MMA issuer (S^T / dP^T two steps ahead of the softmax; descriptors built once and
// advanced by constants -- the issuing warp's instruction stream was the first bottleneck), 2..17 = softmax (TMEM lane =
// key; the four warps of a lane quarter take 16 queries each).  TMEM column plan: see col_s / col_dp below; dV at 256, dK
// at 384.  tcgen05.mma executes in issue order, which orders the reads of P^T(i) / dS^T(i) before S^T(i+2) overwrites
// their buffer.
#include "tc_common.cuh"

namespace isp {
namespace attnbwd {

constexpr int BK = 128;            // keys per CTA
constexpr int BQ = 64;             // queries per step
constexpr int kStages = 4;
constexpr int kSoftmaxWarps = 16;          // four per TMEM lane quarter: each takes 16 of the step's 64 queries
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr uint32_t kKVBytes = 2 * 16384;          // K (or V) tile: two [128 keys x 64 cols] boxes
constexpr uint32_t kQBytes = 2 * 8192;            // Q (or dO) tile: two [64 q x 64 cols] boxes
constexpr uint32_t kStageBytes = 2 * kQBytes;
constexpr uint32_t kPBytes = 16384;               // dS^T smem tile [128 keys x 64 q] (B operand of the dQ^T product)
constexpr uint32_t kSmem = 2 * kKVBytes + kStages * kStageBytes + kPBytes;  // 208 KB
constexpr float kLog2e = 1.4426950408889634f;
// TMEM columns.  Once a step's S^T is in registers its columns are reused: each softmax thread writes P^T (8 packed bf16
// columns) and dS^T (8) over ITS OWN 16 S^T columns.
constexpr uint32_t cDV = 256, cDK = 384;
// S^T(i) / dP^T(i) columns.  Without dQ: buffer i&1 at (i&1)*128 holds both (S^T @+0, dP^T @+64).  With dQ the 64 columns of
// dQ^T need a home of their own (a dQ^T that lived in the dP^T buffer had to be drained before that buffer's next S^T / dP^T
// could be issued, which put the tensor pipe and the softmax warps in series): S^T stays double-buffered (@0, @64), dP^T is
// single-buffered @128 (re-issued as soon as the softmax warps hold it in registers), dQ^T @192.
template <bool DQ> __device__ __forceinline__ uint32_t col_s(int i) { return DQ ? (i & 1) * 64 : (i & 1) * 128; }
template <bool DQ> __device__ __forceinline__ uint32_t col_dp(int i) { return DQ ? 128 : (i & 1) * 128 + 64; }
constexpr uint32_t cDQ = 192;

struct Params {
  int nkeys, heads, nkb;         // real keys per image, heads, key blocks per image
  long long rows;                // queries per image
  int qchunks;                   // query chunks per (image, head, key block)
  int steps_per_chunk;           // 64-query steps per chunk (the last chunk may have fewer)
  int steps_total;               // ceil(rows / 64)
  int HP;                        // head pitch in Q / dO / dK / dV columns (multiple of 16, <= 128)
  const float* lse;              // [B][heads][rows]
  const float* dvec;             // [B][heads][rows]
  float* dK;                     // [B][heads][nkeys][HP] fp32, accumulated into
  float* dV;
  float* dQ;                     // optional fp32 [B*rows][lddq], head h at column h*HP; accumulated into
  long long lddq;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t kAmn = 1u << 15, kBmn = 1u << 16;  // operand is MN-major

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// mbarrier wait with a wall-clock bound (10 s): a protocol bug reports which barrier starved and traps instead of
// spinning through the generic 2^24-probe limit
__device__ __noinline__ void wait_timeout(int tag, uint32_t parity) {
  printf("attention_bwd_kernel: barrier %d (parity %u) starved, block %d warp %d\n", tag, parity, (int)blockIdx.x,
         (int)(threadIdx.x >> 5));
  __trap();
}
__device__ __forceinline__ void bwait(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (!t0) t0 = t;
      else if (t - t0 > 10000000000ull) wait_timeout(tag, parity);
    }
  }
}

template <bool DQ>
__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t kv_full, full_bar[kStages], empty_bar[kStages], s_full[2], p_full, dq_full, dq_free, dp_full, dp_free, acc_full;
  __shared__ __align__(16) float stat[kStages][2][BQ];  // [stage][lse | D][query]
  __shared__ uint32_t tmem_base_s;

  uint8_t* sK = smem;
  uint8_t* sV = sK + kKVBytes;
  uint8_t* sQ0 = sV + kKVBytes;                   // stage s: Q at sQ0 + s*kStageBytes, dO right after it
  uint8_t* sDS = sQ0 + kStages * kStageBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item
  long long it = blockIdx.x;
  const int qc = (int)(it % p.qchunks); it /= p.qchunks;
  const int kb = (int)(it % p.nkb); it /= p.nkb;
  const int h = (int)(it % p.heads);
  const int b = (int)(it / p.heads);
  const int step0 = qc * p.steps_per_chunk;
  const int nsteps = min(p.steps_per_chunk, p.steps_total - step0);
  const long long stat_base = ((long long)b * p.heads + h) * p.rows;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmDO); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
    tc::mbar_init(&kv_full, 1);
    for (int i = 0; i < kStages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(&s_full[0], 1); tc::mbar_init(&s_full[1], 1);
    tc::mbar_init(&p_full, kSoftmaxWarps);
    tc::mbar_init(&dq_full, 1); tc::mbar_init(&dq_free, kSoftmaxWarps);
    tc::mbar_init(&dp_full, 1); tc::mbar_init(&dp_free, kSoftmaxWarps);
    tc::mbar_init(&acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (nsteps <= 0) {  // (cannot happen with the host's chunking; keeps the barriers consistent if it ever does)
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
    return;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int krow = (b * p.heads + h) * p.nkeys + kb * BK;  // row of the [B*heads*nkeys, HP] key / value matrices
      tc::mbar_arrive_expect_tx(&kv_full, 2 * kKVBytes);
      for (int c = 0; c < 2; ++c) {
        tc::tma_load_2d(sK + c * 16384, &tmK, &kv_full, c * 64, krow);
        tc::tma_load_2d(sV + c * 16384, &tmV, &kv_full, c * 64, krow);
      }
    }
    const bool stat_bulk = (p.rows & 3) == 0;  // 16-byte aligned row statistics: 1-D bulk copies; else the warp copies them
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % kStages;
      bwait(&empty_bar[s], ((i / kStages) & 1) ^ 1, 0);
      const long long q0 = (long long)(step0 + i) * BQ;
      if (!stat_bulk) {
#pragma unroll
        for (int j = lane; j < BQ; j += 32) {
          const bool ok = q0 + j < p.rows;
          stat[s][0][j] = ok ? __ldg(p.lse + stat_base + q0 + j) : 0.f;
          stat[s][1][j] = ok ? __ldg(p.dvec + stat_base + q0 + j) : 0.f;
        }
        __syncwarp();  // orders the other lanes' stores before lane 0's (releasing) arrive below
      }
      if (lane == 0) {
        const int qrow = (int)((long long)b * p.rows + q0);
        uint8_t* sq = sQ0 + s * kStageBytes;
        tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes + (stat_bulk ? 2 * BQ * 4 : 0));
        for (int c = 0; c < 2; ++c) {
          tc::tma_load_2d(sq + c * 8192, &tmQ, &full_bar[s], h * p.HP + c * 64, qrow);
          tc::tma_load_2d(sq + kQBytes + c * 8192, &tmDO, &full_bar[s], h * p.HP + c * 64, qrow);
        }
        if (stat_bulk) {
          bulk_load(&stat[s][0][0], p.lse + stat_base + q0, BQ * 4, &full_bar[s]);
          bulk_load(&stat[s][1][0], p.dvec + stat_base + q0, BQ * 4, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = tc::elect_one();
    const int ksteps = p.HP / 16;
    const uint32_t id_s = tc::idesc_bf16_f32(BK, BQ);                 // S^T, dP^T: both operands K-major
    const uint32_t id_g = tc::idesc_bf16_f32(BK, p.HP) | kBmn;        // dV, dK: A K-major (P^T / dS^T), B MN-major
    const uint32_t id_q = tc::idesc_bf16_f32(128, BQ) | kAmn | kBmn;  // dQ^T: K tile and dS^T tile both MN-major
    const uint32_t aK = tc::smem_u32(sK), aV = tc::smem_u32(sV), aDS = tc::smem_u32(sDS);
    // Descriptors are built once; every MMA only adds a compile-time offset to the 14-bit start-address field (>> 4): the
    // single issuing warp's instruction stream is ~6 instructions per MMA (it was the bottleneck when each descriptor was
    // rebuilt from its address: 22-30 MMAs per step).
    const uint64_t dKk = tc::smem_desc_k_sw128(aK), dVk = tc::smem_desc_k_sw128(aV);
    const uint64_t dKmn = desc_mn(aK, 16384), dDSmn = desc_mn(aDS, 8192);
    auto issue_s = [&](int i, bool with_dp) {  // S^T(i) (and dP^T(i)) -> TMEM
      const int s = i % kStages;
      bwait(&full_bar[s], (i / kStages) & 1, 1);
      tc::tc_fence_after();
      const uint32_t aQ = tc::smem_u32(sQ0 + s * kStageBytes);
      const uint64_t dQk = tc::smem_desc_k_sw128(aQ), dDOk = tc::smem_desc_k_sw128(aQ + kQBytes);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k)  // K step k: 32 bytes further inside the 128-byte rows, second box after four steps
          if (k < ksteps)
            tc::umma_bf16(tmem + col_s<DQ>(i), dKk + (((k >> 2) * 16384 + (k & 3) * 32) >> 4),
                          dQk + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), id_s, k ? 1u : 0u);
        if (with_dp) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < ksteps)
              tc::umma_bf16(tmem + col_dp<DQ>(i), dVk + (((k >> 2) * 16384 + (k & 3) * 32) >> 4),
                            dDOk + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), id_s, k ? 1u : 0u);
        }
        tc::umma_commit(&s_full[i & 1]);
      }
    };
    auto issue_dp = [&](int i) {  // dQ variant: dP^T(i) into the single dP^T buffer
      const int s = i % kStages;
      bwait(&full_bar[s], (i / kStages) & 1, 1);
      tc::tc_fence_after();
      const uint64_t dDOk = tc::smem_desc_k_sw128(tc::smem_u32(sQ0 + s * kStageBytes) + kQBytes);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < ksteps)
            tc::umma_bf16(tmem + col_dp<DQ>(i), dVk + (((k >> 2) * 16384 + (k & 3) * 32) >> 4),
                          dDOk + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), id_s, k ? 1u : 0u);
        tc::umma_commit(&dp_full);
      }
    };
    bwait(&kv_full, 0, 2);
    issue_s(0, !DQ);
    if (nsteps > 1) issue_s(1, !DQ);
    if constexpr (DQ) issue_dp(0);
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % kStages;
      if constexpr (DQ) {
        if (i + 1 < nsteps) {
          bwait(&dp_free, i & 1, 3);  // the softmax warps hold dP^T(i) in registers
          tc::tc_fence_after();
          issue_dp(i + 1);
        }
      }
      bwait(&p_full, i & 1, 4);    // P^T(i), dS^T(i) are in TMEM (over S^T(i)'s columns)
      tc::tc_fence_after();
      const uint32_t tb = tmem + col_s<DQ>(i);
      const uint32_t aQ = tc::smem_u32(sQ0 + s * kStageBytes);
      const uint64_t dQmn = desc_mn(aQ, 8192), dDOmn = desc_mn(aQ + kQBytes, 8192);
      if constexpr (DQ) {
        if (i > 0) {
          bwait(&dq_free, (i - 1) & 1, 5);  // dQ^T(i-1) has been drained (the softmax warps do that before p_full(i))
          tc::tc_fence_after();
        }
      }
      if (leader) {
        // A operands straight from TMEM: queries [16k, 16k+16) -> P^T at columns 16k (8 packed columns), dS^T at 16k + 8;
        // B: 16 query rows = 2048 bytes further per K step
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc::umma_bf16_ts(tmem + cDV, tb + k * 16, dDOmn + k * 128, id_g, (i | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) tc::umma_bf16_ts(tmem + cDK, tb + k * 16 + 8, dQmn + k * 128, id_g, (i | k) ? 1u : 0u);
        if constexpr (DQ) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) tc::umma_bf16(tmem + cDQ, dKmn + k * 128, dDSmn + k * 128, id_q, k ? 1u : 0u);
          tc::umma_commit(&dq_full);
        }
        tc::umma_commit(&empty_bar[s]);
      }
      if (i + 2 < nsteps) issue_s(i + 2, !DQ);  // the buffer's previous contents are read by MMAs issued above
    }
    if (leader) tc::umma_commit(&acc_full);
  } else {
    // ------------------------------------------------------------------ softmax warps
    // 16 warps = 4 per SM sub-partition (the arithmetic is latency-bound with fewer: ex2 / LDS / TMEM-load latencies)
    const int sw = warp - 2;
    const int cq = sw >> 2;              // which 16 queries of the step (== the K step of the dV / dK MMAs that reads them)
    const int qd = warp & 3;             // TMEM lane quarter
    const int key = qd * 32 + lane;      // key inside the block == TMEM lane
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const bool key_ok = kb * BK + key < p.nkeys;
    const bool keys_partial = kb * BK + BK > p.nkeys;  // warp-uniform: this key block has padded keys
    uint8_t* dsrow = sDS + key * 128;
    const int sx = key & 7;

    auto drain_dq = [&](int i) {  // dQ^T(i): lane = head-dim index, columns = this warp's 16 queries
      bwait(&dq_full, i & 1, 6);
      tc::tc_fence_after();
      uint32_t v[16];
      tc::tmem_ld16(tmem + lane_addr + cDQ + cq * 16, v);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&dq_free);
      if (key < p.HP) {
        const long long q0 = (long long)(step0 + i) * BQ + cq * 16;
        float* dst = p.dQ + ((long long)b * p.rows + q0) * p.lddq + h * p.HP + key;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (q0 + e < p.rows) red_add(dst + (long long)e * p.lddq, __uint_as_float(v[e]));
      }
    };

    for (int i = 0; i < nsteps; ++i) {
      const int s = i % kStages;
      const uint32_t tb = tmem + lane_addr + col_s<DQ>(i) + cq * 16;
      bwait(&s_full[i & 1], (i >> 1) & 1, 7);
      if constexpr (DQ) bwait(&dp_full, i & 1, 12);
      tc::tc_fence_after();
      uint32_t sv[16], dp[16];
      tc::tmem_ld16(tb, sv);
      tc::tmem_ld16(tmem + lane_addr + col_dp<DQ>(i) + cq * 16, dp);
      tc::tmem_ld_wait();
      if constexpr (DQ) {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&dp_free);
      }
      bwait(&full_bar[s], (i / kStages) & 1, 11);  // long complete (the MMAs read the stage): acquires this stage's lse / D
      const float4* ls = reinterpret_cast<const float4*>(&stat[s][0][cq * 16]);
      const float4* dv = reinterpret_cast<const float4*>(&stat[s][1][cq * 16]);
      const long long q0 = (long long)(step0 + i) * BQ + cq * 16;
      const int nvalid = (int)min((long long)16, p.rows - q0);  // may be <= 0 in the last step
      uint32_t pk[8], dk[8];
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 l4 = ls[e >> 2], d4 = dv[e >> 2];
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), kLog2e, -l4.x));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), kLog2e, -l4.y));
        const float p2 = ex2_approx(fmaf(__uint_as_float(sv[e + 2]), kLog2e, -l4.z));
        const float p3 = ex2_approx(fmaf(__uint_as_float(sv[e + 3]), kLog2e, -l4.w));
        const float g0 = p0 * (__uint_as_float(dp[e]) - d4.x), g1 = p1 * (__uint_as_float(dp[e + 1]) - d4.y);
        const float g2 = p2 * (__uint_as_float(dp[e + 2]) - d4.z), g3 = p3 * (__uint_as_float(dp[e + 3]) - d4.w);
        __nv_bfloat162 a0 = __floats2bfloat162_rn(p0, p1), a1 = __floats2bfloat162_rn(p2, p3);
        __nv_bfloat162 c0 = __floats2bfloat162_rn(g0, g1), c1 = __floats2bfloat162_rn(g2, g3);
        pk[e >> 1] = *reinterpret_cast<uint32_t*>(&a0);
        pk[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&a1);
        dk[e >> 1] = *reinterpret_cast<uint32_t*>(&c0);
        dk[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&c1);
      }
      if (keys_partial || nvalid < 16) {  // warp-uniform, only the last key block / last step: zero what is padding
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t keep = (key_ok && 2 * j < nvalid ? 0x0000FFFFu : 0u) | (key_ok && 2 * j + 1 < nvalid ? 0xFFFF0000u : 0u);
          pk[j] &= keep;
          dk[j] &= keep;
        }
      }
      if constexpr (DQ) {
        // dQ^T(i-1): its MMAs were issued after p_full(i-1), a whole load + arithmetic phase ago.  Waiting for them here also
        // guarantees that the dS^T smem tile of step i-1 has been read before this step overwrites it.
        if (i > 0) drain_dq(i - 1);
      }
      // P^T / dS^T go back into TMEM over this thread's own 16 S^T columns (keys on the lanes = A-operand layout):
      // P^T in the first 8 (16 packed bf16), dS^T in the last 8
      tc::tmem_st8(tb, pk);
      tc::tmem_st8(tb + 8, dk);
      if constexpr (DQ) {  // dQ^T = K^T dS^T takes dS^T as its B operand: also as a 128B-swizzled smem tile [128 keys][64 q]
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int off = ((cq * 2 + c) ^ sx) * 16;
          *reinterpret_cast<uint4*>(dsrow + off) = make_uint4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
        }
        tc::fence_proxy_async();
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full);
    }
    if constexpr (DQ) drain_dq(nsteps - 1);
    // epilogue: dV (warps with cq = 0, 1) and dK (cq = 2, 3) of this chunk -> red.add into the fp32 gradients; the two warps
    // of a matrix split its columns at 64
    bwait(&acc_full, 0, 9);
    tc::tc_fence_after();
    {  // (tcgen05.ld is warp-collective: every lane loads, only valid keys add)
      const bool is_dk = cq >= 2;
      float* dst = (is_dk ? p.dK : p.dV) + (((long long)b * p.heads + h) * p.nkeys + kb * BK + key) * p.HP;
      const uint32_t tsrc = tmem + lane_addr + (is_dk ? cDK : cDV);
      const int c_lo = (cq & 1) ? min(64, p.HP) : 0, c_hi = (cq & 1) ? p.HP : min(64, p.HP);
      for (int c = c_lo; c < c_hi; c += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tsrc + c, v);
        tc::tmem_ld_wait();
        if (key_ok) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            red_add_v4(dst + c + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                       __uint_as_float(v[e + 3]));
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// D[b][h][row] = sum_d dO[row, h*HP + d] * O[row, h*HP + d]: one thread per (row, head), 16-byte loads.
__global__ void rowdot_heads_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ o,
                                    long long ldo, float* __restrict__ out, int B, long long rows, int heads, int HP) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * rows * heads) return;
  const int h = (int)(t % heads);
  const long long r = t / heads;  // global row
  const uint4* pa = reinterpret_cast<const uint4*>(a + r * lda + h * HP);
  const uint4* po = reinterpret_cast<const uint4*>(o + r * ldo + h * HP);
  float acc = 0.f;
  for (int c = 0; c < HP / 8; ++c) {
    const uint4 x = __ldg(pa + c), y = __ldg(po + c);
    const __nv_bfloat162* xa = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* ya = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 u = __bfloat1622float2(xa[e]), v = __bfloat1622float2(ya[e]);
      acc = fmaf(u.x, v.x, fmaf(u.y, v.y, acc));
    }
  }
  const long long bi = r / rows, rl = r % rows;
  out[(bi * heads + h) * rows + rl] = acc;
}

}  // namespace attnbwd
}  // namespace isp

using namespace isp;

// D[B][heads][rows] = per-(row, head) dot product of dO and O (both bf16 [B*rows, ld], head h at column h*HP).
extern "C" int isp_attention_rowdot_heads(const void* dO, long long lddo, const void* O, long long ldo, float* out, int B,
                                          long long rows, int heads, int HP, isp_stream_t stream) {
  ISP_REQUIRE(dO && O && out && B > 0 && rows > 0 && heads > 0, ISP_ERR_BAD_SHAPE, "attention_rowdot_heads: bad shape");
  ISP_REQUIRE(HP % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0 && aligned16(dO) && aligned16(O), ISP_ERR_MISALIGNED,
              "attention_rowdot_heads: HP / strides must be multiples of 8, pointers 16-byte aligned");
  const long long n = (long long)B * rows * heads;
  attnbwd::rowdot_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)dO, lddo, (const __nv_bfloat16*)O, ldo, out, B, rows, heads, HP);
  ISP_CHECK_LAUNCH("rowdot_heads_kernel");
  return ISP_OK;
}

// Q, dO: bf16 [B*rows, ld] (head h at columns [h*HP, +HP); Q pre-scaled as in the forward).  K, V: bf16
// [B, heads, nkeys, HP] (rows = keys).  lse, dvec: fp32 [B][heads][rows] followed by >= 64 floats of slack (the last
// step's 256-byte bulk copy may read past a partial tile).  dK, dV: fp32 [B, heads, nkeys, HP], ACCUMULATED into (zero them
// first).  dQ: optional fp32 [B*rows, lddq], accumulated into (gradient w.r.t. the pre-scaled Q).
extern "C" int isp_attention_bwd_bf16_tc(const void* Q, long long ldq, const void* dO, long long lddo, const void* K,
                                         const void* V, const float* lse, const float* dvec, float* dK, float* dV,
                                         float* dQ, long long lddq, int B, long long rows, int heads, int nkeys, int HP,
                                         isp_stream_t stream) {
  ISP_REQUIRE(Q && dO && K && V && lse && dvec && dK && dV, ISP_ERR_BAD_SHAPE, "attention_bwd_bf16_tc: null pointer");
  ISP_REQUIRE(B > 0 && rows > 0 && heads > 0 && nkeys > 0, ISP_ERR_BAD_SHAPE, "attention_bwd_bf16_tc: bad shape");
  ISP_REQUIRE(HP % 16 == 0 && HP >= 16 && HP <= 128, ISP_ERR_UNSUPPORTED, "attention_bwd_bf16_tc: HP %d (multiple of 16, <= 128)", HP);
  ISP_REQUIRE(ldq % 8 == 0 && lddo % 8 == 0 && (!dQ || lddq >= (long long)heads * HP), ISP_ERR_MISALIGNED,
              "attention_bwd_bf16_tc: ldq / lddo must be multiples of 8");
  ISP_REQUIRE(aligned16(Q) && aligned16(dO) && aligned16(K) && aligned16(V) && aligned16(lse) && aligned16(dvec) &&
                  aligned16(dK) && aligned16(dV),
              ISP_ERR_MISALIGNED, "attention_bwd_bf16_tc: 16-byte alignment");
  ISP_REQUIRE((long long)B * rows < (1ll << 31) && (long long)B * heads * nkeys < (1ll << 31), ISP_ERR_UNSUPPORTED,
              "attention_bwd_bf16_tc: too many rows");
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  if (int e = ensure_dynamic_smem((const void*)attnbwd::attention_bwd_kernel<false>, (int)attnbwd::kSmem)) return e;
  if (int e = ensure_dynamic_smem((const void*)attnbwd::attention_bwd_kernel<true>, (int)attnbwd::kSmem)) return e;
  attnbwd::Params p = {};
  p.nkeys = nkeys; p.heads = heads; p.rows = rows; p.HP = HP;
  p.nkb = (nkeys + attnbwd::BK - 1) / attnbwd::BK;
  p.steps_total = (int)((rows + attnbwd::BQ - 1) / attnbwd::BQ);
  // chunk the queries: about 96 steps per CTA (K / V load, dK / dV drain amortised), then the count that wastes the least
  // of the last wave
  const long long base = (long long)B * heads * p.nkb;
  int best = 1;
  double best_eff = -1.0;
  const int qc_hi = p.steps_total < 16 ? 1 : (p.steps_total + 15) / 16;
  const int qc_mid = p.steps_total / 96 > 1 ? p.steps_total / 96 : 1;
  for (int qc = qc_mid > 12 ? qc_mid - 12 : 1; qc <= qc_mid + 12 && qc <= qc_hi; ++qc) {
    const int spc = (p.steps_total + qc - 1) / qc;
    const int qcr = (p.steps_total + spc - 1) / spc;  // chunks that are not empty
    const long long items = base * qcr;
    const long long waves = (items + num_sms - 1) / num_sms;
    const double eff = (double)items / (double)(waves * num_sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = qcr; }
  }
  p.steps_per_chunk = (p.steps_total + best - 1) / best;
  p.qchunks = (p.steps_total + p.steps_per_chunk - 1) / p.steps_per_chunk;
  p.lse = lse; p.dvec = dvec; p.dK = dK; p.dV = dV; p.dQ = dQ; p.lddq = lddq;
  CUtensorMap tmQ, tmDO, tmK, tmV;
  {
    const uint64_t dims[2] = {(uint64_t)ldq, (uint64_t)B * rows}, str[2] = {2, (uint64_t)ldq * 2};
    const uint32_t box[2] = {64, attnbwd::BQ};
    if (int e = make_tmap_bf16(&tmQ, Q, 2, dims, str, box, "attention_bwd(Q)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)lddo, (uint64_t)B * rows}, str[2] = {2, (uint64_t)lddo * 2};
    const uint32_t box[2] = {64, attnbwd::BQ};
    if (int e = make_tmap_bf16(&tmDO, dO, 2, dims, str, box, "attention_bwd(dO)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)HP, (uint64_t)B * heads * nkeys}, str[2] = {2, (uint64_t)HP * 2};
    const uint32_t box[2] = {64, attnbwd::BK};
    if (int e = make_tmap_bf16(&tmK, K, 2, dims, str, box, "attention_bwd(K)")) return e;
    if (int e = make_tmap_bf16(&tmV, V, 2, dims, str, box, "attention_bwd(V)")) return e;
  }
  const long long grid = base * p.qchunks;
  ISP_REQUIRE(grid < (1ll << 31), ISP_ERR_UNSUPPORTED, "attention_bwd_bf16_tc: grid too large");
  if (dQ)
    attnbwd::attention_bwd_kernel<true><<<(unsigned)grid, attnbwd::kThreads, attnbwd::kSmem, as_stream(stream)>>>(tmQ, tmDO, tmK, tmV, p);
  else
    attnbwd::attention_bwd_kernel<false><<<(unsigned)grid, attnbwd::kThreads, attnbwd::kSmem, as_stream(stream)>>>(tmQ, tmDO, tmK, tmV, p);
  ISP_CHECK_LAUNCH("attention_bwd_kernel");
  return ISP_OK;
}
