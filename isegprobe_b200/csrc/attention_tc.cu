// Flash-style attention on tcgen05 (sm_100a): O = softmax(Q K^T) V without ever
// writing the score matrix to HBM.  Used for
//   * LoftUp cross-attention: 200 704 pixel queries x 1024 low-res keys, 4 heads x 101
//     (nn.MultiheadAttention slow path in the reference, loftup/layers.py:186-202, which
//     materialises [B*4, HW, hw] fp32 = 3.29 GB / image / layer);
//   * DINOv2 self-attention: 1025 tokens, 6 heads x 64 (dinov2/layers/attention.py:54-71).
//
// Work item = (pair of 128-query tiles, head); see the kernel comment for the pipeline.
// Q is expected pre-scaled by 1/sqrt(head_dim) (folded into the projection weights).
// Head dims are zero-padded on the K / V^T side only (DK_STEPS*16 >= head_dim, DV >= head_dim).
#include "tc_common.cuh"

namespace isp {
namespace attn {

constexpr int kSoftmaxWarps = 8;                 // two groups of four: group t owns query tile t of the pair
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr int BQ = 128, BKEY = 128;
constexpr uint32_t kQTile = 2 * 16384;           // one Q tile: two 64-column chunks [128 rows x 64]
constexpr uint32_t kKStage = 2 * 16384;          // K block: two 64-column chunks [128 keys x 64]
constexpr uint32_t kVStage = 2 * 16384;          // V^T block: two 64-key chunks [DV x 64] (DV <= 128)
constexpr uint32_t kSmem = 2 * kQTile + 2 * kKStage + 2 * kVStage;  // 196608
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThresh = 8.0f;           // log2 units: P stays <= 2^8

struct Params {
  int nkeys, nblocks;          // real keys, blocks of 128 (K / V^T padded with zeros)
  int heads;
  long long rows_per_img;      // queries per image (tile rows never straddle stored rows of 2 images)
  int pairs_per_img;           // pairs of 128-query tiles per image
  long long nitems;            // B * pairs_per_img * heads
  int q_head_stride;           // column offset between heads in Q (elements)
  void* out;                   // bf16 [B*rows_per_img, ldo]
  int ldo, o_head_stride;      // column offset between heads in out
};

__device__ __forceinline__ float ex2_approx(float x) {  // single MUFU.EX2 (ftz); inputs are <= 8
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Work item = (pair of 128-query tiles, head).  Both tiles share every K / V^T block that TMA brings
// in; each tile has its own S and O accumulators in TMEM and its own group of four softmax warps, so
// while one tile's rows are in exp2 (MUFU-bound) the tensor pipe runs the other tile's MMAs:
//   warp 0 (TMA)  : Q pair once per item; K block [128 keys x DK], V^T block [DV x 128 keys] -> 2-deep rings
//   warp 1 (MMA)  : S_t = Q_t K^T (TMEM);  O_t += P_t V  with P_t read FROM TMEM (tcgen05.mma A operand),
//                   so probabilities never touch shared memory
//   warps 2-9     : thread per query row: tcgen05.ld S row (once), online softmax in fp32 (exp2, lazy
//                   rescale of O in TMEM only when the row max grows by > 2^8), P -> bf16 -> tcgen05.st
//                   over the first 64 columns of S_t (the row's S values are already in registers),
//                   final O / l -> global bf16.
// TMEM columns: S_0 @0, S_1 @128, O_0 @256, O_1 @384.  tcgen05.mma instructions execute in issue order,
// which orders PV_t(j) (reads P_t) before QK_t(j+1) (overwrites S_t / P_t).
template <int DK_CHUNKS, int DK_STEPS, int DV>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t q_full, q_empty, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2],
      o_done[2];
  __shared__ uint32_t tmem_base_s;

  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * kQTile;
  uint8_t* sV = sK + 2 * kKStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nblocks;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
    tc::mbar_init(&q_full, 1); tc::mbar_init(&q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&k_full[i], 1); tc::mbar_init(&k_empty[i], 1);
      tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_full[i], 4); tc::mbar_init(&o_done[i], 1);
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
        tc::mbar_arrive_expect_tx(&q_full, 2 * DK_CHUNKS * 16384);
        for (int t = 0; t < 2; ++t)
          for (int c = 0; c < DK_CHUNKS; ++c)
            tc::tma_load_2d(sQ + t * kQTile + c * 16384, &tmQ, &q_full, h * p.q_head_stride + c * 64,
                            (int)(row0 + t * BQ));
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
            tc::tma_load_2d(sV + s * kVStage + c * 16384, &tmV, &v_full[s], j * BKEY + c * 64, (int)(bh * DV));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_qk = tc::idesc_bf16_f32(BQ, BKEY);
    const uint32_t idesc_pv = tc::idesc_bf16_f32(BQ, DV);
    uint32_t kv_it = 0, item_it = 0, p_it = 0;
    // S_t = Q_t K(stage s)^T, then signal the softmax group of tile t
    auto issue_qk = [&](int t, int s) {
      if (lane == 0) {
        const uint32_t d = tmem + t * 128;
#pragma unroll
        for (int k = 0; k < DK_STEPS; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          tc::umma_bf16(d, tc::smem_desc_k_sw128(tc::smem_u32(sQ) + t * kQTile + off),
                        tc::smem_desc_k_sw128(tc::smem_u32(sK) + s * kKStage + off), idesc_qk, k ? 1u : 0u);
        }
        tc::umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      tc::mbar_wait(&q_full, item_it & 1);
      {
        const int s = kv_it & 1;
        tc::mbar_wait(&k_full[s], (kv_it >> 1) & 1);
        tc::tc_fence_after();
        issue_qk(0, s);
        issue_qk(1, s);
        if (lane == 0) {
          tc::umma_commit(&k_empty[s]);
          if (nb == 1) tc::umma_commit(&q_empty);
        }
        __syncwarp();
      }
      for (int j = 0; j < nb; ++j, ++p_it) {
        const uint32_t kvi = kv_it + j;
        const int s = kvi & 1, sn = (kvi + 1) & 1;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // O_t += P_t(j) V(j): P_t from TMEM (16 keys = 8 packed columns per MMA), V^T from smem
          tc::mbar_wait(&p_full[t], p_it & 1);
          if (t == 0) tc::mbar_wait(&v_full[s], (kvi >> 1) & 1);
          tc::tc_fence_after();
          if (lane == 0) {
            const uint32_t dO = tmem + 256 + t * 128, aP = tmem + t * 128;
#pragma unroll
            for (int k = 0; k < BKEY / 16; ++k) {
              const uint32_t boff = (k >> 2) * 16384 + (k & 3) * 32;
              tc::umma_bf16_ts(dO, aP + k * 8, tc::smem_desc_k_sw128(tc::smem_u32(sV) + s * kVStage + boff), idesc_pv,
                               (j | k) ? 1u : 0u);
            }
            if (t == 1) tc::umma_commit(&v_empty[s]);
            if (j == nb - 1) tc::umma_commit(&o_done[t]);
          }
          __syncwarp();
          if (j + 1 < nb) {
            if (t == 0) {
              tc::mbar_wait(&k_full[sn], ((kvi + 1) >> 1) & 1);
              tc::tc_fence_after();
            }
            issue_qk(t, sn);
            if (t == 1 && lane == 0) {
              tc::umma_commit(&k_empty[sn]);
              if (j + 2 == nb) tc::umma_commit(&q_empty);  // last QK of the item issued: Q smem free when it retires
            }
            __syncwarp();
          }
        }
      }
      kv_it += nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int t = (warp - 2) >> 2;   // query tile of the pair this group owns
    const int q = warp & 3;          // TMEM lane quarter
    const int r = q * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem + t * 128 + lane_addr, tO = tmem + 256 + t * 128 + lane_addr;
    uint32_t s_it = 0, item_it = 0;
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      const int h = (int)(it % p.heads);
      const long long tq = it / p.heads;
      const int b = (int)(tq / p.pairs_per_img);
      const int pr = (int)(tq % p.pairs_per_img);
      float m_ref = -INFINITY;  // exponent reference, in log2 units (s * log2e)
      float l = 0.f;
      for (int j = 0; j < nb; ++j, ++s_it) {
        tc::mbar_wait(&s_full[t], s_it & 1);
        tc::tc_fence_after();
        uint32_t sv[128];  // the whole S row of this block, read from TMEM once
#pragma unroll
        for (int c = 0; c < 4; ++c) tc::tmem_ld32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        tc::tmem_ld_wait();
        const int kbase = j * BKEY;
        if (kbase + BKEY > p.nkeys) {  // only the last key block can hold padded keys (block-uniform)
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (kbase + e >= p.nkeys) sv[e] = __float_as_uint(-INFINITY);
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four chains: the row max is latency-bound otherwise
#pragma unroll
        for (int e = 0; e < 128; e += 4) {
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
          need = (j > 0);
        }
        if (__any_sync(0xffffffffu, need)) {  // rescale this warp's 32 rows of O_t (PV_t(j-1) has retired: s_full follows it)
          const float f = need ? scale : 1.f;
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t o[16];
            tc::tmem_ld16(tO + c, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tc::tmem_st16(tO + c, o);
          }
        }
        const float neg_m = -m_ref;
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 16 keys -> 8 packed bf16x2 columns of P_t
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c * 16 + e]), kLog2e, neg_m));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c * 16 + e + 1]), kLog2e, neg_m));
            rs[(e >> 1) & 3] += p0 + p1;
            __nv_bfloat162 bb = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
          }
          tc::tmem_st8(tS + c * 8, pk);
        }
        l = l * scale + (rs[0] + rs[1]) + (rs[2] + rs[3]);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full[t]);
      }
      // epilogue: O_t / l -> global
      tc::mbar_wait(&o_done[t], item_it & 1);
      tc::tc_fence_after();
      const long long row_local = (long long)pr * (2 * BQ) + t * BQ + r;
      const bool row_ok = row_local < p.rows_per_img;
      const float inv = 1.f / l;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                           ((long long)b * p.rows_per_img + row_local) * p.ldo + h * p.o_head_stride;
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t o[16];
        tc::tmem_ld16(tO + c, o);
        tc::tmem_ld_wait();
        if (row_ok) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(o[2 * e]) * inv, __uint_as_float(o[2 * e + 1]) * inv);
            w[e] = *reinterpret_cast<uint32_t*>(&bb);
          }
          *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      // the next item's PV_t(0) overwrites O_t only after this group's next p_full arrival, which
      // follows these reads in program order; tcgen05.ld completion is covered by the wait above
      tc::tc_fence_before();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace attn
}  // namespace isp

using namespace isp;

// Q: bf16 [B*rows_per_img, ldq]; head h reads columns [h*q_head_stride, +DK) (zero K padding
// makes any extra columns harmless).  K: bf16 [B, heads, nblocks*128, DKC] (DKC = 64 or 128,
// zero padded).  Vt: bf16 [B, heads, DV, nblocks*128].  out: bf16 [B*rows_per_img, ldo], head h
// writes DV columns at h*o_head_stride.  variant 0: head_dim <= 64 (DV = 64); 1: <= 112 (DV = 112).
extern "C" int isp_attention_bf16_tc(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                                     void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                                     int heads, int nkeys, int variant, isp_stream_t stream) {
  ISP_REQUIRE(Q && K && Vt && out, ISP_ERR_BAD_SHAPE, "attention_bf16_tc: null pointer");
  ISP_REQUIRE(B > 0 && rows_per_img > 0 && heads > 0 && nkeys > 0, ISP_ERR_BAD_SHAPE, "attention_bf16_tc: bad shape");
  ISP_REQUIRE(variant == 0 || variant == 1, ISP_ERR_UNSUPPORTED, "attention_bf16_tc: variant %d", variant);
  const int DV = variant ? 112 : 64, DKC = variant ? 128 : 64;
  ISP_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && o_head_stride % 8 == 0 && q_head_stride % 8 == 0, ISP_ERR_MISALIGNED,
              "attention_bf16_tc: ldq/ldo/q_head_stride/o_head_stride must be multiples of 8 (TMA box starts and "
              "vector stores need 16-byte alignment)");
  ISP_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(Vt) && aligned16(out), ISP_ERR_MISALIGNED,
              "attention_bf16_tc: 16-byte alignment");
  ISP_REQUIRE((long long)B * rows_per_img < (1ll << 31), ISP_ERR_UNSUPPORTED, "attention_bf16_tc: too many rows");
  attn::Params p = {};
  p.nkeys = nkeys;
  p.nblocks = (nkeys + attn::BKEY - 1) / attn::BKEY;
  p.heads = heads;
  p.rows_per_img = rows_per_img;
  p.pairs_per_img = (int)((rows_per_img + 2 * attn::BQ - 1) / (2 * attn::BQ));
  p.nitems = (long long)B * p.pairs_per_img * heads;
  p.q_head_stride = q_head_stride;
  p.out = out; p.ldo = (int)ldo; p.o_head_stride = o_head_stride;
  const long long nkp = (long long)p.nblocks * attn::BKEY;
  CUtensorMap tmQ, tmK, tmV;
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
  static int num_sms = 0;
  static bool attr_set = false;
  if (!num_sms) {
    int dev = 0;
    ISP_CUDA(cudaGetDevice(&dev));
    ISP_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!attr_set) {
    ISP_CUDA(cudaFuncSetAttribute(attn::attention_kernel<2, 7, 112>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)attn::kSmem));
    ISP_CUDA(cudaFuncSetAttribute(attn::attention_kernel<1, 4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)attn::kSmem));
    attr_set = true;
  }
  const int grid = (int)(p.nitems < num_sms ? p.nitems : num_sms);
  if (variant)
    attn::attention_kernel<2, 7, 112><<<grid, attn::kThreads, attn::kSmem, as_stream(stream)>>>(tmQ, tmK, tmV, p);
  else
    attn::attention_kernel<1, 4, 64><<<grid, attn::kThreads, attn::kSmem, as_stream(stream)>>>(tmQ, tmK, tmV, p);
  ISP_CHECK_LAUNCH("attention_kernel");
  return ISP_OK;
}
