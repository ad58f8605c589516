// Flash-style attention on tcgen05 (sm_100a): O = softmax(Q K^T) V without ever
// writing the score matrix to HBM.  Used for
//   * LoftUp cross-attention: 200 704 pixel queries x 1024 low-res keys, 4 heads x 101
//     (nn.MultiheadAttention slow path in the reference, loftup/layers.py:186-202, which
//     materialises [B*4, HW, hw] fp32 = 3.29 GB / image / layer);
//   * DINOv2 self-attention: 1025 tokens, 6 heads x 64 (dinov2/layers/attention.py:54-71).
//
// Work item = (128-query tile, head).  Per item, for each block of 128 keys:
//   warp 0 (TMA)   : K block [128 keys x DK] and V^T block [DV x 128 keys] -> smem rings
//   warp 1 (MMA)   : S = Q K^T  -> TMEM (double-buffered, so QK(j+1) overlaps softmax(j));
//                    O += P V   (A = P from smem, B = V^T from smem) -> TMEM
//   warps 2-5      : one thread per query row: tcgen05.ld S, online softmax in fp32
//                    (exp2, lazy rescale of O in TMEM only when the row max grows by > 2^8),
//                    P -> bf16 -> 128B-swizzled smem tile, final O / l -> global bf16
// Q is expected pre-scaled by 1/sqrt(head_dim) (folded into the projection weights).
// Head dims are zero-padded on the K / V^T side only (DK_STEPS*16 >= head_dim, DV >= head_dim).
#include "tc_common.cuh"

namespace isp {
namespace attn {

constexpr int kThreads = 192;
constexpr int BQ = 128, BKEY = 128;
constexpr uint32_t kQBytes = 2 * 16384;          // two 64-column chunks of the Q tile
constexpr uint32_t kKStage = 2 * 16384;          // K block: two 64-column chunks [128 keys x 64]
constexpr uint32_t kVStage = 2 * 16384;          // V^T block: two 64-key chunks [DV x 64] (DV <= 128)
constexpr uint32_t kPBytes = 2 * 16384;          // P tile: two 64-key chunks [128 rows x 64]
constexpr uint32_t kSmem = kQBytes + 2 * kKStage + 2 * kVStage + kPBytes;  // 196608
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThresh = 8.0f;           // log2 units: P stays <= 2^8

struct Params {
  int nkeys, nblocks;          // real keys, blocks of 128 (K / V^T padded with zeros)
  int heads;
  long long rows_per_img;      // queries per image (tile rows never straddle stored rows of 2 images)
  int tiles_per_img;
  long long nitems;            // B * tiles_per_img * heads
  int q_head_stride;           // column offset between heads in Q (elements)
  void* out;                   // bf16 [B*rows_per_img, ldo]
  int ldo, o_head_stride;      // column offset between heads in out
};

__device__ __forceinline__ float ex2_approx(float x) {  // single MUFU.EX2 (ftz); inputs are <= 8
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DK_CHUNKS, int DK_STEPS, int DV>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t q_full, q_empty, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2],
      s_empty[2], p_full, p_empty, o_done, o_empty;
  __shared__ uint32_t tmem_base_s;

  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + 2 * kKStage;
  uint8_t* sP = sV + 2 * kVStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nblocks;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmK); tc::prefetch_tmap(&tmV);
    tc::mbar_init(&q_full, 1); tc::mbar_init(&q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&k_full[i], 1); tc::mbar_init(&k_empty[i], 1);
      tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 4);
    }
    tc::mbar_init(&p_full, 4); tc::mbar_init(&p_empty, 1);
    tc::mbar_init(&o_done, 1); tc::mbar_init(&o_empty, 4);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS0 = tmem, tO = tmem + 256;  // S buffers at columns 0 and 128, O at 256

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
            tc::tma_load_2d(sV + s * kVStage + c * 16384, &tmV, &v_full[s], j * BKEY + c * 64, (int)(bh * DV));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_qk = tc::idesc_bf16_f32(BQ, BKEY);
    const uint32_t idesc_pv = tc::idesc_bf16_f32(BQ, DV);
    uint32_t kv_it = 0, item_it = 0, s_it = 0, p_it = 0;
    auto issue_qk = [&](uint32_t kvi, uint32_t si) {
      const int s = kvi & 1;
      tc::mbar_wait(&k_full[s], (kvi >> 1) & 1);
      tc::mbar_wait(&s_empty[si & 1], ((si >> 1) & 1) ^ 1);
      tc::tc_fence_after();
      if (lane == 0) {
        const uint32_t d = tS0 + (si & 1) * 128;
#pragma unroll
        for (int k = 0; k < DK_STEPS; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          tc::umma_bf16(d, tc::smem_desc_k_sw128(tc::smem_u32(sQ) + off),
                        tc::smem_desc_k_sw128(tc::smem_u32(sK) + s * kKStage + off), idesc_qk, k ? 1u : 0u);
        }
        tc::umma_commit(&s_full[si & 1]);
        tc::umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      tc::mbar_wait(&q_full, item_it & 1);
      issue_qk(kv_it, s_it);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) {
          issue_qk(kv_it + j + 1, s_it + j + 1);
        } else if (lane == 0) {
          tc::umma_commit(&q_empty);  // all QK MMAs of this item issued: Q smem free when they retire
        }
        __syncwarp();
        // O += P(j) V(j)
        const uint32_t kvi = kv_it + j;
        const int s = kvi & 1;
        tc::mbar_wait(&p_full, p_it & 1);
        tc::mbar_wait(&v_full[s], (kvi >> 1) & 1);
        if (j == 0) tc::mbar_wait(&o_empty, (item_it & 1) ^ 1);
        tc::tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < BKEY / 16; ++k) {
            const uint32_t aoff = (k >> 2) * 16384 + (k & 3) * 32;
            tc::umma_bf16(tO, tc::smem_desc_k_sw128(tc::smem_u32(sP) + aoff),
                          tc::smem_desc_k_sw128(tc::smem_u32(sV) + s * kVStage + aoff), idesc_pv, (j | k) ? 1u : 0u);
          }
          tc::umma_commit(&p_empty);
          tc::umma_commit(&v_empty[s]);
          if (j == nb - 1) tc::umma_commit(&o_done);
        }
        __syncwarp();
        ++p_it;
      }
      kv_it += nb;
      s_it += nb;
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t s_it = 0, p_it = 0, item_it = 0;
    for (long long it = blockIdx.x; it < p.nitems; it += gridDim.x, ++item_it) {
      const int h = (int)(it % p.heads);
      const long long tq = it / p.heads;
      const int b = (int)(tq / p.tiles_per_img);
      const int qt = (int)(tq % p.tiles_per_img);
      float m_ref = -INFINITY;  // exponent reference, in log2 units (s * log2e)
      float l = 0.f;
      for (int j = 0; j < nb; ++j, ++s_it, ++p_it) {
        tc::mbar_wait(&s_full[s_it & 1], (s_it >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t ts = tS0 + (s_it & 1) * 128 + lane_addr;
        uint32_t pk[64];  // P row as packed bf16 pairs
        float mj = -INFINITY;
        const int kbase = j * BKEY;
        // pass 1: row max of this block.  S is read from TMEM twice (max, then exp) instead of
        // being held in 128 registers; TMEM reads are cheap.  Only the last key block can hold
        // padded keys, so the masked variant is a block-uniform slow path.
        const bool tail = kbase + BKEY > p.nkeys;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tc::tmem_ld32(ts + c * 32, v);
          tc::tmem_ld_wait();
          if (!tail) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mj = fmaxf(mj, __uint_as_float(v[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (kbase + c * 32 + e < p.nkeys) mj = fmaxf(mj, __uint_as_float(v[e]));
          }
        }
        mj *= kLog2e;
        // lazy max update: only move the reference when the row max grew by more than 2^8
        float scale = 1.f;
        bool need = false;
        if (mj > m_ref + kRescaleThresh) {
          scale = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - mj);
          m_ref = mj;
          need = (j > 0);
        }
        float rs0 = 0.f, rs1 = 0.f;
        const float neg_m = -m_ref;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tc::tmem_ld32(ts + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), kLog2e, neg_m));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), kLog2e, neg_m));
            if (tail) {
              const int key = kbase + c * 32 + e;
              if (key >= p.nkeys) p0 = 0.f;
              if (key + 1 >= p.nkeys) p1 = 0.f;
            }
            rs0 += p0;
            rs1 += p1;
            __nv_bfloat162 bb = __floats2bfloat162_rn(p0, p1);
            pk[c * 16 + (e >> 1)] = *reinterpret_cast<uint32_t*>(&bb);
          }
        }
        const float rs = rs0 + rs1;
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&s_empty[s_it & 1]);  // S(j) consumed; QK(j+2) may overwrite it
        l = l * scale + rs;
        // P smem (and O) are free once PV(j-1) has retired
        if (p_it > 0) tc::mbar_wait(&p_empty, (p_it - 1) & 1);
        const bool any_need = __any_sync(0xffffffffu, need);
        if (any_need) {  // rescale this warp's 32 rows of O in TMEM (warp-uniform branch)
          tc::tc_fence_after();
          const float f = need ? scale : 1.f;
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t o[16];
            tc::tmem_ld16(tO + lane_addr + c, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tc::tmem_st16(tO + lane_addr + c, o);
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
        }
        // write P row into the 128B-swizzled K-major tile: chunk = key/64, 16-byte unit u -> u ^ (r & 7)
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int chunk = u >> 3, uu = u & 7;
          uint4 val = make_uint4(pk[u * 4 + 0], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
          *reinterpret_cast<uint4*>(sP + chunk * 16384 + r * 128 + ((uu ^ (r & 7)) << 4)) = val;
        }
        tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full);
      }
      // epilogue: O / l -> global
      tc::mbar_wait(&o_done, item_it & 1);
      tc::tc_fence_after();
      const long long row_local = (long long)qt * BQ + r;
      const bool row_ok = row_local < p.rows_per_img;
      const float inv = 1.f / l;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) +
                           ((long long)b * p.rows_per_img + row_local) * p.ldo + h * p.o_head_stride;
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t o[16];
        tc::tmem_ld16(tO + lane_addr + c, o);
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
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&o_empty);
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
  p.tiles_per_img = (int)((rows_per_img + attn::BQ - 1) / attn::BQ);
  p.nitems = (long long)B * p.tiles_per_img * heads;
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
