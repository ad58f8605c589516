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
#include <atomic>

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

#ifdef ISP_GEMM_TRACE
// developer instrumentation (tools/trace_gemm.py): clock64 stamps of the first 64 tiles of CTA 0
__device__ long long g_trace[4][64][4];
#define ISP_TRACE(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 64) g_trace[role][tile][slot] = clock64(); } while (0)
#define ISP_TRACE_ADD(role, tile, slot, v) do { if (blockIdx.x == 0 && (tile) < 64) g_trace[role][tile][slot] += (v); } while (0)
#else
#define ISP_TRACE(role, tile, slot) do { } while (0)
#define ISP_TRACE_ADD(role, tile, slot, v) do { } while (0)
#endif

constexpr int BM = 128, BK = 64;
#ifndef ISP_GEMM_EPI_WARPS
#define ISP_GEMM_EPI_WARPS 16
#endif
constexpr int kEpiWarps = ISP_GEMM_EPI_WARPS;      // kParts warps per TMEM lane quarter (they split the chunks of its 32 rows)
constexpr int kParts = kEpiWarps / 4;
constexpr int kFirstEpiWarp = 4;                   // warp group 0 = warps 0, 2, 3 TMA producers, warp 1 MMA; the rest epilogue
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);
// Registers: the block is compiled for 96 per thread (65536 / 640); warp group 0 gives most of its share back (setmaxnreg)
// and the epilogue warp groups raise theirs, which removes the spills of the epilogue's per-tile state (their reloads
// were ~16 % of its stall samples)
constexpr int kRegsLow = 32, kRegsEpi = 112;  // 128 x (96 - 32) released = 4 x 128 x (112 - 96) acquired: inc draws only on what dec freed
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 8;
constexpr int kSmemBytes = 224 * 1024;             // operand ring + epilogue staging (4 KB chunks of 32 rows x 128 B, one or
                                                   // two per epilogue warp); also forces one CTA per SM (TMEM: 512 columns)

struct Params {
  long long M;        // rows (GEMM) or Nimg*H*W (conv)
  int N, K;           // K = padded reduction length actually looped (multiple of 64)
  int BN;             // tile width, multiple of 16, <= 256
  int stages;
  int pair;           // 1: a CTA works on TWO vertically adjacent 128-row tiles at once (both TMEM accumulators), loading
                      // the weight tile once for both -- 0.67x the L2->SM operand bytes per MAC (the conv kernels sit at
                      // the ~12 TB/s L2->SM ceiling); the epilogue then does not overlap the next tile's MMAs
  int batched;        // > 0: `batched` independent problems per "head" x `nbatch` "batches" with their own A / W / D
  int nbatch;         // (4-D tensor maps: column, row, head, batch); tiles_m then counts the row tiles of ALL problems
  int tn;             // batched only, bit 0: A stored reduction-major [K][M] (M contiguous), bit 1: W stored [K][N].
                      // Such an operand arrives as 64-row x 64-column TMA boxes (128-byte rows, 128B swizzle) and the
                      // UMMA descriptor walks it "MN-major", as in wgrad_tc.cu; BN is a multiple of 64 with bit 1.
  int cta2;           // 1: launched as clusters of two CTAs that form a CTA pair (tcgen05 cta_group::2).  The pair computes
                      // two vertically adjacent tiles (or tile pairs) of one column tile with ONE 256-row MMA: each CTA
                      // loads its own 128 A rows and only HALF of the weight tile, the tensor cores read both halves.
                      // Per-SM operand fill per MAC x0.67 (x0.78 on top of pair mode) -- the fill rate (~49 B/clk/SM)
                      // is what bounds these kernels -- while every CTA keeps its own accumulators and epilogue.
  int wres;           // 1 (CTA-pair GEMM with a short reduction): the CTA's share of the weight tile, all k-blocks of it, is loaded
                      // ONCE and stays in shared memory; every cluster works on one column tile only and streams just the A
                      // rows.  Per 128 x 208 tile that is 115 KB instead of 208 KB through the ~49 B/clk L2->SM port, which is
                      // what bounded the K ~ 400 GEMMs of the LoftUp transformer (not HBM, not the tensor pipe)
  int epi_bufs;       // staging chunks per epilogue warp: 2, or 1 (pair mode without residual: the 32 KB go to a third
                      // pipeline stage instead)
  int last_steps;     // 16-wide MMA steps that hold real data in the last k-block (GEMM) / last chunk of a tap (conv)
  // conv mode (TW == 0 -> plain GEMM)
  int TW, TH, H, W, cin_chunks;
  int tiles_m, tiles_n;  // tiles_m * tiles_n < 2^31 (checked by the host): tile arithmetic is 32-bit unsigned
  int tiles_w, tiles_h;
  // epilogue
  const float* bias;   // [N] or null
  float alpha;         // out = alpha * act(acc + bias) + resid
  // fused LayerNorm of the A rows (see isp_gemm_bf16_tc_ex): acc -> rstd * (acc - mean * ln_g[n]) before bias / act
  const float* ln_stats;  // [M][ln_slots][2] partial (sum, sum of squares) of every A row, or null
  const float* ln_g;      // [N] column sums of the (gamma-scaled) weights
  int ln_slots;
  float ln_invC, ln_eps;
  // row statistics of the OUTPUT (over columns < N, of the values as stored), for the next fused LayerNorm
  float* stats_out;       // [M][stats_slots][2] or null; slot = n_tile * chunks_per_tile + chunk: no atomics, deterministic
  int stats_slots;
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

// MN-major operand, 128-byte swizzle: 64 contiguous M (or N) elements per reduction row, 8-row groups 1024 B apart
// (SBO), the next 64 elements one 8 KB box further (LBO).  Same canonical layout as wgrad_tc.cu.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((8192u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TileCoord {
  int m0;
  int n0;
  int img, h0, w0;
};

__device__ __forceinline__ TileCoord tile_coord(const Params& p, uint32_t t) {
  TileCoord c;
  const uint32_t tm = t / (uint32_t)p.tiles_n;
  c.n0 = (int)(t - tm * (uint32_t)p.tiles_n) * p.BN;
  c.m0 = tm * BM;
  c.img = c.h0 = c.w0 = 0;
  if (p.batched) {  // problem z = (batch, head), row tile inside it
    const uint32_t per = (uint32_t)p.tiles_m / (uint32_t)(p.batched * p.nbatch);
    const uint32_t z = tm / per;
    c.m0 = (tm % per) * BM;
    c.h0 = (int)(z % p.batched);
    c.img = (int)(z / p.batched);
  } else if (p.TW) {
    const int per_img = p.tiles_w * p.tiles_h;
    c.img = (int)(tm / per_img);
    const int r = (int)(tm % per_img);
    c.h0 = (r / p.tiles_w) * p.TH;
    c.w0 = (r % p.tiles_w) * p.TW;
  }
  return c;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); }

// The kernel is instantiated per (output dtype, activation, residual) so the epilogue is fully
// inlined: no local-memory arrays, no ABI stack traffic (the earlier non-inlined epilogue spent
// 60 % of its stalls on STL/LDL of its own arguments).
//
// Epilogue, per warp and 128-byte-wide chunk (CW = 64 bf16 / 32 f32 columns) of its 32 rows:
//   residual chunk: TMA-loaded (by this warp's lane 0, before it waits for the accumulator, so the
//   load overlaps the MMAs of the tile) into the 128B-swizzled staging buffer;
//   tcgen05.ld 32 columns -> bias (from smem) / activation / alpha -> + residual read from the
//   staging buffer (row per thread, swizzle makes it conflict-free) -> written back IN PLACE ->
//   TMA store of the chunk.  All global traffic is bulk; OOB rows / columns are clipped by TMA.
template <bool OUT_BF16, int ACT, bool RESID, bool CTA2>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmDt,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmRt,
               const __grid_constant__ CUtensorMap tmBh, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ __align__(8) uint64_t resid_bar[kEpiWarps][2];
  __shared__ __align__(8) uint64_t wfull_bar;
  __shared__ __align__(16) float bias_s[256], lng_s[256];  // single-buffered: two epilogue barriers per tile
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_half = BM * BK * 2;
  // this CTA's share of the weight tile: all of it, or (CTA pair) its half
  const uint32_t a_bytes = a_half << p.pair, b_bytes = ((uint32_t)p.BN >> (CTA2 ? 1 : 0)) * BK * 2;
  const uint32_t stage_bytes = a_bytes + (p.wres ? 0u : b_bytes);
  // work unit of a CTA = one tile, or (pair mode) two vertically adjacent tiles; the two CTAs of a CTA pair take
  // adjacent units of the same column tile.  tiles_m is a multiple of the tiles per cluster unit.
  const uint32_t crank = CTA2 ? tc::cluster_ctarank() : 0u;
  const int gshift = p.pair + (CTA2 ? 1 : 0);
  const uint32_t ntiles = (uint32_t)(p.tiles_m >> gshift) * (uint32_t)p.tiles_n;   // units per CTA stream
  const uint32_t u0 = blockIdx.x >> (CTA2 ? 1 : 0), ustep = gridDim.x >> (CTA2 ? 1 : 0);
  auto unit_tile = [&](uint32_t u, int sub) -> uint32_t {
    return gshift ? ((((u / p.tiles_n) << gshift) + (crank << p.pair) + sub) * p.tiles_n + u % p.tiles_n) : u;
  };
  const int kblocks = p.K / BK;
  uint8_t* epi_smem = smem + kSmemBytes - kEpiWarps * 4096 * p.epi_bufs;
  uint8_t* wres_smem = epi_smem - (size_t)kblocks * b_bytes;  // resident weights (wres): k-block kb at + kb * b_bytes

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::prefetch_tmap(&tmD);
    if (RESID) tc::prefetch_tmap(&tmR);
    // CTA pair: the leader's full barrier also counts the peer's producer, its tempty barrier the peer's epilogue warps
    const uint32_t nroles = 1u + (p.wres ? 0u : 1u) + (uint32_t)p.pair;  // producer threads arriving on a full barrier (per CTA)
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], nroles << (CTA2 ? 1 : 0)); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull_bar[a], 1); tc::mbar_init(&tempty_bar[a], kEpiWarps << (CTA2 ? 1 : 0)); }
    for (int w = 0; w < kEpiWarps; ++w) { tc::mbar_init(&resid_bar[w][0], 1); tc::mbar_init(&resid_bar[w][1], 1); }
    tc::mbar_init(&wfull_bar, 1u << (CTA2 ? 1 : 0));
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) tc::tmem_alloc_2sm(&tmem_base_s, kTmemCols);
    else tc::tmem_alloc(&tmem_base_s, kTmemCols);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CTA2) tc::cluster_sync();  // the peer's barriers exist before anything is signalled across the pair
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < kFirstEpiWarp) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLow));
  if (warp == 0 || warp == 2 || warp == 3) {
    // ------------------------------------------------------------- TMA producers
    // THREE issuing threads, one per operand stream (role 0 on warp 0: the A tile; role 1 on warp 2: the weight tile; role 2
    // on warp 3: the second A tile of pair mode).  One thread gets a tensor load accepted only every ~400 cycles whatever
    // the box size (tools/micro/tma_row_rate.cu: 407 cycles per `cp.async.bulk.tensor` from one thread for 2 KB .. 16 KB
    // boxes, 204 with two issuing warps, 103 with four), so a single producer thread paced a k-block at 2 x 407 cycles
    // (GEMM) or 3 x 407 (conv pair) against 416 / 832 cycles of MMA work -- the "operand-stream" bound of round 1.
    // Every role arrives on the stage's full barrier with its own byte count (barrier count = number of roles).
    const int role = warp == 0 ? 0 : warp - 1;
    const bool active = role == 0 || (role == 1 && !p.wres) || (role == 2 && p.pair);
    if (lane == 0 && active) {
      uint32_t ps = 0, pph = 0;  // ring position of the producer
      if (role == 0 && CTA2 && p.wres && u0 < ntiles) {  // every unit of this cluster has the same column tile
        const int n0 = tile_coord(p, unit_tile(u0, 0)).n0 + (int)crank * (p.BN >> 1);
        if (crank == 0) tc::mbar_arrive_expect_tx(&wfull_bar, 2u * kblocks * b_bytes);
        else tc::mbar_arrive_leader(&wfull_bar);
        for (int kb = 0; kb < kblocks; ++kb) tc::tma_load_2d_2sm(wres_smem + (size_t)kb * b_bytes, &tmBh, &wfull_bar, kb * BK, n0);
      }
      const uint32_t my_bytes = role == 1 ? b_bytes : a_half;
      for (uint32_t t = u0; t < ntiles; t += ustep) {
        const TileCoord tc_ = tile_coord(p, unit_tile(t, role == 2 ? 1 : 0));
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = (int)ps;
          const uint32_t ph = pph;
          if (++ps == (uint32_t)p.stages) { ps = 0; pph ^= 1; }
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes + (role == 2 ? a_half : 0u);
          uint8_t* sb = smem + (size_t)s * stage_bytes + a_bytes;
          if (CTA2) {
            // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
            if (crank == 0) tc::mbar_arrive_expect_tx(&full_bar[s], 2 * my_bytes);
            else tc::mbar_arrive_leader(&full_bar[s]);
            if (role == 1) {
              tc::tma_load_2d_2sm(sb, &tmBh, &full_bar[s], kb * BK, tc_.n0 + (int)crank * (p.BN >> 1));
            } else if (p.TW) {
              const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
              const int r = tap / 3, q = tap - r * 3;
              tc::tma_load_4d_2sm(sa, &tmA, &full_bar[s], cc * BK, tc_.w0 + q - 1, tc_.h0 + r - 1, tc_.img);
            } else {
              tc::tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * BK, (int)tc_.m0);
            }
            continue;
          }
          tc::mbar_arrive_expect_tx(&full_bar[s], my_bytes);
          if (p.batched) {
            if (role == 0) {
              if (p.tn & 1) {  // boxes of 64 reduction rows x 64 columns
                for (int i = 0; i < BM / 64; ++i)
                  tc::tma_load_4d(sa + i * 8192, &tmA, &full_bar[s], (int)tc_.m0 + i * 64, kb * BK, tc_.h0, tc_.img);
              } else {
                tc::tma_load_4d(sa, &tmA, &full_bar[s], kb * BK, (int)tc_.m0, tc_.h0, tc_.img);
              }
            } else if (p.tn & 2) {
              for (int i = 0; i < p.BN / 64; ++i)
                tc::tma_load_4d(sb + i * 8192, &tmB, &full_bar[s], tc_.n0 + i * 64, kb * BK, tc_.h0, tc_.img);
            } else {
              tc::tma_load_4d(sb, &tmB, &full_bar[s], kb * BK, tc_.n0, tc_.h0, tc_.img);
            }
            continue;
          }
          if (role == 1) {
            tc::tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, tc_.n0);
          } else if (p.TW) {
            const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            const int r = tap / 3, q = tap - r * 3;
            tc::tma_load_4d(sa, &tmA, &full_bar[s], cc * BK, tc_.w0 + q - 1, tc_.h0 + r - 1, tc_.img);
          } else {
            tc::tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, (int)tc_.m0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (CTA pair: the leader CTA only)
    // The loop stays warp-converged and everything it computes is warp-uniform (ring position kept as a counter, no
    // division; descriptors = one add on a constant), with only the tcgen05 instructions themselves predicated on the
    // elected lane: ptxas then keeps the descriptors in uniform registers.  The earlier form (lane 0 inside a divergent
    // branch, `it % stages`) cost ~250 instructions = ~1000 cycles per k-block against 416 cycles of MMA work, which is
    // what bounded every K ~ 400 GEMM (0.35 ms per 802816 x 404 x 448 product with the epilogue switched off).
    // reduction-major operands: both "MN-major" (instruction-descriptor bits 15 / 16)
    const uint32_t idesc = tc::idesc_bf16_f32(BM << (CTA2 ? 1 : 0), p.BN) | ((p.tn & 1) ? (1u << 15) : 0u) |
                           ((p.tn & 2) ? (1u << 16) : 0u);
    const bool leader = tc::elect_one();
    const int last_in = p.TW ? p.cin_chunks - 1 : kblocks - 1;  // k-block (within a tap / the row) that may be partial
    const int period = p.TW ? p.cin_chunks : kblocks;
    const uint32_t ring_lo = tc::smem_u32(smem), wres_lo = tc::smem_u32(wres_smem);
    const bool wres = CTA2 && p.wres;
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accum) {
      if (CTA2) tc::umma_bf16_2sm(d, a, b, idesc, accum);
      else tc::umma_bf16(d, a, b, idesc, accum);
    };
    auto commit = [&](uint64_t* bar) {
      if (CTA2) tc::umma_commit_2sm(bar);
      else tc::umma_commit(bar);
    };
    if (wres && crank == 0 && u0 < ntiles) {
      tc::mbar_wait(&wfull_bar, 0);
      tc::tc_fence_after();
    }
    uint32_t tl = 0, s = 0, ph = 0;
    if (!p.tn && !p.pair && crank == 0) {
      // Fast path (K-major operands, one tile per unit: every plain GEMM).  Per k-block: one barrier probe on a precomputed
      // address, four MMAs whose descriptors differ by a constant in the low word, one commit -- ~30 instructions.  The
      // generic loop below spends ~85 on the same work, and the uniform datapath runs them at ~6-9 cycles each: more than
      // the 416 cycles the four MMAs take, so the issuing warp, not the tensor pipe, set the pace of the K ~ 400 GEMMs.
      const uint32_t full0 = tc::smem_u32(&full_bar[0]), empty0 = tc::smem_u32(&empty_bar[0]);
      const uint32_t tfull0 = tc::smem_u32(&tfull_bar[0]), tempty0 = tc::smem_u32(&tempty_bar[0]);
      const uint32_t a_lo0 = ((ring_lo & 0x3FFFFu) >> 4) | (1u << 16);               // descriptor low word of stage 0's A tile
      const uint32_t b_lo0 = wres ? (((wres_lo & 0x3FFFFu) >> 4) | (1u << 16)) : a_lo0 + (a_bytes >> 4);
      const uint32_t a_step = stage_bytes >> 4, b_step = wres ? (b_bytes >> 4) : a_step;
      const int last_steps = p.last_steps;
      const uint32_t nstages = (uint32_t)p.stages;
      uint32_t a_lo = a_lo0, b_ring = b_lo0;
      for (uint32_t t = u0; t < ntiles; t += ustep, ++tl) {
        const uint32_t acc = tl & 1;
        if (leader) ISP_TRACE(0, tl, 0);
        tc::mbar_wait_u32(tempty0 + 8u * acc, ((tl >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        if (leader) ISP_TRACE(0, tl, 1);
        const uint32_t d_tmem = tmem_base + acc * 256u;
        uint32_t b_lo = wres ? b_lo0 : b_ring;
        for (int kb = 0; kb < kblocks; ++kb) {
#ifdef ISP_GEMM_TRACE
          const long long w0 = clock64();
#endif
          tc::mbar_wait_u32(full0 + 8u * s, ph);
          tc::tc_fence_after();
#ifdef ISP_GEMM_TRACE
          if (leader) ISP_TRACE_ADD(0, tl, 3, clock64() - w0);
#endif
          if (leader) {
            if (kb + 1 < kblocks || last_steps == 4) {
              tc::umma_bf16_lo<CTA2>(d_tmem, a_lo, b_lo, idesc, kb ? 1u : 0u);
              tc::umma_bf16_lo<CTA2>(d_tmem, a_lo + 2, b_lo + 2, idesc, 1u);
              tc::umma_bf16_lo<CTA2>(d_tmem, a_lo + 4, b_lo + 4, idesc, 1u);
              tc::umma_bf16_lo<CTA2>(d_tmem, a_lo + 6, b_lo + 6, idesc, 1u);
            } else {  // the zero-filled K tail of the last k-block is skipped
              tc::umma_bf16_lo<CTA2>(d_tmem, a_lo, b_lo, idesc, kb ? 1u : 0u);
              if (last_steps > 1) tc::umma_bf16_lo<CTA2>(d_tmem, a_lo + 2, b_lo + 2, idesc, 1u);
              if (last_steps > 2) tc::umma_bf16_lo<CTA2>(d_tmem, a_lo + 4, b_lo + 4, idesc, 1u);
            }
            tc::umma_commit_u32<CTA2>(empty0 + 8u * s);
            if (kb + 1 == kblocks) { tc::umma_commit_u32<CTA2>(tfull0 + 8u * acc); ISP_TRACE(0, tl, 2); }
          }
          __syncwarp();
          a_lo += a_step;
          b_ring += a_step;
          b_lo += b_step;
          if (++s == nstages) { s = 0; ph ^= 1; a_lo = a_lo0; b_ring = b_lo0; }
          if (!wres) b_lo = b_ring;
        }
      }
    }
    for (uint32_t t = u0; t < ntiles && crank == 0 && (p.tn || p.pair); t += ustep, ++tl) {
      // pair mode: the unit owns both accumulators (virtual tiles 2*tl and 2*tl+1 of the epilogue's numbering)
      const uint32_t acc = p.pair ? 0u : (tl & 1), aph = p.pair ? (tl & 1) : ((tl >> 1) & 1);
      tc::mbar_wait(&tempty_bar[acc], aph ^ 1);
      if (p.pair) tc::mbar_wait(&tempty_bar[1], aph ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;  // accumulator stages at columns 0 and 256
      int kin = 0;
      uint32_t wb = wres_lo;
      for (int kb = 0; kb < kblocks; ++kb) {
        tc::mbar_wait(&full_bar[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = ring_lo + s * stage_bytes;
        // the zero-filled K tail of a row / tap is skipped: those MMAs would add exact zeros
        const int nsteps = (kin == last_in) ? p.last_steps : BK / 16;
        const uint32_t accum = kb ? 1u : 0u;
        if (p.tn) {  // reduction-major operand: 16 reduction rows = 2048 B per step inside the 64-row boxes
          const uint64_t ta = (p.tn & 1) ? smem_desc_mn_sw128(sa) : tc::smem_desc_k_sw128(sa);
          const uint64_t tb = (p.tn & 2) ? smem_desc_mn_sw128(sa + a_bytes) : tc::smem_desc_k_sw128(sa + a_bytes);
          const uint64_t ia = (p.tn & 1) ? 128 : 2, ib = (p.tn & 2) ? 128 : 2;
          if (leader) {
            mma(d_tmem, ta, tb, accum);
            if (nsteps > 1) mma(d_tmem, ta + ia, tb + ib, 1u);
            if (nsteps > 2) mma(d_tmem, ta + 2 * ia, tb + 2 * ib, 1u);
            if (nsteps > 3) mma(d_tmem, ta + 3 * ia, tb + 3 * ib, 1u);
          }
        } else {
          // K-major, 128B swizzle: +32 B (descriptor + 2) per 16-element K step inside the swizzle row
          const uint64_t adesc = tc::smem_desc_k_sw128(sa);
          const uint64_t bdesc = tc::smem_desc_k_sw128(wres ? wb : sa + a_bytes);
          if (nsteps == BK / 16) {
            if (leader) {
              mma(d_tmem, adesc, bdesc, accum);
              mma(d_tmem, adesc + 2, bdesc + 2, 1u);
              mma(d_tmem, adesc + 4, bdesc + 4, 1u);
              mma(d_tmem, adesc + 6, bdesc + 6, 1u);
            }
          } else if (leader) {
            mma(d_tmem, adesc, bdesc, accum);
            if (nsteps > 1) mma(d_tmem, adesc + 2, bdesc + 2, 1u);
            if (nsteps > 2) mma(d_tmem, adesc + 4, bdesc + 4, 1u);
          }
          if (p.pair) {  // second tile of the pair: A rows 128..255 of the stage (+16 KB), same weight tile
            const uint64_t adesc2 = tc::smem_desc_k_sw128(sa + a_half);
            if (leader) {
              mma(d_tmem + 256u, adesc2, bdesc, accum);
              if (nsteps > 1) mma(d_tmem + 256u, adesc2 + 2, bdesc + 2, 1u);
              if (nsteps > 2) mma(d_tmem + 256u, adesc2 + 4, bdesc + 4, 1u);
              if (nsteps > 3) mma(d_tmem + 256u, adesc2 + 6, bdesc + 6, 1u);
            }
          }
        }
        if (leader) {
          commit(&empty_bar[s]);          // smem slot free (CTA pair: in both CTAs) when these MMAs retire
          if (kb == kblocks - 1) {        // accumulator(s) complete (in both CTAs)
            commit(&tfull_bar[acc]);
            if (p.pair) commit(&tfull_bar[1]);
          }
        }
        __syncwarp();
        if (++kin == period) kin = 0;
        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        wb += b_bytes;
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ------------------------------------------------------------- epilogue (warps kFirstEpiWarp ..)
    // The warps are independent of each other: each owns the chunks part, part + kParts, ... of its 32 rows of every tile
    // ("items"), with its own staging buffer(s), residual barriers and TMA store groups; the only CTA-wide step is the
    // reload of the bias slice when a tile starts at another column than the previous one.
    constexpr int CW = OUT_BF16 ? 64 : 32;    // columns per 128-byte chunk
    constexpr int ESZ = OUT_BF16 ? 2 : 4;
    constexpr int UPS = OUT_BF16 ? 4 : 8;     // 16-byte units per 32 accumulator columns
    constexpr int CPU_ = 16 / ESZ;            // columns per 16-byte unit
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    const int ew = warp - kFirstEpiWarp;
    const int q = warp & 3;      // TMEM lane quarter this warp may access == its 32 rows of the tile
    const int part = ew >> 2;    // which of the kParts warps sharing that quarter
    const int etid = ew * 32 + lane;
    uint8_t* stage = epi_smem + ew * 4096 * p.epi_bufs;
    uint64_t* rbar = resid_bar[ew];
    const int nchunks = (p.BN + CW - 1) / CW;
    const int n_my = (nchunks - part + kParts - 1) / kParts;  // 0: this warp only takes part in the accumulator hand-over
    const int row = lane, r7 = lane & 7;
    const uint32_t bmask = (uint32_t)p.epi_bufs - 1u;
    uint32_t tl = 0, st_seq = 0, rph0 = 0, rph1 = 0;
    int bias_n0 = -1;
    // where this warp's 32 rows of a tile live in the output tensor
    auto out_coords = [&](const TileCoord& t, int& c1, int& c2, int& c3) {
      c2 = c3 = 0;
      if (p.TW) {
        const int bw = p.TW < 32 ? p.TW : 32;
        const int pix = q * 32;
        c1 = t.w0 + (p.TW >= 32 ? pix % p.TW : 0);
        c2 = t.h0 + (p.TW >= 32 ? pix / p.TW : q * (32 / bw));
        c3 = t.img;
      } else {
        c1 = (int)t.m0 + q * 32;
        if (p.batched) { c2 = t.h0; c3 = t.img; }
      }
    };
    // this thread's row of a tile: global row index (conv mode: linear pixel index) and validity
    auto row_of = [&](const TileCoord& t, long long& grow, bool& ok) {
      if (p.TW) {
        const int pix = q * 32 + lane;
        const int hh = t.h0 + pix / p.TW, ww = t.w0 + pix % p.TW;
        ok = hh < p.H && ww < p.W;
        grow = ((long long)t.img * p.H + hh) * p.W + ww;
      } else {
        grow = t.m0 + q * 32 + lane;
        ok = grow < p.M;
      }
    };
    // residual chunk i of the tile at t -> staging buffer seq & bmask (one lane)
    auto issue_resid = [&](const TileCoord& t, int i, uint32_t seq) {
      int c1, c2, c3;
      out_coords(t, c1, c2, c3);
      const int col0 = (part + kParts * i) * CW;
      const int ncols = min(CW, p.BN - col0);
      const uint32_t b = seq & bmask;
      uint8_t* buf = stage + b * 4096;
      tc::mbar_arrive_expect_tx(&rbar[b], 32u * ncols * ESZ);
      if (p.TW) tc::tma_load_4d(buf, ncols < CW ? &tmRt : &tmR, &rbar[b], t.n0 + col0, c1, c2, c3);
      else tc::tma_load_2d(buf, ncols < CW ? &tmRt : &tmR, &rbar[b], t.n0 + col0, c1);
    };
    // LayerNorm statistics of a row are loaded ONE TILE AHEAD (an even number of slots <= kLnPf, as float4 pairs kept in
    // registers): loaded at the start of their own tile they cost the warp a global-memory latency per tile before it
    // could touch the accumulator
    constexpr int kLnPf = 8;
    const bool ln_on = p.ln_stats != nullptr && n_my > 0;
    const bool ln_pf = ln_on && p.ln_slots <= kLnPf && (p.ln_slots & 1) == 0;
    const int ln_pairs = p.ln_slots >> 1;
    const bool has_alpha = p.alpha != 1.f;
    const float2 alpha2 = make_float2(p.alpha, p.alpha);
    float4 sv[kLnPf / 2];
    auto load_stats = [&](long long grow, bool ok) {
      const float4* sp4 = reinterpret_cast<const float4*>(p.ln_stats + grow * (2 * p.ln_slots));
#pragma unroll
      for (int k = 0; k < kLnPf / 2; ++k) sv[k] = (ok && k < ln_pairs) ? __ldg(sp4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // The per-tile bookkeeping is on this warp's critical path (one item of 32 x 64 outputs per tile): unit coordinates advance
    // incrementally (unit u = um * tiles_n + un), no division per tile outside the conv / batched index split
    const uint32_t tn_u = (uint32_t)p.tiles_n;
    const uint32_t step_m = ustep / tn_u, step_n = ustep - step_m * tn_u;
    uint32_t u = u0, um = u0 / tn_u, un = u0 - um * tn_u;
    int sub = 0;
    auto coord_of = [&](uint32_t um_, uint32_t un_, int sub_) -> TileCoord {
      const uint32_t tm = gshift ? ((um_ << gshift) + (crank << p.pair) + (uint32_t)sub_) : um_;
      TileCoord c;
      c.n0 = (int)un_ * p.BN;
      c.m0 = (int)(tm * BM);
      c.img = c.h0 = c.w0 = 0;
      if (p.batched) {
        const uint32_t per = (uint32_t)p.tiles_m / (uint32_t)(p.batched * p.nbatch);
        const uint32_t z = tm / per;
        c.m0 = (int)((tm - z * per) * BM);
        c.img = (int)(z / (uint32_t)p.batched);
        c.h0 = (int)(z - (uint32_t)c.img * (uint32_t)p.batched);
      } else if (p.TW) {
        const uint32_t per_img = (uint32_t)(p.tiles_w * p.tiles_h);
        c.img = (int)(tm / per_img);
        const uint32_t r = tm - (uint32_t)c.img * per_img;
        const uint32_t rh = r / (uint32_t)p.tiles_w;
        c.h0 = (int)rh * p.TH;
        c.w0 = (int)(r - rh * (uint32_t)p.tiles_w) * p.TW;
      }
      return c;
    };
    TileCoord tc_ = coord_of(um, un, 0);  // garbage beyond the last tile, unused
    long long grow;
    bool row_ok;
    row_of(tc_, grow, row_ok);
    if (u < ntiles) {
      if (ln_pf) load_stats(grow, row_ok);
      if (RESID && p.epi_bufs == 2 && n_my > 0 && lane == 0) issue_resid(tc_, 0, 0);
    }
    for (; u < ntiles; ++tl) {
      const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
      int c1, c2, c3;
      out_coords(tc_, c1, c2, c3);
      // bias / g slice of this column tile -> smem (zeros beyond N, so padded columns come out as exact zeros); the tile
      // sequence is the same for every epilogue warp, so the condition is CTA-uniform
      if (tc_.n0 != bias_n0) {
        epi_bar_sync();  // every epilogue warp is done with the previous slice
        if (etid < p.BN) {
          const int n = tc_.n0 + etid;
          bias_s[etid] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
          lng_s[etid] = (p.ln_stats && n < p.N) ? __ldg(p.ln_g + n) : 0.f;
        }
        epi_bar_sync();
        bias_n0 = tc_.n0;
      }
      float ln_rstd = 1.f, ln_nrm = 0.f;  // x_norm . W = rstd * acc + (-rstd * mean) * g[n]
      if (ln_on && row_ok) {
        float su = 0.f, sq = 0.f;
        if (ln_pf) {
#pragma unroll
          for (int k = 0; k < kLnPf / 2; ++k) { su += sv[k].x; sq += sv[k].y; su += sv[k].z; sq += sv[k].w; }  // fixed order
        } else {
          const float2* sp2 = reinterpret_cast<const float2*>(p.ln_stats) + grow * p.ln_slots;
#pragma unroll 1
          for (int k = 0; k < p.ln_slots; ++k) {
            const float2 v = __ldg(sp2 + k);
            su += v.x;
            sq += v.y;
          }
        }
        const float mean = su * p.ln_invC;
        ln_rstd = rsqrtf(fmaxf(sq * p.ln_invC - mean * mean, 0.f) + p.ln_eps);
        ln_nrm = -ln_rstd * mean;
      }
      // the warp's next tile: coordinates, row, and its statistics on their way
      uint32_t u_next = u, um_next = um, un_next = un;
      int sub_next = 0;
      if (p.pair && sub == 0) {
        sub_next = 1;
      } else {
        u_next = u + ustep;
        um_next = um + step_m;
        un_next = un + step_n;
        if (un_next >= tn_u) { un_next -= tn_u; ++um_next; }
      }
      const bool has_next = u_next < ntiles;
      const TileCoord tc_next = coord_of(um_next, un_next, sub_next);
      long long grow_next;
      bool row_ok_next;
      row_of(tc_next, grow_next, row_ok_next);
      if (ln_pf && has_next) load_stats(grow_next, row_ok_next);
      if (ew == 0 && lane == 0) ISP_TRACE(1, tl, 0);
      if (ew == kEpiWarps - 1 && lane == 0) ISP_TRACE(2, tl, 0);
      tc::mbar_wait(&tfull_bar[acc], aph);
      tc::tc_fence_after();
      if (ew == 0 && lane == 0) ISP_TRACE(1, tl, 1);
      if (ew == kEpiWarps - 1 && lane == 0) ISP_TRACE(2, tl, 1);
      const uint32_t t_addr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);
      const float* bs = bias_s;
      const float* gs = lng_s;
      const float2 rstd2 = make_float2(ln_rstd, ln_rstd), nrm2 = make_float2(ln_nrm, ln_nrm);
      for (int i = 0; i < n_my; ++i, ++st_seq) {
        float2 st_sum2 = make_float2(0.f, 0.f), st_sq2 = st_sum2;  // even / odd columns, added at the end
        const int col0 = (part + kParts * i) * CW;  // column inside the tile
        const int ncols = min(CW, p.BN - col0);     // multiple of 16
        const bool is_tail = ncols < CW;
        const int row_bytes = ncols * ESZ;          // dense row pitch of a tail chunk
        const int nglob = tc_.n0 + col0;            // first global column of the chunk
        const bool full = nglob + ncols <= p.N;     // no column masking needed
        const uint32_t b = st_seq & bmask;
        uint8_t* obuf = stage + b * 4096;
        if (RESID) {
          if (lane == 0) {
            // the store that last used the buffer about to be loaded has read it (a few hundred cycles after its issue)
            tc::tma_store_wait_read<0>();
            if (p.epi_bufs == 2) {
              // two buffers: the residual of the NEXT item (of this tile or of the warp's next tile) is requested one
              // whole item ahead, so its HBM latency is not part of this warp's per-tile time
              if (i + 1 < n_my) issue_resid(tc_, i + 1, st_seq + 1);
              else if (has_next) issue_resid(tc_next, 0, st_seq + 1);
            } else {
              issue_resid(tc_, i, st_seq);
            }
          }
          __syncwarp();
          tc::mbar_wait(&rbar[b], b ? rph1 : rph0);
          if (b) rph1 ^= 1; else rph0 ^= 1;
        } else {
          if (lane == 0) {  // the store that used this buffer (two chunks ago / the previous one) has read it
            if (p.epi_bufs == 2) tc::tma_store_wait_read<1>();
            else tc::tma_store_wait_read<0>();
          }
          __syncwarp();
        }
        uint8_t* orow = obuf + row * (is_tail ? row_bytes : 128);
#pragma unroll
        for (int sb = 0; sb < CW / 32; ++sb) {
          const int scol = sb * 32;
          if (scol < ncols) {
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
            for (int uu = 0; uu < UPS; ++uu) {
              const int u = sb * UPS + uu;          // 16-byte unit inside the chunk row
              if (u * 16 < row_bytes) {
                uint8_t* sp = orow + (is_tail ? (u << 4) : ((u ^ r7) << 4));
                // packed fp32 pairs (FFMA2 / FADD2 / FMUL2): LayerNorm of the A row folded as a = rstd * acc + (nrm * g[n] + bias[n])
                // -- with rstd = 1, nrm = 0 and g = 0 when no LayerNorm is fused, so there is one code path
                float2 x2[CPU_ / 2];
#pragma unroll
                for (int e4 = 0; e4 < CPU_; e4 += 4) {
                  const float4 bb = *reinterpret_cast<const float4*>(bs + col0 + scol + uu * CPU_ + e4);
                  const float4 gg = *reinterpret_cast<const float4*>(gs + col0 + scol + uu * CPU_ + e4);
                  const float2 c0 = __ffma2_rn(nrm2, make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
                  const float2 c1v = __ffma2_rn(nrm2, make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
                  float2 a0 = __ffma2_rn(rstd2, make_float2(__uint_as_float(vr[uu * CPU_ + e4 + 0]),
                                                            __uint_as_float(vr[uu * CPU_ + e4 + 1])), c0);
                  float2 a1 = __ffma2_rn(rstd2, make_float2(__uint_as_float(vr[uu * CPU_ + e4 + 2]),
                                                            __uint_as_float(vr[uu * CPU_ + e4 + 3])), c1v);
                  if constexpr (ACT != 0 && ACT != 5) {
                    a0.x = act_fn<ACT>(a0.x); a0.y = act_fn<ACT>(a0.y);
                    a1.x = act_fn<ACT>(a1.x); a1.y = act_fn<ACT>(a1.y);
                  }
                  if (has_alpha) { a0 = __fmul2_rn(a0, alpha2); a1 = __fmul2_rn(a1, alpha2); }
                  x2[e4 / 2] = a0;
                  x2[e4 / 2 + 1] = a1;
                }
                if (RESID) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(sp);
                  const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
                  if constexpr (ACT == 5) {
                    // ReLU backward: the "residual" operand is the forward activation; keep the gradient where it is > 0
                    if constexpr (OUT_BF16) {
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        if (!(__uint_as_float(rw[k] << 16) > 0.f)) x2[k].x = 0.f;
                        if (!(__uint_as_float(rw[k] & 0xffff0000u) > 0.f)) x2[k].y = 0.f;
                      }
                    } else {
                      if (!(__uint_as_float(rr.x) > 0.f)) x2[0].x = 0.f;
                      if (!(__uint_as_float(rr.y) > 0.f)) x2[0].y = 0.f;
                      if (!(__uint_as_float(rr.z) > 0.f)) x2[1].x = 0.f;
                      if (!(__uint_as_float(rr.w) > 0.f)) x2[1].y = 0.f;
                    }
                  } else if constexpr (OUT_BF16) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // bf16 -> f32 is a shift: the add happens in fp32, one rounding
                      x2[k] = __fadd2_rn(x2[k], make_float2(__uint_as_float(rw[k] << 16), __uint_as_float(rw[k] & 0xffff0000u)));
                  } else {
                    x2[0] = __fadd2_rn(x2[0], make_float2(__uint_as_float(rr.x), __uint_as_float(rr.y)));
                    x2[1] = __fadd2_rn(x2[1], make_float2(__uint_as_float(rr.z), __uint_as_float(rr.w)));
                  }
                }
                if (!full) {
#pragma unroll
                  for (int e = 0; e < CPU_; e += 2) {
                    if (nglob + scol + uu * CPU_ + e >= p.N) x2[e / 2].x = 0.f;
                    if (nglob + scol + uu * CPU_ + e + 1 >= p.N) x2[e / 2].y = 0.f;
                  }
                }
                uint4 w;
                if constexpr (OUT_BF16) {
                  __nv_bfloat162 b0 = __floats2bfloat162_rn(x2[0].x, x2[0].y), b1 = __floats2bfloat162_rn(x2[1].x, x2[1].y);
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(x2[2].x, x2[2].y), b3 = __floats2bfloat162_rn(x2[3].x, x2[3].y);
                  w = make_uint4(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1),
                                 *reinterpret_cast<uint32_t*>(&b2), *reinterpret_cast<uint32_t*>(&b3));
                  if (p.stats_out) {  // statistics of the values the consumer will read (after bf16 rounding)
                    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const float2 v = make_float2(__uint_as_float(ww[k] << 16), __uint_as_float(ww[k] & 0xffff0000u));
                      st_sum2 = __fadd2_rn(st_sum2, v);
                      st_sq2 = __ffma2_rn(v, v, st_sq2);
                    }
                  }
                } else {
                  if (p.stats_out) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) { st_sum2 = __fadd2_rn(st_sum2, x2[k]); st_sq2 = __ffma2_rn(x2[k], x2[k], st_sq2); }
                  }
                  w = make_uint4(__float_as_uint(x2[0].x), __float_as_uint(x2[0].y), __float_as_uint(x2[1].x), __float_as_uint(x2[1].y));
                }
                *reinterpret_cast<uint4*>(sp) = w;
              }
            }
          }
        }
        if (p.stats_out && row_ok) {
          const int slot = (tc_.n0 / p.BN) * nchunks + (part + kParts * i);
          reinterpret_cast<float2*>(p.stats_out)[grow * p.stats_slots + slot] = make_float2(st_sum2.x + st_sum2.y, st_sq2.x + st_sq2.y);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const CUtensorMap* m = is_tail ? &tmDt : &tmD;
          if (p.TW || p.batched) tc::tma_store_4d(m, obuf, nglob, c1, c2, c3);
          else tc::tma_store_2d(m, obuf, nglob, c1);
          tc::tma_store_commit();
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) ISP_TRACE(1, tl, 2);
      if (ew == kEpiWarps - 1 && lane == 0) ISP_TRACE(2, tl, 2);
      if (lane == 0) {  // accumulator drained (CTA pair: only the leader's MMA warp waits, on the leader's barrier)
        if (CTA2) tc::mbar_arrive_leader(&tempty_bar[acc]);
        else tc::mbar_arrive(&tempty_bar[acc]);
      }
      u = u_next; um = um_next; un = un_next; sub = sub_next;
      tc_ = tc_next; grow = grow_next; row_ok = row_ok_next;
    }
    if (lane == 0) tc::tma_store_wait_all();  // global writes complete before the CTA retires
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CTA2) tc::cluster_sync();  // no CTA leaves (or frees TMEM) while its peer may still signal / the pair MMA runs
  if (warp == 1) {
    if (CTA2) tc::tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

static int pick_bn(int N, int max_bn) {
  // widest tile <= max_bn that wastes the least: N itself if it fits, else an even split
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= max_bn) return n16;
  for (int parts = 2; parts <= 64; ++parts) {
    const int bn = ((N + parts - 1) / parts + 15) / 16 * 16;
    if (bn <= max_bn) return bn;
  }
  return max_bn;
}

typedef void (*kernel_fn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Params);

template <bool OB, bool RS, bool C2>
static kernel_fn pick_act(int act) {
  switch (act) {
    case 0: return gemm_tc_kernel<OB, 0, RS, C2>;
    case 1: return gemm_tc_kernel<OB, 1, RS, C2>;
    case 2: return gemm_tc_kernel<OB, 2, RS, C2>;
    case 3: return gemm_tc_kernel<OB, 3, RS, C2>;
    default: return gemm_tc_kernel<OB, 4, RS, C2>;
  }
}

// C2: the CTA-pair instantiation (cta_group::2 instructions: it can only be launched in clusters of two CTAs)
template <bool C2>
static kernel_fn pick_kernel(int out_bf16, int act, bool resid) {
  if (act == 5) return out_bf16 ? gemm_tc_kernel<true, 5, true, C2> : gemm_tc_kernel<false, 5, true, C2>;  // ReLU-mask epilogue
  if (resid) return out_bf16 ? pick_act<true, true, C2>(act) : pick_act<false, true, C2>(act);
  return out_bf16 ? pick_act<true, false, C2>(act) : pick_act<false, false, C2>(act);
}

static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmDt,
                  const CUtensorMap& tmR, const CUtensorMap& tmRt, const CUtensorMap& tmBh, Params& p, int out_bf16,
                  int act, bool resid, cudaStream_t stream) {
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  // pairing halves the number of work units: only when there are at least two waves of pairs
  if (p.pair && (p.batched || (p.tiles_m & 1) || (p.tiles_m / 2) * p.tiles_n < 2LL * num_sms)) p.pair = 0;
  auto prepare = [](kernel_fn f) -> int {  // opt in to 224 KB of dynamic shared memory, once per (device, instantiation)
    return ensure_dynamic_smem((const void*)f, kSmemBytes);
  };
  // CTA-pair mode: tiles per cluster unit divide tiles_m (the weight tile always splits into two swizzle-aligned
  // halves: BN % 16 == 0) and there are at least two waves of cluster units
  // co-resident 2-CTA clusters of the pair kernel (224 KB smem: one CTA per SM); queried once per device
  static std::atomic<int> max_clusters_dev[64];  // zero-initialised: 0 = not queried yet, else the count + 1
  int dev = 0;
  ISP_CUDA(cudaGetDevice(&dev));
  ISP_REQUIRE(dev >= 0 && dev < 64, ISP_ERR_UNSUPPORTED, "gemm_tc: device ordinal %d", dev);
  int max_clusters = max_clusters_dev[dev].load() - 1;
  if (max_clusters < 0) {
    kernel_fn f2 = pick_kernel<true>(out_bf16, act, resid);
    if (int e = prepare(f2)) return e;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(num_sms & ~1); q.blockDim = dim3(kThreads); q.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, f2, &q) != cudaSuccess) { n = 0; cudaGetLastError(); }
    max_clusters = n;
    max_clusters_dev[dev].store(n + 1);
  }
  const int G = 2 << p.pair;
  p.cta2 = (!p.batched && 2 * max_clusters >= num_sms - 4 && p.tiles_m % G == 0 &&
            (p.tiles_m / G) * p.tiles_n >= 2LL * max_clusters) ? 1 : 0;
  // staging chunks per epilogue warp: one, or two where the second hides latency: the residual of the next item is
  // prefetched into it (pair mode exposes its epilogue anyway and needs the shared memory for its wider operand stages)
  p.epi_bufs = (resid && !p.pair) ? 2 : 1;
  // resident weights: plain CTA-pair GEMM without residual whose share of the weight tile (all k-blocks) fits next to the
  // staging chunks and a >= 5-stage ring of A tiles (fewer leave the MMA warp waiting for operands: the TMA round trip is
  // ~3000 cycles; 802816 x 448 x 404 with 3 stages: 4000 cycles of operand wait per 128 x 224 tile, with 5: 1000)
  const int a_bytes = BM * BK * 2 << p.pair;
  const int w_bytes = (p.K / BK) * (p.BN >> 1) * BK * 2;
  p.wres = (p.cta2 && !p.pair && !p.TW && !resid && max_clusters >= p.tiles_n &&
            kSmemBytes - kEpiWarps * 4096 - w_bytes >= 5 * a_bytes) ? 1 : 0;
  const int stage_bytes = a_bytes + (p.wres ? 0 : (p.BN >> p.cta2) * BK * 2);
  p.stages = (kSmemBytes - kEpiWarps * 4096 * p.epi_bufs - (p.wres ? w_bytes : 0)) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  ISP_REQUIRE(p.stages >= 2, ISP_ERR_UNSUPPORTED, "gemm_tc: tile too large for a 2-stage pipeline");
  if (p.cta2) {
    kernel_fn fn = pick_kernel<true>(out_bf16, act, resid);
    if (int e = prepare(fn)) return e;
    const long long nunits = (p.tiles_m / G) * p.tiles_n;
    int ncl = (int)(nunits < max_clusters ? nunits : max_clusters);
    if (p.wres) ncl = ncl / (int)p.tiles_n * (int)p.tiles_n;  // cluster c then only ever sees column tile c % tiles_n
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ncl); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ISP_CUDA(cudaLaunchKernelEx(&cfg, fn, tmA, tmB, tmD, tmDt, tmR, tmRt, tmBh, p));
    ::isp::count_launch(1);
    return ISP_OK;
  }
  kernel_fn fn = pick_kernel<false>(out_bf16, act, resid);
  if (int e = prepare(fn)) return e;
  const long long ntiles = (p.tiles_m >> p.pair) * p.tiles_n;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  fn<<<grid, kThreads, kSmemBytes, stream>>>(tmA, tmB, tmD, tmDt, tmR, tmRt, tmBh, p);
  ISP_CHECK_LAUNCH("gemm_tc_kernel");
  return ISP_OK;
}

}  // namespace gemm
}  // namespace isp

using namespace isp;

static int gemm_common(const void* A, long long lda, const void* W, long long ldw, const float* bias, const void* resid,
                       int resid_bf16, long long ldr, float alpha, int act, void* D, long long ldd, int out_bf16,
                       long long M, int N, int K, const float* ln_stats, int ln_slots, const float* ln_g, float ln_eps,
                       float* stats_out, int stats_slots, isp_stream_t stream) {
  ISP_REQUIRE(A && W && D, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: null pointer");
  ISP_REQUIRE(M > 0 && N > 0 && K > 0, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: bad shape M=%lld N=%d K=%d", M, N, K);
  ISP_REQUIRE(lda >= K && ldw >= K && ldd >= N, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc: leading dimensions too small");
  // the epilogue stores whole 32 / 64-column chunks, clipped at column ldd by the tensor map, with columns >= N zeroed: a D
  // that is a column slice of a wider matrix (ldd beyond the padding of N) would have its neighbours overwritten with zeros
  ISP_REQUIRE(ldd <= (N + 15) / 16 * 16, ISP_ERR_BAD_SHAPE,
              "gemm_bf16_tc: ldd = %lld exceeds N = %d rounded up to 16: D[:, N:ldd) is zero-filled, so D must own its padding "
              "columns (not be a column slice of a wider matrix)", ldd, N);
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
  // with a residual every epilogue warp should own at most two 128-byte chunks per tile, so that
  // both residual loads are in flight before the accumulator is ready (f32: 32 columns per chunk)
  p.BN = gemm::pick_bn(N, (resid && !out_bf16) ? 128 : 256);
  p.last_steps = (K - (p.K - gemm::BK) + 15) / 16;
  p.TW = 0;
  p.tiles_m = (int)((M + gemm::BM - 1) / gemm::BM);
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.pair = 0;  // K is short here: the exposed epilogue costs more than the shared weight tile saves (measured)
  p.bias = bias; p.alpha = alpha;
  if (ln_stats) {
    ISP_REQUIRE(ln_g && ln_slots > 0, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc_ex: ln_stats needs ln_g and ln_slots > 0");
    p.ln_stats = ln_stats; p.ln_g = ln_g; p.ln_slots = ln_slots; p.ln_invC = 1.f / (float)K; p.ln_eps = ln_eps;
  }
  if (stats_out) {
    const int need = (int)p.tiles_n * ((p.BN + (out_bf16 ? 64 : 32) - 1) / (out_bf16 ? 64 : 32));
    ISP_REQUIRE(stats_slots == need, ISP_ERR_BAD_SHAPE,
                "gemm_bf16_tc_ex: stats_slots must be %d for N=%d (ask isp_gemm_stats_slots)", need, N);
    p.stats_out = stats_out; p.stats_slots = stats_slots;
  }
  CUtensorMap tmA, tmB, tmD, tmBh;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {2, (uint64_t)lda * 2};
    const uint32_t box[2] = {gemm::BK, gemm::BM};
    if (int e = make_tmap_bf16(&tmA, A, 2, dims, str, box, "gemm_bf16_tc(A)")) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[2] = {2, (uint64_t)ldw * 2};
    const uint32_t box[2] = {gemm::BK, (uint32_t)p.BN};
    if (int e = make_tmap_bf16(&tmB, W, 2, dims, str, box, "gemm_bf16_tc(W)")) return e;
    const uint32_t boxh[2] = {gemm::BK, (uint32_t)p.BN / 2};  // half tile (CTA-pair mode: one half per CTA)
    if (int e = make_tmap_bf16(&tmBh, W, 2, dims, str, boxh, "gemm_bf16_tc(W half)")) return e;
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
  CUtensorMap tmR = tmD, tmRt = tmDt;
  if (resid) {
    const uint64_t dims[2] = {(uint64_t)ldr, (uint64_t)M}, str[2] = {(uint64_t)esz, (uint64_t)ldr * esz};
    const uint32_t box[2] = {cw, 32}, boxt[2] = {tailw ? tailw : cw, 32};
    if (int e = make_tmap(&tmR, esz, resid, 2, dims, str, box, "gemm_bf16_tc(resid)", true)) return e;
    if (int e = make_tmap(&tmRt, esz, resid, 2, dims, str, boxt, "gemm_bf16_tc(resid tail)", false)) return e;
  }
  return gemm::launch(tmA, tmB, tmD, tmDt, tmR, tmRt, tmBh, p, out_bf16, act, resid != nullptr, as_stream(stream));
}

#ifdef ISP_GEMM_TRACE
extern "C" int isp_gemm_trace_read(long long* host_out, int reset) {
  ISP_CUDA(cudaDeviceSynchronize());
  ISP_CUDA(cudaMemcpyFromSymbol(host_out, gemm::g_trace, sizeof(gemm::g_trace)));
  if (reset) {
    static long long zeros[4 * 64 * 4];
    ISP_CUDA(cudaMemcpyToSymbol(gemm::g_trace, zeros, sizeof(zeros)));
  }
  return ISP_OK;
}
#endif

extern "C" int isp_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                const void* resid, int resid_bf16, long long ldr, float alpha, int act, void* D,
                                long long ldd, int out_bf16, long long M, int N, int K, isp_stream_t stream) {
  return gemm_common(A, lda, W, ldw, bias, resid, resid_bf16, ldr, alpha, act, D, ldd, out_bf16, M, N, K, nullptr, 0,
                     nullptr, 0.f, nullptr, 0, stream);
}

// Number of (sum, sum of squares) slots per row that a GEMM / conv with N output columns writes.
extern "C" int isp_gemm_stats_slots(int N, int out_bf16, int has_resid) {
  const int bn = gemm::pick_bn(N, (has_resid && !out_bf16) ? 128 : 256);
  const int cw = out_bf16 ? 64 : 32;
  return ((N + bn - 1) / bn) * ((bn + cw - 1) / cw);
}

// isp_gemm_bf16_tc with a LayerNorm fused on either side:
//  * ln_stats != NULL: D = alpha * act( LN(A) W^T + bias ) + resid, where W must hold the gamma-scaled weights
//    W[n,k] * gamma[k] (bf16), ln_g[n] = sum_k of those bf16 values, bias[n] = sum_k W[n,k] beta[k] + b[n], and
//    ln_stats[M][ln_slots][2] the partial (sum, sum of squares) of every A row over its K real columns:
//    LN(A) W^T = rstd * (A W'^T - mean * ln_g).  The normalised activations are never materialised.
//  * stats_out != NULL: additionally writes those partial sums for the rows of D (values as stored),
//    [M][stats_slots][2] with stats_slots = isp_gemm_stats_slots(N, out_bf16, resid != NULL).
extern "C" int isp_gemm_bf16_tc_ex(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                   const void* resid, int resid_bf16, long long ldr, float alpha, int act, void* D,
                                   long long ldd, int out_bf16, long long M, int N, int K, const float* ln_stats,
                                   int ln_slots, const float* ln_g, float ln_eps, float* stats_out, int stats_slots,
                                   isp_stream_t stream) {
  return gemm_common(A, lda, W, ldw, bias, resid, resid_bf16, ldr, alpha, act, D, ldd, out_bf16, M, N, K, ln_stats,
                     ln_slots, ln_g, ln_eps, stats_out, stats_slots, stream);
}

// Batched GEMM: for every (batch b, head h)   D[b,h] (M x N) = alpha * A[b,h] (M x K) . W[b,h]^T (N x K),
// bf16 operands with unit stride along K, D bf16 | f32 with unit stride along N; every other stride (row, head,
// batch; in elements) is explicit, so heads can be column slices of a packed [tokens, 3C] qkv matrix.  Rows /
// columns beyond M / N / K are zero-filled on load and clipped on store by TMA (no bleed between problems).
// Used by the attention backward of the frozen ViT (dinov2/layers/attention.py:54-71 under autograd).
static int gemm_batched_common(const void* A, long long a_sm, long long a_sh, long long a_sb, const void* W, long long w_sn,
                               long long w_sh, long long w_sb, void* D, long long d_sm, long long d_sh, long long d_sb,
                               int out_bf16, int M, int N, int K, int H, int B, float alpha, int tn, isp_stream_t stream) {
  ISP_REQUIRE(A && W && D, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc_batched: null pointer");
  ISP_REQUIRE(M > 0 && N > 0 && K > 0 && H > 0 && B > 0, ISP_ERR_BAD_SHAPE, "gemm_bf16_tc_batched: bad shape");
  const int esz = out_bf16 ? 2 : 4;
  ISP_REQUIRE(a_sm % 8 == 0 && a_sh % 8 == 0 && a_sb % 8 == 0 && w_sn % 8 == 0 && w_sh % 8 == 0 && w_sb % 8 == 0,
              ISP_ERR_MISALIGNED, "gemm_bf16_tc_batched: A / W strides must be multiples of 8 elements");
  ISP_REQUIRE((d_sm * esz) % 16 == 0 && (d_sh * esz) % 16 == 0 && (d_sb * esz) % 16 == 0, ISP_ERR_MISALIGNED,
              "gemm_bf16_tc_batched: D strides must be multiples of 16 bytes");
  ISP_REQUIRE(aligned16(A) && aligned16(W) && aligned16(D), ISP_ERR_MISALIGNED, "gemm_bf16_tc_batched: 16-byte alignment");
  gemm::Params p = {};
  p.M = M; p.N = N;
  p.K = (K + gemm::BK - 1) / gemm::BK * gemm::BK;
  p.BN = gemm::pick_bn(N, 256);
  if (tn & 2) {  // column boxes of W are 64 wide
    p.BN = (N + 63) / 64 * 64;
    if (p.BN > 256) p.BN = 256;
  }
  p.tn = tn;
  p.last_steps = (K - (p.K - gemm::BK) + 15) / 16;
  p.TW = 0;
  p.batched = H; p.nbatch = B;
  ISP_REQUIRE((long long)((M + gemm::BM - 1) / gemm::BM) * H * B < (1ll << 24), ISP_ERR_UNSUPPORTED, "gemm_bf16_tc_batched: too many row tiles");
  p.tiles_m = ((M + gemm::BM - 1) / gemm::BM) * H * B;
  p.tiles_n = (N + p.BN - 1) / p.BN;
  p.pair = 0;
  p.bias = nullptr; p.alpha = alpha;
  CUtensorMap tmA, tmB, tmD, tmDt;
  const uint32_t box_red[4] = {64, gemm::BK, 1, 1};  // reduction-major operand: 64 columns x 64 reduction rows
  if (tn & 1) {  // a_sm is the stride of a REDUCTION row; M is the contiguous dimension
    const uint64_t dims[4] = {(uint64_t)M, (uint64_t)K, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {2, (uint64_t)a_sm * 2, (uint64_t)a_sh * 2, (uint64_t)a_sb * 2};
    if (int e = make_tmap_bf16(&tmA, A, 4, dims, str, box_red, "gemm_bf16_tc_batched(A, reduction-major)")) return e;
  } else {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {2, (uint64_t)a_sm * 2, (uint64_t)a_sh * 2, (uint64_t)a_sb * 2};
    const uint32_t box[4] = {gemm::BK, gemm::BM, 1, 1};
    if (int e = make_tmap_bf16(&tmA, A, 4, dims, str, box, "gemm_bf16_tc_batched(A)")) return e;
  }
  if (tn & 2) {
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)K, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {2, (uint64_t)w_sn * 2, (uint64_t)w_sh * 2, (uint64_t)w_sb * 2};
    if (int e = make_tmap_bf16(&tmB, W, 4, dims, str, box_red, "gemm_bf16_tc_batched(W, reduction-major)")) return e;
  } else {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {2, (uint64_t)w_sn * 2, (uint64_t)w_sh * 2, (uint64_t)w_sb * 2};
    const uint32_t box[4] = {gemm::BK, (uint32_t)p.BN, 1, 1};
    if (int e = make_tmap_bf16(&tmB, W, 4, dims, str, box, "gemm_bf16_tc_batched(W)")) return e;
  }
  const uint32_t cw = out_bf16 ? 64 : 32;
  const uint32_t tailw = (uint32_t)p.BN % cw;
  {
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)H, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)esz, (uint64_t)d_sm * esz, (uint64_t)d_sh * esz, (uint64_t)d_sb * esz};
    const uint32_t box[4] = {cw, 32, 1, 1}, boxt[4] = {tailw ? tailw : cw, 32, 1, 1};
    if (int e = make_tmap(&tmD, esz, D, 4, dims, str, box, "gemm_bf16_tc_batched(D)", true)) return e;
    if (int e = make_tmap(&tmDt, esz, D, 4, dims, str, boxt, "gemm_bf16_tc_batched(D tail)", false)) return e;
  }
  return gemm::launch(tmA, tmB, tmD, tmDt, tmD, tmDt, tmB, p, out_bf16, 0, false, as_stream(stream));
}

extern "C" int isp_gemm_bf16_tc_batched(const void* A, long long a_sm, long long a_sh, long long a_sb, const void* W,
                                        long long w_sn, long long w_sh, long long w_sb, void* D, long long d_sm,
                                        long long d_sh, long long d_sb, int out_bf16, int M, int N, int K, int H, int B,
                                        float alpha, isp_stream_t stream) {
  return gemm_batched_common(A, a_sm, a_sh, a_sb, W, w_sn, w_sh, w_sb, D, d_sm, d_sh, d_sb, out_bf16, M, N, K, H, B, alpha, 0,
                             stream);
}

// Same with the operands stored reduction-major: A[b,h] is [K][M] and W[b,h] is [K][N] (a_sk / w_sk = stride of a
// reduction row, M / N contiguous):  D[b,h] (M x N) = alpha * A[b,h]^T W[b,h].  This is the shape of the attention
// backward's dK = dS^T Q and dV = P^T dO (reduction over the queries) on the tensors as they are stored -- no
// transposed copies of the score-sized matrices.
extern "C" int isp_gemm_bf16_tc_batched_tn(const void* A, long long a_sk, long long a_sh, long long a_sb, const void* W,
                                           long long w_sk, long long w_sh, long long w_sb, void* D, long long d_sm,
                                           long long d_sh, long long d_sb, int out_bf16, int M, int N, int K, int H, int B,
                                           float alpha, isp_stream_t stream) {
  return gemm_batched_common(A, a_sk, a_sh, a_sb, W, w_sk, w_sh, w_sb, D, d_sm, d_sh, d_sb, out_bf16, M, N, K, H, B, alpha, 3,
                             stream);
}

// Mixed form: A row-major [M][K] as in isp_gemm_bf16_tc_batched, W stored reduction-major [K][N]:
// D[b,h] = alpha * A[b,h] W[b,h]  (dQ = dS K with K as stored, keys x head_dim).
extern "C" int isp_gemm_bf16_tc_batched_nn(const void* A, long long a_sm, long long a_sh, long long a_sb, const void* W,
                                           long long w_sk, long long w_sh, long long w_sb, void* D, long long d_sm,
                                           long long d_sh, long long d_sb, int out_bf16, int M, int N, int K, int H, int B,
                                           float alpha, isp_stream_t stream) {
  return gemm_batched_common(A, a_sm, a_sh, a_sb, W, w_sk, w_sh, w_sb, D, d_sm, d_sh, d_sb, out_bf16, M, N, K, H, B, alpha, 2,
                             stream);
}

static int conv3x3_common(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16, int Nimg,
                          int H, int Wd, int Cin, int ldx, int Cout, int ldy, const void* mask, int ldm,
                          float* stats_out, int stats_slots, isp_stream_t stream) {
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
  if (mask) {
    ISP_REQUIRE(act == 0, ISP_ERR_UNSUPPORTED, "conv3x3_dgrad_bf16_tc: the ReLU mask replaces the activation");
    ISP_REQUIRE(ldm >= Cout && (ldm * esz) % 16 == 0 && aligned16(mask), ISP_ERR_MISALIGNED,
                "conv3x3_dgrad_bf16_tc: mask pixels must be 16-byte aligned (ldm=%d)", ldm);
    act = 5;
  }
  gemm::Params p = {};
  p.M = (long long)Nimg * H * Wd; p.N = Cout; p.K = 9 * Cin_pad;
  // with a mask operand every epilogue warp should own at most two chunks per tile (see isp_gemm_bf16_tc)
  p.BN = gemm::pick_bn(Cout, (mask && !out_bf16) ? 128 : 256);
  p.last_steps = (Cin - (Cin_pad - 64) + 15) / 16;
  // tile = TH x TW output pixels; prefer wide rows (fewer halo re-reads), fall back to 16x8
  int TW = 16;
  for (int cand : {128, 64, 32, 16, 8}) {
    if (Wd % cand == 0 && H % (128 / cand) == 0) { TW = cand; break; }
  }
  p.TW = TW; p.TH = 128 / TW; p.H = H; p.W = Wd; p.cin_chunks = Cin_pad / 64;
  p.tiles_w = (Wd + p.TW - 1) / p.TW; p.tiles_h = (H + p.TH - 1) / p.TH;
  ISP_REQUIRE((long long)Nimg * p.tiles_w * p.tiles_h < (1ll << 24), ISP_ERR_UNSUPPORTED, "conv3x3_bf16_tc: too many pixel tiles");
  p.tiles_m = Nimg * p.tiles_w * p.tiles_h;
  p.tiles_n = (Cout + p.BN - 1) / p.BN;
  p.pair = 1;  // K loop of >= 9 k-blocks: the un-overlapped epilogue is small against the saved operand traffic
  p.bias = bias; p.alpha = 1.f;
  if (stats_out) {
    const int need = (int)p.tiles_n * ((p.BN + (out_bf16 ? 64 : 32) - 1) / (out_bf16 ? 64 : 32));
    ISP_REQUIRE(stats_slots == need, ISP_ERR_BAD_SHAPE, "conv3x3_bf16_tc_ex: stats_slots must be %d for Cout=%d", need, Cout);
    p.stats_out = stats_out; p.stats_slots = stats_slots;
  }
  CUtensorMap tmA, tmB, tmD, tmDt, tmBh;
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
    const uint32_t boxh[2] = {gemm::BK, (uint32_t)p.BN / 2};
    if (int e = make_tmap_bf16(&tmBh, Wp, 2, dims, str, boxh, "conv3x3_bf16_tc(W half)")) return e;
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
  CUtensorMap tmR = tmD, tmRt = tmDt;
  if (mask) {
    const uint32_t bw = TW < 32 ? TW : 32;
    const uint64_t dims[4] = {(uint64_t)ldm, (uint64_t)Wd, (uint64_t)H, (uint64_t)Nimg};
    const uint64_t str[4] = {(uint64_t)esz, (uint64_t)ldm * esz, (uint64_t)Wd * ldm * esz, (uint64_t)H * Wd * ldm * esz};
    const uint32_t cw = out_bf16 ? 64u : 32u, tailw = (uint32_t)p.BN % cw;
    const uint32_t box[4] = {cw, bw, 32 / bw, 1}, boxt[4] = {tailw ? tailw : cw, bw, 32 / bw, 1};
    if (int e = make_tmap(&tmR, esz, mask, 4, dims, str, box, "conv3x3_dgrad_bf16_tc(mask)", true)) return e;
    if (int e = make_tmap(&tmRt, esz, mask, 4, dims, str, boxt, "conv3x3_dgrad_bf16_tc(mask tail)", false)) return e;
  }
  return gemm::launch(tmA, tmB, tmD, tmDt, tmR, tmRt, tmBh, p, out_bf16, act, mask != nullptr, as_stream(stream));
}

extern "C" int isp_conv3x3_bf16_tc(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                                   int Nimg, int H, int Wd, int Cin, int ldx, int Cout, int ldy, isp_stream_t stream) {
  return conv3x3_common(X, Wp, bias, act, Y, out_bf16, Nimg, H, Wd, Cin, ldx, Cout, ldy, nullptr, 0, nullptr, 0, stream);
}

// isp_conv3x3_bf16_tc that also writes the row statistics of Y (see isp_gemm_bf16_tc_ex); rows = pixels in
// NHWC order.  stats_slots = isp_gemm_stats_slots(Cout, out_bf16, 0).
extern "C" int isp_conv3x3_bf16_tc_ex(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                                      int Nimg, int H, int Wd, int Cin, int ldx, int Cout, int ldy, float* stats_out,
                                      int stats_slots, isp_stream_t stream) {
  return conv3x3_common(X, Wp, bias, act, Y, out_bf16, Nimg, H, Wd, Cin, ldx, Cout, ldy, nullptr, 0, stats_out,
                        stats_slots, stream);
}

// Data gradient of the same convolution: dX = conv3x3(dY, W') with W'[ci][tap][co] = W[co][ci][8 - tap]
// (the caller packs the flipped, transposed weights like forward ones), optionally masked by the ReLU
// of the layer that produced X:  dX = relu_mask > 0 ? conv : 0.  Same kernel, ReLU-mask epilogue.
extern "C" int isp_conv3x3_dgrad_bf16_tc(const void* dY, const void* Wp_flipped, const void* relu_mask, int ldm,
                                         void* dX, int out_bf16, int Nimg, int H, int Wd, int Cout, int ldy, int Cin,
                                         int ldx, isp_stream_t stream) {
  return conv3x3_common(dY, Wp_flipped, nullptr, 0, dX, out_bf16, Nimg, H, Wd, Cout, ldy, Cin, ldx, relu_mask, ldm,
                        nullptr, 0, stream);
}
