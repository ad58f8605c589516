// Weight gradient of the 3x3 convolution (stride 1, padding 1) on tcgen05 (sm_100a):
//
//   dW[co][tap][ci] += sum over pixels p of  dY[p][co] * X[p + tap][ci]        bf16 operands, fp32 accumulate
//
// This is a GEMM whose reduction dimension is the PIXEL axis, i.e. both operands are "MN-major" in
// shared memory (channels contiguous, pixels along K).  TMA brings NHWC boxes of 64 pixels x 64
// channels (128-byte rows, 128B swizzle) -- for X shifted by the tap, out-of-bounds pixels zero-filled,
// which IS the padding -- and the UMMA descriptors walk them MN-major (LBO = one 64-channel box,
// SBO = 8 pixel rows).  Nothing is transposed or im2col'ed in memory.
//
// Work item = (tap, 128-wide Cout tile, BN-wide Cin tile, K split); K runs over (image, row, 64-pixel
// chunk).  One persistent CTA per SM: warp 0 TMA, warp 1 MMA (accumulator in TMEM for the whole K range
// of the item), warps 2-5 epilogue: tcgen05.ld -> red.global.add.v4.f32 into the fp32 gradient (the K
// splits and, in training, micro-batches accumulate there).
// Replaces cuDNN's wgrad under ConvSegHead.convs (core/model/heads/conv_heads.py:58-66) in the
// trainer's backward (core/training/trainer.py:213-221).
#include "tc_common.cuh"

namespace isp {
namespace wgrad {

constexpr int BM = 128;           // Cout tile
constexpr int KPIX = 64;          // pixels per k-block
constexpr int kThreads = 192;     // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kMaxStages = 6;
constexpr int kBudget = 200 * 1024;

struct Params {
  int Nimg, H, W, Cout, Cin;
  int BN;                 // Cin tile (multiple of 64, <= 256; 384 = two 192-wide MMAs per step)
  int tiles_m, tiles_n, wchunks, splits, stages;
  long long kblocks;      // Nimg * H * wchunks
  float* dW;              // [Cout][9][Cin] fp32, accumulated into
};

// MN-major operand, 128-byte swizzle: 64 channels (128 B) per pixel row, 8-row groups 1024 B apart (SBO),
// the next 64 channels one box further (LBO).  cute::UMMA canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16_f32_mn(int M, int N) {  // both operands MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box_bytes = KPIX * 128;                 // 64 pixels x 64 channels bf16
  const uint32_t a_bytes = 2 * box_bytes, b_bytes = (uint32_t)(p.BN / 64) * box_bytes, stage_bytes = a_bytes + b_bytes;
  const long long nitems = (long long)9 * p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmDY);
    tc::prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&acc_full, 1);
    tc::mbar_init(&acc_empty, 4);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // item -> (tile_n, tile_m, tap, split), split SLOWEST: the CTAs running at the same time work on all
  // (tap, Cout tile, Cin tile) combinations of the same pixel window, so X and dY come from DRAM once and are
  // re-read from L2 (with the split fastest every CTA streamed its own window: 5.4x the DRAM reads, ncu)
  auto decode = [&](long long it, int& tap, int& mt, int& nt, long long& kb0, long long& kb1) {
    const long long tiles = (long long)9 * p.tiles_m * p.tiles_n;
    const int split = (int)(it / tiles);
    long long r = it % tiles;
    nt = (int)(r % p.tiles_n); r /= p.tiles_n;
    mt = (int)(r % p.tiles_m); r /= p.tiles_m;
    tap = (int)r;
    const long long per = (p.kblocks + p.splits - 1) / p.splits;
    kb0 = split * per;
    kb1 = kb0 + per < p.kblocks ? kb0 + per : p.kblocks;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0;
      for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        int tap, mt, nt;
        long long kb0, kb1;
        decode(it, tap, mt, nt, kb0, kb1);
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        for (long long kb = kb0; kb < kb1; ++kb, ++n) {
          const int s = n % p.stages;
          tc::mbar_wait(&empty_bar[s], ((n / p.stages) & 1) ^ 1);
          const int wc = (int)(kb % p.wchunks);
          const long long rr = kb / p.wchunks;
          const int h = (int)(rr % p.H), img = (int)(rr / p.H);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          tc::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          for (int i = 0; i < 2; ++i)
            tc::tma_load_4d(sa + i * box_bytes, &tmDY, &full_bar[s], mt * BM + i * 64, wc * KPIX, h, img);
          for (int i = 0; i < p.BN / 64; ++i)
            tc::tma_load_4d(sb + i * box_bytes, &tmX, &full_bar[s], nt * p.BN + i * 64, wc * KPIX + dx, h + dy, img);
        }
      }
    }
  } else if (warp == 1) {
    // Cin tiles wider than one MMA (N <= 256) are issued as two half-width MMAs that share the dY operand: a 384-wide tile
    // moves 21 operand bytes per kMAC through L2 -> SM instead of the 25 of two 192-wide items
    const int halves = p.BN > 256 ? 2 : 1;
    const int bn_mma = p.BN / halves;
    const uint32_t idesc = idesc_bf16_f32_mn(BM, bn_mma);
    const bool leader = tc::elect_one();
    uint32_t n = 0, item_n = 0;
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x, ++item_n) {
      int tap, mt, nt;
      long long kb0, kb1;
      decode(it, tap, mt, nt, kb0, kb1);
      tc::mbar_wait(&acc_empty, (item_n & 1) ^ 1);
      tc::tc_fence_after();
      for (long long kb = kb0; kb < kb1; ++kb, ++n) {
        const int s = n % p.stages;
        tc::mbar_wait(&full_bar[s], (n / p.stages) & 1);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
        for (int k = 0; k < KPIX / 16; ++k) {  // 16 pixel rows = 2048 B per MMA
          const uint64_t da = smem_desc_mn_sw128(sa + k * 2048, box_bytes);
          const uint64_t db = smem_desc_mn_sw128(sa + a_bytes + k * 2048, box_bytes);
          if (leader) tc::umma_bf16(tmem, da, db, idesc, (kb > kb0 || k) ? 1u : 0u);
          if (halves == 2) {
            const uint64_t db2 = smem_desc_mn_sw128(sa + a_bytes + (bn_mma / 64) * box_bytes + k * 2048, box_bytes);
            if (leader) tc::umma_bf16(tmem + bn_mma, da, db2, idesc, (kb > kb0 || k) ? 1u : 0u);
          }
        }
        if (leader) tc::umma_commit(&empty_bar[s]);
      }
      if (leader) tc::umma_commit(&acc_full);
    }
  } else {
    const int q = warp & 3;
    const uint32_t t_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t item_n = 0;
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x, ++item_n) {
      int tap, mt, nt;
      long long kb0, kb1;
      decode(it, tap, mt, nt, kb0, kb1);
      tc::mbar_wait(&acc_full, item_n & 1);
      tc::tc_fence_after();
      const int co = mt * BM + q * 32 + lane;
      float* drow = p.dW + ((long long)co * 9 + tap) * p.Cin + nt * p.BN;
      const bool live = co < p.Cout && kb1 > kb0;
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(t_addr + c, v);
        tc::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (nt * p.BN + c + e < p.Cin)
              red_add_v4(drow + c + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                         __uint_as_float(v[e + 3]));
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace wgrad
}  // namespace isp

using namespace isp;

extern "C" int isp_conv3x3_wgrad_bf16_tc(const void* X, int ldx, const void* dY, int ldy, float* dW, int Nimg, int H,
                                         int Wd, int Cin, int Cout, isp_stream_t stream) {
  ISP_REQUIRE(X && dY && dW, ISP_ERR_BAD_SHAPE, "conv3x3_wgrad_bf16_tc: null pointer");
  ISP_REQUIRE(Nimg > 0 && H > 0 && Wd > 0 && Cin > 0 && Cout > 0, ISP_ERR_BAD_SHAPE, "conv3x3_wgrad_bf16_tc: bad shape");
  ISP_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, ISP_ERR_UNSUPPORTED,
              "conv3x3_wgrad_bf16_tc: Cin and Cout must be multiples of 64 (got %d, %d)", Cin, Cout);
  ISP_REQUIRE(ldx >= Cin && ldy >= Cout && ldx % 8 == 0 && ldy % 8 == 0, ISP_ERR_MISALIGNED,
              "conv3x3_wgrad_bf16_tc: channel strides must be multiples of 8");
  ISP_REQUIRE(aligned16(X) && aligned16(dY) && aligned16(dW), ISP_ERR_MISALIGNED, "conv3x3_wgrad_bf16_tc: 16-byte alignment");
  wgrad::Params p = {};
  p.Nimg = Nimg; p.H = H; p.W = Wd; p.Cout = Cout; p.Cin = Cin;
  p.BN = Cin % 192 == 0 ? 192 : (Cin % 128 == 0 ? 128 : 64);
  if (Cin % 256 == 0) p.BN = 256;
  if (Cin % 384 == 0) p.BN = 384;  // two 192-wide MMAs per step on one dY tile (TMEM: 384 of 512 columns)
  p.tiles_m = (Cout + wgrad::BM - 1) / wgrad::BM;
  p.tiles_n = Cin / p.BN;
  p.wchunks = (Wd + wgrad::KPIX - 1) / wgrad::KPIX;
  p.kblocks = (long long)Nimg * H * p.wchunks;
  p.dW = dW;
  int num_sms = 0;
  if (int e = device_sm_count(&num_sms)) return e;
  if (int e = ensure_dynamic_smem((const void*)wgrad::wgrad_tc_kernel, wgrad::kBudget)) return e;
  const int stage_bytes = (2 + p.BN / 64) * wgrad::KPIX * 128;
  p.stages = wgrad::kBudget / stage_bytes;
  if (p.stages > wgrad::kMaxStages) p.stages = wgrad::kMaxStages;
  // K splits: a pixel window of ~32 MB of X + dY per split (a few windows in flight fit the 126 MB L2) and enough
  // items for ~3 waves of the grid, but at least 64 k-blocks each
  const long long tiles = (long long)9 * p.tiles_m * p.tiles_n;
  long long splits = (3LL * num_sms + tiles - 1) / tiles;
  const long long win_kblocks = (32LL << 20) / ((long long)(Cin + Cout) * 2 * wgrad::KPIX);
  if (win_kblocks > 0 && (p.kblocks + win_kblocks - 1) / win_kblocks > splits) splits = (p.kblocks + win_kblocks - 1) / win_kblocks;
  const long long max_splits = p.kblocks / 64 > 0 ? p.kblocks / 64 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  CUtensorMap tmDY, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wd, (uint64_t)H, (uint64_t)Nimg};
    const uint64_t str[4] = {2, (uint64_t)ldy * 2, (uint64_t)Wd * ldy * 2, (uint64_t)H * Wd * ldy * 2};
    const uint32_t box[4] = {64, wgrad::KPIX, 1, 1};
    if (int e = make_tmap_bf16(&tmDY, dY, 4, dims, str, box, "conv3x3_wgrad_bf16_tc(dY)")) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Wd, (uint64_t)H, (uint64_t)Nimg};
    const uint64_t str[4] = {2, (uint64_t)ldx * 2, (uint64_t)Wd * ldx * 2, (uint64_t)H * Wd * ldx * 2};
    const uint32_t box[4] = {64, wgrad::KPIX, 1, 1};
    if (int e = make_tmap_bf16(&tmX, X, 4, dims, str, box, "conv3x3_wgrad_bf16_tc(X)")) return e;
  }
  const long long nitems = tiles * p.splits;
  const int grid = (int)(nitems < num_sms ? nitems : num_sms);
  wgrad::wgrad_tc_kernel<<<grid, wgrad::kThreads, wgrad::kBudget, as_stream(stream)>>>(tmDY, tmX, p);
  ISP_CHECK_LAUNCH("wgrad_tc_kernel");
  return ISP_OK;
}
