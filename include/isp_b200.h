/* libisp_b200 -- C ABI of the B200-native (sm_100a) hot path of iSegProbe.
 *
 * Drop-in boundary: the reference is pure Python, so the
 * "FFI" a maintainer binds is ctypes from the reference's plugin classes
 * (UPSAMPLER_REGISTRY / HEAD_REGISTRY entries, DistMaps).  Every entry point
 * below names the reference function it replaces.  Conventions:
 *   - plain device pointers + sizes, no torch types; the caller owns all memory;
 *   - nothing allocates, synchronises or keeps global state (apart from a
 *     process-wide TMA-encode function pointer resolved once);
 *   - work is enqueued on `stream` (a cudaStream_t) and is stream-ordered;
 *   - returns ISP_OK or a negative code; isp_last_error() gives the text;
 *   - there is NO CPU fallback: without a sm_100 device the launch fails.
 * Layouts: "NHWC" = channels innermost.  Unless said otherwise tensors are dense.
 */
#ifndef ISP_B200_H
#define ISP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* isp_stream_t; /* cudaStream_t */

#define ISP_OK 0
#define ISP_ERR_BAD_SHAPE (-1)
#define ISP_ERR_MISALIGNED (-2)
#define ISP_ERR_CUDA (-3)
#define ISP_ERR_UNSUPPORTED (-4)
#define ISP_ERR_WORKSPACE (-5)

#define ISP_ABI_VERSION 1

int isp_version(void);
const char* isp_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long isp_launch_count(void);

/* ---- click-map encoding -------------------------------------------------
 * Replaces DistMaps.get_coord_features, torch path
 * (core/model/ops.py:35-77): points [B,2P,3] float32 (row, col, order) ->
 * out [B,2,H,W] float32.  use_disks: 1.0 where min squared distance <=
 * (norm_radius*spatial_scale)^2 else 0.0; otherwise tanh(2*sqrt(d2)) with
 * coordinates divided by norm_radius*spatial_scale.  Bit-exact (disks). */
int isp_distmaps_fwd(const float* points, float* out, int B, int P, int H, int W,
                     float norm_radius, float spatial_scale, int use_disks, isp_stream_t stream);

/* Replaces the Cython `get_dist_maps` (core/utils/cython/_get_dist_maps.pyx:18-64)
 * as dispatched by DistMaps cpu_mode (ops.py:21-34): click coordinates are
 * ROUNDED, validity tests the row only, output is the squared distance
 * divided by norm_delimeter^2 (1e6 where no click), [B,2,H,W]. */
int isp_distmaps_rounded_sqdist_fwd(const float* points, float* out, int B, int P, int H, int W,
                                    float norm_delimeter, isp_stream_t stream);

/* Fuses prepare_input + get_coord_features (core/model/iseg_base_model.py:91-110,
 * BatchImageNormalize ops.py:96-105): image [B,Cin,H,W] (Cin = 3, or 4 with the
 * previous mask as channel 3) -> norm_image [B,3,H,W] and coord [B,Cc,H,W] with
 * Cc = 2 (+1 leading prev-mask channel when Cin == 4). */
int isp_prepare_input_fwd(const float* image, const float* points, float* norm_image, float* coord,
                          int B, int Cin, int P, int H, int W, const float* mean3, const float* std3,
                          float norm_radius, float spatial_scale, int use_disks, isp_stream_t stream);

/* ---- layout / dtype movers ---------------------------------------------- */
/* [B,C,H,W] f32 (arbitrary element strides) -> dense NHWC f32 or bf16 */
int isp_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int H, int W,
                         long long sb, long long sc, long long sh, long long sw, isp_stream_t stream);
int isp_nchw_to_nhwc_bf16(const float* in, void* out_bf16, int B, int C, int H, int W, int Cpad,
                          long long sb, long long sc, long long sh, long long sw, isp_stream_t stream);
/* bilinear, align_corners=True, NHWC f32 -> NHWC f32 (iseg_probe_model.py:120-129,
 * basic_upsamplers.py:26-33).  out_bf16 != 0 writes bf16 with Cpad channels. */
int isp_bilinear_ac_nhwc(const float* in, void* out, int B, int C, int Hin, int Win, int Hout, int Wout,
                         int out_bf16, int Cpad, isp_stream_t stream);
/* Gradient of isp_bilinear_ac_nhwc w.r.t. its input (adjoint of the align_corners=True resize,
 * core/model/iseg_probe_model.py:120-129 under autograd): gin [B,Hin,Win,C] from gout [B,Hout,Wout,C], fp32 NHWC. */
int isp_bilinear_ac_nhwc_bwd(const float* gout, float* gin, int B, int C, int Hin, int Win, int Hout, int Wout,
                             isp_stream_t stream);
/* Same resize writing BOTH an f32 result and its bf16 copy (the operand of the tensor-core 1x1
 * conv that follows) in one pass over the input.  C % 4 == 0. */
int isp_bilinear_ac_nhwc_dual(const float* in, float* out_f32, void* out_bf16, int B, int C, int Hin, int Win,
                              int Hout, int Wout, isp_stream_t stream);
/* out = resize(in) + bias[c] (fp32 NHWC -> fp32 or bf16 NHWC, align_corners=True, bias NULL or C floats): the resize the reference
 * applies after the upsampler (core/model/iseg_probe_model.py:120-129) with the bias of JBUStack's final
 * `fixup_proj(x) * 0.1 + x` (FeatUp, via core/model/upsamplers/JBUFeatUp.py:30-32) folded in; the conv's matrix part
 * is applied at source resolution (isegprobe_b200/upsamplers.py `_mix_channels`). */
int isp_bilinear_ac_nhwc_bias(const float* in, void* out, int out_bf16, const float* bias, int B, int C, int Hin, int Win,
                              int Hout, int Wout, isp_stream_t stream);

/* ---- FeatUp JBU stack (external to the reference tree: JBUFeatUp.py:30-32) --- */
/* F.adaptive_avg_pool2d(guidance, (OH,OW)): NCHW f32 [B,3,H,W] -> NHWC4 f32 [B,OH,OW,4] (4th = 0) */
int isp_jbu_pool_guidance(const float* guidance, float* out, int B, int H, int W, int OH, int OW,
                          long long sb, long long sc, long long sh, long long sw, isp_stream_t stream);
/* range_proj: Conv1x1(3->32) -> GELU -> Conv1x1(32->32).  g NHWC4 -> proj [B,H,W,32].
 * w0 [32,3], b0 [32], w1 [32,32], b1 [32] (row-major, out-channel first). */
int isp_jbu_range_proj(const float* g, float* proj, long long npix, const float* w0, const float* b0,
                       const float* w1, const float* b1, isp_stream_t stream);
/* combined kernel of JBULearnedRange.forward: softmax_49(temp*<proj nbr, proj>) * spatial,
 * renormalised, + 0.1*fixup([k,g]).  filters out [B,H,W,49].
 * fw0 [49,52], fb0 [49], fw1 [49,49], fb1 [49]; temp = clamp(exp(range_temp),1e-4,1e4). */
/* out_ld selects the filter layout: 49 = dense [B,H,W,49]; 56 = row-padded [B,H,W,7,8] (8th tap of
 * each filter row is 0), which isp_adaptive_conv_fwd fetches with one TMA box per tile. */
int isp_jbu_filters(const float* proj, const float* g, float* filters, int B, int H, int W,
                    float temp, float sigma_spatial, const float* fw0, const float* fb0,
                    const float* fw1, const float* fb1, int out_ld, isp_stream_t stream);
/* Same result with the 52 -> 49 -> 49 fix-up MLP on the fp32 pipe for both layouts (isp_jbu_filters runs it on the tensor
 * cores, tcgen05 kind::tf32 with hi / lo split activations, for the padded layout): cross-check and A/B timing. */
int isp_jbu_filters_simt(const float* proj, const float* guidance4, float* filters, int B, int H, int W, float temp,
                         float sigma_spatial, const float* fw0, const float* fb0, const float* fw1, const float* fb1,
                         int out_ld, isp_stream_t stream);
/* bicubic x2 (align_corners=False, A=-0.75) followed by reflect pad 3:
 * src NHWC [B,h,w,C] -> out NHWC [B,2h+6,2w+6,C] */
int isp_jbu_bicubic2x_reflectpad(const float* src, float* out, int B, int h, int w, int C,
                                 isp_stream_t stream);
/* Its adjoint (gradient w.r.t. src, for the activation backward of the frozen JBU stack):
 * gsrc [B,h,w,C] from gpad [B,2h+6,2w+6,C]. */
int isp_jbu_bicubic2x_reflectpad_bwd(const float* gpad, float* gsrc, int B, int h, int w, int C, isp_stream_t stream);
/* AdaptiveConv.forward: out[b,y,x,c] = sum_{i,j<7} in[b,y+i,x+j,c] * filt[b,y,x,i*7+j].
 * NHWC fast path: in [B,H+6,W+6,C], out [B,H,W,C], C % 64 == 0.  filt_ld: 49 = dense filters
 * [B,H,W,49] (FeatUp's layout), 56 = row-padded [B,H,W,7,8] (see isp_jbu_filters). */
int isp_adaptive_conv_fwd(const float* in_padded, const float* filters, float* out,
                          int B, int H, int W, int C, int filt_ld, isp_stream_t stream);
/* AdaptiveConv.backward wrt the padded input (NHWC): gi [B,H+6,W+6,C] */
int isp_adaptive_conv_grad_input(const float* grad_out, const float* filters, float* grad_in,
                                 int B, int H, int W, int C, isp_stream_t stream);

/* ---- SIMT fp32 GEMM (bring-up / cross-check of the tensor-core path) -----
 * C[M,N] = A[M,K] * W[N,K]^T + bias[N]; optional residual: C = alpha*C + resid */
int isp_gemm_f32_simt(const float* A, const float* W, const float* bias, const float* resid, float alpha,
                      float* C, long long M, int N, int K, isp_stream_t stream);

/* ---- tcgen05 tensor-core GEMM (TMA -> smem ring -> tcgen05.mma -> TMEM -> epilogue) ----
 * D[M,N] = alpha * act(A[M,K] * W[N,K]^T + bias[N]) + resid[M,N]
 * A, W bf16 row-major (K contiguous; lda, ldw in elements, multiples of 8); D bf16 or f32
 * with row stride N <= ldd <= round_up(N, 16): the padding columns N..ldd-1 are ZERO-FILLED, so D must own them (it cannot
 * be a column slice of a wider matrix);
 * resid bf16 or f32 with row stride ldr, or NULL; act: 0 none, 1 ReLU, 2 GELU(erf), 3 QuickGELU,
 * 4 GELU in tanh form (one MUFU; differs from the erf form by < 5e-4, used by the bf16 pipelines).
 * Replaces the cuBLAS fp32 GEMMs under nn.Linear / Conv2d(1x1) in LoftUp
 * (loftup/layers.py:161-202, loftup/loftup.py:67-70) and the ViT blocks
 * (featurizers/dinov2/layers/attention.py:54-71, mlp.py:34-40). */
int isp_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     const void* resid, int resid_bf16, long long ldr, float alpha, int act, void* D,
                     long long ldd, int out_bf16, long long M, int N, int K, isp_stream_t stream);

/* Implicit-GEMM 3x3 convolution, stride 1, padding 1, on the same core:
 * X NHWC bf16 [Nimg,H,W,ldx] (Cin real channels), Wp bf16 [Cout][9][Cin_pad] with
 * Cin_pad = ceil(Cin/64)*64 (tap-major, zero padded), Y NHWC [Nimg,H,W,ldy] bf16|f32,
 * Y = act(conv(X) + bias).  Padding comes from TMA out-of-bounds zero fill.
 * Replaces cuDNN under LoftUp.first_conv (loftup/loftup.py:55-65, BatchNorm folded by the
 * caller) and ConvSegHead.convs (heads/conv_heads.py:58-66). */
int isp_conv3x3_bf16_tc(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                        int Nimg, int H, int W, int Cin, int ldx, int Cout, int ldy, isp_stream_t stream);

/* The same two kernels with a LayerNorm fused on either side, so that LN(x) between two GEMMs of the
 * LoftUp transformer (loftup/layers.py:186-202: norm_q -> in_proj, :161-174: LayerNorm -> Linear) is never
 * materialised:
 *  - ln_stats != NULL: D = alpha * act(LN(A) W^T + bias) + resid.  W must hold W[n,k]*gamma[k] (bf16),
 *    ln_g[n] = sum_k of those bf16 values, bias[n] = sum_k W[n,k]*beta[k] + b[n]; ln_stats is
 *    [M][ln_slots][2] partial (sum, sum of squares) of each A row over its K real columns.
 *    LN(A) W^T = rstd * (A W'^T - mean * ln_g), applied per row in the epilogue.
 *  - stats_out != NULL: also writes those partial sums for the rows of D / pixels of Y (of the values as
 *    stored, columns < N), [M][stats_slots][2], stats_slots = isp_gemm_stats_slots(N, out_bf16, resid != NULL);
 *    one slot per (column tile, 128-byte chunk): no atomics, bit-reproducible. */
int isp_gemm_stats_slots(int N, int out_bf16, int has_resid);
int isp_gemm_bf16_tc_ex(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                        const void* resid, int resid_bf16, long long ldr, float alpha, int act, void* D,
                        long long ldd, int out_bf16, long long M, int N, int K, const float* ln_stats, int ln_slots,
                        const float* ln_g, float ln_eps, float* stats_out, int stats_slots, isp_stream_t stream);
int isp_conv3x3_bf16_tc_ex(const void* X, const void* Wp, const float* bias, int act, void* Y, int out_bf16,
                           int Nimg, int H, int W, int Cin, int ldx, int Cout, int ldy, float* stats_out,
                           int stats_slots, isp_stream_t stream);

/* Backward of that convolution (trainer backward, core/training/trainer.py:213-221, of
 * ConvSegHead.convs, heads/conv_heads.py:58-66):
 * dgrad: dX = conv3x3(dY, W') with W' = flipped / transposed weights packed like forward ones
 *        ([Cin][9][Cout_pad], tap 8 - t); if relu_mask != NULL (the bf16|f32 activation X itself,
 *        same dtype as dX, pixel stride ldm) the result is zeroed where relu_mask <= 0.
 * wgrad: dW[co][tap][ci] += sum_pixels dY[p][co] * X[p + tap][ci]  (fp32 [Cout][9][Cin], ACCUMULATED:
 *        zero it first for a fresh gradient).  Cin, Cout multiples of 64.  Pixel-major tcgen05 GEMM
 *        with MN-major operands straight from the NHWC tensors. */
int isp_conv3x3_dgrad_bf16_tc(const void* dY, const void* Wp_flipped, const void* relu_mask, int ldm, void* dX,
                              int out_bf16, int Nimg, int H, int W, int Cout, int ldy, int Cin, int ldx,
                              isp_stream_t stream);
int isp_conv3x3_wgrad_bf16_tc(const void* X, int ldx, const void* dY, int ldy, float* dW, int Nimg, int H, int W,
                              int Cin, int Cout, isp_stream_t stream);

/* IS-head classifier backward (num_classes == 1; heads/base_head.py:8-18) fused with the ReLU mask
 * of the last 3x3 layer: dz = act > 0 ? dlogits * wc : 0 (bf16), dwc += act^T dlogits,
 * dbc += sum dlogits, dbz += column sums of dz (that layer's bias gradient).  All three are
 * ACCUMULATED.  isp_colsum_bf16: out[c] += sum_m x[m, c] (bias gradient of the other layers). */
int isp_head_classifier_bwd(const void* act_bf16, long long lda, const float* dlogits, const float* wc,
                            void* dz_bf16, long long ldz, float* dwc, float* dbc, float* dbz, long long M, int C,
                            isp_stream_t stream);
int isp_colsum_bf16(const void* x_bf16, long long ld, float* out, long long M, int C, isp_stream_t stream);

/* LoftUp FeedForward block (loftup/layers.py:161-174: x + Linear2(GELU(Linear1(LayerNorm(x))))) in one tcgen05 kernel: the
 * [128 x NH] hidden tile of a row block stays in shared memory (A operand of the second GEMM), so the hidden activations
 * never reach HBM.  LayerNorm folded as in isp_gemm_bf16_tc_ex: W1g = W1 * gamma (bf16 [NH, ldw1]), g1[n] = sum_k W1g[n,k],
 * b1 = W1 beta + bias1; ln_stats fp32 [M][ln_slots][2] = partial (sum, sum of squares) of every x row.  x bf16 [M, ldx]
 * (K1 real columns, padding zero, ldx <= 448); NH <= 384, multiple of 64; W2 bf16 [N2, ldw2], N2 <= 416; out bf16 [M, ldo]
 * with N2 <= ldo <= round_up(N2, 16) (padding zero-filled); GELU in tanh form (act 4 of isp_gemm_bf16_tc).
 * stats_out (or NULL): fp32 [M][4][2] partial (sum, sum of squares) of every stored output row (ln_slots = 4 downstream). */
int isp_ffn_fused_bf16_tc(const void* x, long long ldx, int K1, const void* W1g, long long ldw1, const float* g1,
                          const float* b1, int NH, const void* W2, long long ldw2, const float* b2, int N2, void* out,
                          long long ldo, long long M, const float* ln_stats, int ln_slots, float ln_eps, float* stats_out,
                          isp_stream_t stream);

/* Device side of the NoC evaluation loop, one sample per call, fp32.
 * isp_zoom_in_fwd: ZoomIn._transform + AddHorizontalFlip.transform (core/inference/transforms/zoom_in.py:51-104,216-240,
 *   flip.py:13-29): out [with_flip ? 2 : 1][4][S0][S1] = bilinear (align_corners=True) resize of the ROI rows rmin..rmax, columns
 *   cmin..cmax (inclusive) of image [3,Hs,Ws] || prev [Hs,Ws] (NULL = zeros), and its horizontal flip.
 * isp_unzoom_probs: the inverse chain (flip.py:31-36, base_transform.py:38-39, zoom_in.py:106-130) -- probabilities
 *   sigmoid(mean of the logits and the flipped logits) resized back to the ROI, zeros elsewhere -> prob [Hs,Ws]; pred_mask =
 *   prob > pred_thr (uint8); stats int[8] = {intersection, union with gt (label -1 ignored; core/inference/utils.py:107-120),
 *   #pixels with prob > box_thr, their first / last row and column (INT_MAX / -1 if none), 0}.  gt may be NULL.
 * isp_noc_next_click: Clicker._get_next_click (core/inference/clicker.py:58-91): exact Euclidean distance transform of the
 *   zero-padded false-negative / false-positive masks (cv2.distanceTransform(DIST_L2, 0)), clicked pixels zeroed;
 *   best[m] = (float bits of the largest distance << 32) | (0xFFFFFFFF - row-major index of its first occurrence), m = 0
 *   false negatives, 1 false positives.  work: 2*H*W ints. */
int isp_zoom_in_fwd(const float* image, const float* prev, int Hs, int Ws, int rmin, int rmax, int cmin, int cmax, float* out,
                    int S0, int S1, int with_flip, isp_stream_t stream);
int isp_unzoom_probs(const float* logits, int S0, int S1, int with_flip, int Hs, int Ws, int rmin, int rmax, int cmin,
                     int cmax, float* prob, const int* gt, float pred_thr, float box_thr, unsigned char* pred_mask, int* stats,
                     isp_stream_t stream);
int isp_noc_next_click(const int* gt, const unsigned char* pred_mask, const unsigned char* clicked, int H, int W, int* work,
                       unsigned long long* best, isp_stream_t stream);

/* BatchNorm2d in training mode for the frozen LoftUp / LiFT conv stacks (loftup/loftup.py:55-65, LiFT.py:17-24 under the
 * trainer's net.train(), core/training/trainer.py:213-214): the conv writes its raw bf16 output, then
 * isp_col_moments_bf16 writes per-channel (sum, sum of squares) partials for slabs of ISP_COL_MOMENTS_SLAB_ROWS rows --
 * partial fp32 [slabs][2][ldp], no atomics -- and isp_bn_relu_rows_bf16 applies y = max(x*scale[c] + shift[c], 0) in place
 * and (optionally) writes stats fp32 [M][2], the (sum, sum of squares) of every stored row for a fused LayerNorm with
 * ln_slots = 1 (isp_gemm_bf16_tc_ex). */
#define ISP_COL_MOMENTS_SLAB_ROWS 1024
int isp_col_moments_bf16(const void* x_bf16, long long ld, long long M, int C, float* partial, int ldp, isp_stream_t stream);
int isp_bn_relu_rows_bf16(void* x_bf16, long long ld, const float* scale, const float* shift, long long M, int C,
                          float* stats, isp_stream_t stream);

/* Activation backward through the frozen ViT: the reference trains the click embedding THROUGH the frozen
 * backbone (core/model/featurizers/DINOv2.py:518-523 injects it before the blocks; trainer backward
 * core/training/trainer.py:213-221), so d(loss)/d(additional_features) needs every block's input gradient
 * (dinov2/layers/block.py:92-117, attention.py:54-71, mlp.py:34-40).  Weight-side GEMMs of that backward are
 * isp_gemm_bf16_tc with transposed packed weights; the per-head attention products use
 *   isp_gemm_bf16_tc_batched: for every (batch b, head h)  D[b,h] (M x N) = alpha * A[b,h] (M x K) . W[b,h]^T (N x K),
 *     bf16 operands with unit stride along K, D bf16 | f32 with unit stride along N, all other strides
 *     (row, head, batch; in elements) explicit -- heads may be column slices of a packed [tokens, 3C] matrix.
 * Row / element kernels:
 *   isp_layernorm_rows_bwd: dx = LN'(x)^T (gamma * dy) + resid (x fp32 | bf16, dx fp32 + optional bf16 copy), C <= 1024;
 *   isp_gelu_bwd_bf16:      dpre = dh * act'(pre)  (quick = 0: nn.GELU erf form; 1: CLIP QuickGELU), n even;
 *   isp_softmax_rows:       P[r, :ncols] = softmax(S[r, :ncols]) (fp32 -> bf16), zeros up to ncols_pad;
 *   isp_attn_ds_rows:       dS = P * (dP - sum_j P dP) per row (dP fp32 | bf16), zeros up to ncols_pad;
 *   isp_transpose_bf16_batched: dst[z][c][r] = src[z][r][c]. */
int isp_gemm_bf16_tc_batched(const void* A, long long a_sm, long long a_sh, long long a_sb, const void* W,
                             long long w_sn, long long w_sh, long long w_sb, void* D, long long d_sm, long long d_sh,
                             long long d_sb, int out_bf16, int M, int N, int K, int H, int B, float alpha,
                             isp_stream_t stream);
/* ... and with both operands stored reduction-major (A[b,h] = [K][M], W[b,h] = [K][N]; a_sk / w_sk = stride of a
 * reduction row): D[b,h] = alpha * A[b,h]^T W[b,h] -- dK = dS^T Q and dV = P^T dO without transposed copies. */
int isp_gemm_bf16_tc_batched_tn(const void* A, long long a_sk, long long a_sh, long long a_sb, const void* W,
                                long long w_sk, long long w_sh, long long w_sb, void* D, long long d_sm, long long d_sh,
                                long long d_sb, int out_bf16, int M, int N, int K, int H, int B, float alpha,
                                isp_stream_t stream);
/* Mixed: A [M][K] row-major, W stored reduction-major [K][N]: D[b,h] = alpha * A[b,h] W[b,h] (dQ = dS K). */
int isp_gemm_bf16_tc_batched_nn(const void* A, long long a_sm, long long a_sh, long long a_sb, const void* W,
                                long long w_sk, long long w_sh, long long w_sb, void* D, long long d_sm, long long d_sh,
                                long long d_sb, int out_bf16, int M, int N, int K, int H, int B, float alpha,
                                isp_stream_t stream);
int isp_layernorm_rows_bwd(const float* dy, long long lddy, const void* x, int x_bf16, long long ldx, const float* gamma,
                           const float* resid, long long ldr, float* dx, long long lddx, void* dx_bf16, long long ldb,
                           long long M, int C, float eps, isp_stream_t stream);
/* Affine-parameter gradients of a trainable LayerNorm (simple_vit click embedding, simple_ViT.py:31-93):
 * dgamma[c] += sum_rows dy * xhat, dbeta[c] += sum_rows dy (accumulated; C <= 1024). */
int isp_layernorm_affine_bwd(const float* dy, long long lddy, const void* x, int x_bf16, long long ldx, float* dgamma,
                             float* dbeta, long long M, int C, float eps, isp_stream_t stream);
int isp_gelu_bwd_bf16(const void* dh, const void* pre, void* out, long long n, int quick, isp_stream_t stream);
int isp_softmax_rows(const float* S, long long lds, void* P_bf16, long long ldp, long long R, int ncols, int ncols_pad,
                     isp_stream_t stream);
int isp_attn_ds_rows(const void* P_bf16, long long ldp, const void* dP, int dp_bf16, long long lddp, void* dS_bf16,
                     long long ldds, long long R, int ncols, int ncols_pad, isp_stream_t stream);
int isp_transpose_bf16_batched(const void* src, long long lds, long long src_z, void* dst, long long ldd, long long dst_z,
                               int Z, int R, int C, isp_stream_t stream);

/* Flash-style attention on tcgen05: out = softmax(Q K^T) V, scores never leave the SM.
 * Q bf16 [B*rows_per_img, ldq] (pre-scaled by 1/sqrt(d)), head h at column h*q_head_stride;
 * K bf16 [B, heads, ceil128(nkeys), DKC] and Vt bf16 [B, heads, DV, ceil128(nkeys)], zero
 * padded (see isp_repack_heads); out bf16 [B*rows_per_img, ldo], head h writes DV columns at
 * h*o_head_stride.  variant 0: head_dim <= 64 (DKC = DV = 64); variant 1: head_dim <= 112
 * (DKC = 128, DV = 112).  Replaces nn.MultiheadAttention in LoftUp's CrossAttentionLayer
 * (loftup/layers.py:186-202) and Attention.forward of the ViT (dinov2/layers/attention.py:54-71). */
int isp_attention_bf16_tc(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                          void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                          int heads, int nkeys, int variant, isp_stream_t stream);
/* Same, and also writes lse fp32 [B][heads][rows_per_img]: log2 of sum_k 2^(s_k log2 e) of every score row, the
 * statistic isp_attention_bwd_bf16_tc recomputes the probabilities from. */
int isp_attention_bf16_tc_lse(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                              void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                              int heads, int nkeys, int variant, float* lse, isp_stream_t stream);
/* Tuned variant-1 entry (LoftUp cross-attention, loftup/layers.py:186-202).  lsum_col >= 0: Vt row `lsum_col` of every head
 * (one of the head's zero-padding rows, e.g. 101 for head_dim 101 in DV = 112) holds ones, so the softmax denominator
 * accumulates in that O column on the tensor pipe; poly in {0,2,3,4}: that many of every 8 exponentials are evaluated by an
 * FMA-pipe polynomial (relative error 1e-4) instead of MUFU.EX2.  lse may be NULL.  out column lsum_col receives 1.0. */
int isp_attention_bf16_tc_opt(const void* Q, long long ldq, int q_head_stride, const void* K, const void* Vt,
                              void* out, long long ldo, int o_head_stride, int B, long long rows_per_img,
                              int heads, int nkeys, int variant, float* lse, int lsum_col, int poly,
                              isp_stream_t stream);
/* D[B][heads][rows] = sum_d dO[row, h*HP + d] * O[row, h*HP + d] (bf16 [B*rows, ld] operands). */
int isp_attention_rowdot_heads(const void* dO, long long lddo, const void* O, long long ldo, float* out, int B,
                               long long rows, int heads, int HP, isp_stream_t stream);
/* Flash-style attention backward on tcgen05 (scores, probabilities and their gradients never leave the SM): the autograd of
 * nn.MultiheadAttention in LoftUp's CrossAttentionLayer (loftup/layers.py:186-202) under the trainer's backward
 * (core/training/trainer.py:213-221).  Q, dO bf16 [B*rows, ld], head h at columns [h*HP, +HP), Q pre-scaled as in the
 * forward; K, V bf16 [B, heads, nkeys, HP] (rows = keys); lse, dvec fp32 [B][heads][rows] (+ 64 floats of slack) from
 * isp_attention_bf16_tc_lse / isp_attention_rowdot_heads.  dK, dV fp32 [B, heads, nkeys, HP] and the optional dQ fp32
 * [B*rows, lddq] are ACCUMULATED into (zero them first).  HP: multiple of 16, <= 128. */
int isp_attention_bwd_bf16_tc(const void* Q, long long ldq, const void* dO, long long lddo, const void* K, const void* V,
                              const float* lse, const float* dvec, float* dK, float* dV, float* dQ, long long lddq, int B,
                              long long rows, int heads, int nkeys, int HP, isp_stream_t stream);

/* LayerNorm over the last dim of a row-major matrix (f32 or bf16 in/out, biased variance):
 * nn.LayerNorm in loftup/layers.py:26-35,161-202,222-228 and the channel LayerNorm :38-58.
 * Output columns C..ldo-1 are zero-filled. */
int isp_layernorm_rows(const void* in, int in_bf16, long long ldi, void* out, int out_bf16, long long ldo,
                       const float* gamma, const float* beta, long long M, int C, float eps, isp_stream_t stream);

/* MinMaxScaler statistics (loftup/layers.py:66-71): per-channel min/max of a 3-channel image
 * over batch and space -> mm6 (device int[6], order-preserving encoding consumed by
 * isp_loftup_fourier_chnorm).  sb/sc: batch/channel strides in elements; H*W must be dense. */
int isp_minmax_per_channel(const float* img, int* mm6, int B, int H, int W, long long sb, long long sc,
                           isp_stream_t stream);

/* LoftUp query producer: MinMaxScaler -> ImplicitFeaturizer(colour, 20 freqs, learned bias) ->
 * ChannelNorm(203) (loftup/layers.py:61-158, loftup/loftup.py:50-56) -> bf16 NHWC [B,H,W,ldo].
 * gridr[H], gridc[W], freqs20[20]: the host's torch.linspace / torch.exp tables;
 * bias_sin/bias_cos[100]: biases[0/1].flatten(); gamma/beta[203]. */
int isp_loftup_fourier_chnorm(const float* img, long long sb, long long sc, long long sh, long long sw,
                              const int* mm6, const float* gridr, const float* gridc, const float* freqs20,
                              const float* bias_sin, const float* bias_cos, const float* gamma,
                              const float* beta, void* out_bf16, int B, int H, int W, int ldo, float eps,
                              isp_stream_t stream);

/* LoftUp key/value source: optional ChannelNorm(C) of the LR features (loftup.py:141-149)
 * concatenated with the 20-channel sine PE of the LR grid (loftup.py:113-118):
 * lr [B,C,h,w] f32 (strided) -> out f32 [B*h*w, C+20].  cn_w/cn_b may be NULL. */
int isp_loftup_lr_prepare(const float* lr, long long sb, long long sc, long long sh, long long sw,
                          const float* cn_w, const float* cn_b, const float* gridr, const float* gridc,
                          const float* freqs5, const float* bias_sin, const float* bias_cos, float* out,
                          int B, int C, int h, int w, float eps, isp_stream_t stream);

/* Split heads out of a [B*T, ld] projection (columns col0 + head*head_dim + d) into the
 * attention kernel's operand layouts: transpose == 0 -> [B, heads, Tpad, D] (keys),
 * transpose == 1 -> [B, heads, D, Tpad] (values, transposed); bf16, zero padded. */
int isp_repack_heads(const void* src, int src_bf16, long long ld, int col0, int head_dim, void* dst_bf16,
                     int B, int T, int Tpad, int heads, int D, int transpose, isp_stream_t stream);

/* ---- LiFT helpers (core/model/upsamplers/LiFT.py:47-122) ------------------------------
 * conv3x3 stride 2 pad 1, Cout = 32, Cin <= 32, + bias (BatchNorm folded by the caller) + ReLU:
 * in f32 with element strides (sb,sc,sh,sw) -> out NHWC f32 [B,ceil(Hi/2),ceil(Wi/2),32];
 * w [32][Cin][3][3].  LiFT.image_convs_1 / image_convs_2 (LiFT.py:69-90). */
int isp_conv3x3_s2_c32(const float* in, long long sb, long long sc, long long sh, long long sw,
                       const float* w, const float* bias, float* out, int B, int Cin, int Hi, int Wi,
                       isp_stream_t stream);
/* isp_conv3x3_s2_c32 without the ReLU (bias = the conv's own bias): the raw output whose batch statistics a BatchNorm2d in
 * train() normalises with (LiFT.py:69-90 under core/training/trainer.py:213-214). */
int isp_conv3x3_s2_c32_raw(const float* in, long long sb, long long sc, long long sh, long long sw,
                       const float* w, const float* bias, float* out, int B, int Cin, int Hi, int Wi,
                       isp_stream_t stream);
/* F.adaptive_max_pool2d on NHWC f32 (LiFT.py:110) */
int isp_adaptive_maxpool_nhwc(const float* in, float* out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                              isp_stream_t stream);
/* dst[b,y,x,c] (strides db,dh,dw; channel stride 1) = src[b,c,y,x] (strides sb,sc,sh,sw), f32|bf16
 * either side: channel concat (LiFT.py:42,117), pixel shuffle of the k=2,s=2 transposed conv. */
int isp_copy_channels(const void* src, int src_bf16, long long sb, long long sc, long long sh, long long sw,
                      void* dst, int dst_bf16, long long db, long long dh, long long dw, int B, int C, int H,
                      int W, isp_stream_t stream);

/* IS-head classifier Conv2d(C -> K, 1x1) on an NHWC activation (heads/base_head.py:16,
 * conv_heads.py:72): out[m,k] = sum_c x[m,c]*w[k,c] + b[k]; x f32|bf16 [M,ld], K <= 8, out f32 [M,K]. */
int isp_rowdot(const void* x, int x_bf16, long long ld, const float* w, const float* b, float* out,
               long long M, int C, int K, isp_stream_t stream);

/* ViT patch embedding front half: [B,Cin,H,W] f32 -> bf16 [B*(H/P)*(W/P), ldo], column order
 * c*P*P + i*P + j (Conv2d(k=s=P) weight flattening; dinov2/layers/patch_embed.py:25-100,
 * featurizers/utils/patch_embed.py:36-42). */
int isp_vit_patchify(const float* img, long long sb, long long sc, long long sh, long long sw, void* out_bf16,
                     int B, int Cin, int H, int W, int P, int ldo, isp_stream_t stream);

/* ViT token assembly (DINOv2.py:518-529): out[b,0]=cls+pos[0]; out[b,1+n]=patch[b,n]+extra[b,n]+pos[1+n] */
int isp_vit_assemble_tokens(const float* patch, const float* extra, const float* cls, const float* pos,
                            float* out, int B, int N, int C, isp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ISP_B200_H */
