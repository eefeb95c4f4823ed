/* pangu_b200 -- C ABI of the B200 (sm_100a) Pangu-Weather hot path.
 *
 * The reference (zhaoshan2/pangu-pytorch) has no FFI layer: its hot path is the nn.Module
 * code in models/layers.py and models/pangu_model.py.  This library replaces the ATen op
 * chains of those modules with hand-written kernels; the Python host
 * (pangu_pytorch_b200/models/*.py) binds these entry points with ctypes and keeps the
 * reference's module API and state_dict layout.  Each entry point cites the reference
 * code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates,
 *     frees or retains memory.  Workspaces are caller-provided.
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation inside.
 *   - `fp16` selects the 16-bit operand format of activations/weights: 0 = bf16, 1 = fp16.
 *     Accumulation, LayerNorm, softmax and the residual stream are always fp32.
 *   - return 0 on success, negative on error (bad shape / alignment / CUDA error);
 *     pangu_last_error() returns a thread-local message.  No fallbacks: unsupported shapes
 *     are rejected.
 *   - token grids: (Z, H, W) is the un-padded grid (8,181,360 / 8,91,180 at 0.25 deg; W may be
 *     any multiple of 12).  "window order" means row = (lon_window*types + type)*144 + k with
 *     the +5 zero-pad latitude rows included; "natural order" means
 *     row = (z*H + h)*W + w.
 */
#ifndef PANGU_B200_H
#define PANGU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* pangu_last_error(void);
int pangu_version(void);
/* 0 when the current device is sm_100 and the kernels are loadable, negative otherwise. */
int pangu_check_device(void);

/* fp32 -> 16-bit cast of a [rows, k_src] matrix into [rows, k_dst] (zero padded columns).
 * Weight preparation for the nn.Linear / nn.Conv1d(k=1) weights (layout (out, in), as stored
 * in the reference state_dict; models/onnx2torch.py:41-44). */
int pangu_cast16(const float* src, void* dst, int rows, int k_src, int k_dst, int fp16, void* stream);

/* F.pad + torch.roll + window partition of EarthSpecificBlock.forward (models/layers.py:188-221)
 * applied to a plain fp32 residual stream: x32 [T,C] natural -> x16w [Tp,C] window order for
 * roll state `roll` (pad rows zero); roll < 0 = plain cast in natural order.  Only needed when a
 * block is entered from an fp32 tensor; inside the model the producing GEMM writes this layout. */
int pangu_to_window16(const float* x32, void* x16w, int Z, int H, int W, int C, int roll, int fp16, void* stream);

/* PatchEmbedding_pretrain.forward (models/layers.py:40-93).
 * Normalise + zero-pad + concat constants + 2x4x4 / 4x4 im2col into `ws_a_upper`
 * [7*Hh*Ww,192] and `ws_a_surface` [Hh*Ww,128], then two GEMMs (+bias).
 * Outputs: x32 [8*Hh*Ww,192] fp32 natural order; x16w: 16-bit copy in window order (roll 0)
 * for the first block's QKV GEMM (pad rows must have been zeroed once by the caller). */
int pangu_patch_embed(const float* upper, const float* surface,
                      const float* surface_mean, const float* surface_std,
                      const float* upper_mean, const float* upper_std,
                      const float* maps, const float* const_h,
                      const void* w_upper16 /*[192,192]*/, const float* b_upper,
                      const void* w_surface16 /*[192,128] K zero-padded*/, const float* b_surface,
                      void* ws_a_upper, void* ws_a_surface,
                      float* x32, void* x16w,
                      int lat, int lon, int fp16, void* stream);

/* EarthAttention3D.linear1 + head split + q*scale (models/layers.py:365-374).
 * x16w [Tp, C] window order -> qkv16 head-major [3C/32 planes][Tp rounded up to 128][32] (plane = s*heads + head,
 * s in {q,k,v}; window-order rows), q pre-scaled by 32^-0.5. */
int pangu_qkv(const void* x16w, const void* w16 /*[3C,C]*/, const float* bias /*[3C]*/,
              void* qkv16, int Z, int H, int W, int C, int fp16, void* stream);

/* q k^T + earth_specific_bias (+ shift mask, gen_mask models/layers.py:153-181) -> softmax ->
 * P v, heads merged (models/layers.py:378-415).  bias: the fp32 parameter
 * [1, types, heads, 144, 144] as stored in the state_dict.
 * qkv16: head-major [3*heads planes][Tp rounded up to 128][32] as written by pangu_qkv.
 * att16: window_order_out == 0 -> [Z*H*W, C] in NATURAL token order (window reverse, un-roll and crop of
 *        models/layers.py:227-243 folded into the store; pad rows dropped);
 *        window_order_out != 0 -> [Tp, C] in window order, pad rows included (EarthAttention3D.forward). */
int pangu_window_attention(const void* qkv16, const float* earth_bias, void* att16,
                           int Z, int H, int W, int C, int heads, int roll, int window_order_out, int fp16,
                           void* stream);

/* EarthAttention3D.linear2 + norm1 + residual (models/layers.py:418, 250) on the natural-order attention
 * output of pangu_window_attention:  x32[tok] += res_scale * LN1(att[tok] W2^T + b2).
 * Also writes x16 (natural order) = cast(x32) for the MLP. In-place on x32.  `roll` is unused (kept for ABI). */
int pangu_proj_ln_residual(const void* att16, const void* w16 /*[C,C]*/, const float* bias,
                           const float* gamma, const float* beta,
                           float* x32, void* x16, int Z, int H, int W, int C, int roll,
                           float res_scale, int fp16, void* stream);

/* Mlp.forward + norm2 + residual (models/layers.py:264-270, 251):
 *   x32 += res_scale * LN2(GELU(x W1^T + b1) W2^T + b2)      (exact erf GELU)
 * x16_in natural order.  ws_hidden [T, 4C]: receives the 16-bit GELU activations (the training tape keeps them); NULL:
 * the activation is not materialised -- one kernel keeps it on the SM (C = 192: mlp_fused_kernel, C = 384: the CTA-pair
 * mlp_fused2_kernel).  All device pointers 16 B aligned.
 * x16_out: 16-bit copy of the new
 * residual stream, in natural order when roll_out < 0, else in window order for a following
 * block with roll state roll_out (pad rows never written). */
int pangu_mlp_ln_residual(const void* x16_in, const void* w1_16 /*[4C,C]*/, const float* b1,
                          const void* w2_16 /*[C,4C]*/, const float* b2,
                          const float* gamma, const float* beta,
                          void* ws_hidden, float* x32, void* x16_out,
                          int Z, int H, int W, int C, int roll_out,
                          float res_scale, int fp16, void* stream);

/* DownSample.forward (models/layers.py:432-459): pad + 2x2 merge + LN(4C) -> ws_a [T2,4C];
 * GEMM (no bias) -> x32_out [T2, 2C] natural, x16w_out window order (roll 0) on the low grid. */
int pangu_downsample(const float* x32_in, const float* gamma, const float* beta,
                     const void* w16 /*[2C,4C]*/, void* ws_a, float* x32_out, void* x16w_out,
                     int Z, int H, int W, int C, int fp16, void* stream);

/* UpSample.forward (models/layers.py:474-499): GEMM1 (no bias) with pixel-shuffle + crop +
 * LN(C_out) fused in the epilogue -> ws_a [T, C_out]; GEMM2 (no bias) -> x32_out natural,
 * x16w_out window order (roll 0) on the high grid (Z, H, W). x16_in: low grid, natural. */
int pangu_upsample(const void* x16_in, const void* w1_16 /*[4Co,Ci]*/, const float* gamma, const float* beta,
                   const void* w2_16 /*[Co,Co]*/, void* ws_a, float* x32_out, void* x16w_out,
                   int Z, int H, int W, int C_in, int C_out, int fp16, void* stream);

/* torch.cat((skip, x), -1) + PatchRecovery_pretrain.forward (models/pangu_model.py:81,
 * models/layers.py:511-545): the concat is folded into the GEMM K loop (two A sources);
 * un-patchify + crop (14->13 levels, 4*H->lat) in the epilogue.  Outputs normalised fields
 * out_upper [5,13,lat,lon], out_surface [4,lat,lon]. */
int pangu_patch_recover(const void* skip16, const void* x16, const void* w_upper16 /*[160,384]*/,
                        const float* b_upper, const void* w_surface16 /*[64,384]*/, const float* b_surface,
                        float* out_upper, float* out_surface,
                        int Z, int H, int W, int C, int lat, int lon, int fp16, void* stream);

/* normBackData (era5_data/utils_data.py:324-330) in place on the normalised model outputs
 * upper [5,13,lat,lon] / surface [4,lat,lon], with the INPUT-order statistics the model takes
 * (surface [4], upper [13,5] whose level axis is reversed w.r.t. the data; the level reversal of
 * weatherStatistics_output, :225-233, is applied inside).  Glue between autoregressive steps. */
int pangu_denorm_fields(float* upper, float* surface, const float* surface_mean, const float* surface_std,
                        const float* upper_mean, const float* upper_std, int lat, int lon, void* stream);

/* Training loss of models/pangu_sample.py:57-67: normData(target) (era5_data/utils_data.py:315-321) then
 * mean(|out - tgt| * w_upper) + 0.25 * mean(|out_s - tgt_s| * w_surface), weights era5_data/config.py:45-46.
 * Targets are PHYSICAL fields; statistics are the input-order arrays (level reversal applied inside).
 * upper_weights_host[5] / surface_weights_host[4] are HOST pointers (copied into the launch).  loss: device
 * float[1]; ws_acc: device double[2] scratch; grad_*: optional dL/d(output) (seed of the backward pass). */
int pangu_l1_loss(const float* out_upper, const float* out_surface, const float* tgt_upper, const float* tgt_surface,
                  const float* surface_mean, const float* surface_std, const float* upper_mean, const float* upper_std,
                  const float* upper_weights_host, const float* surface_weights_host, float* loss, double* ws_acc,
                  float* grad_upper, float* grad_surface, int lat, int lon, void* stream);

/* Evaluation scores of the reference's test loop (models/pangu_sample.py:236-270): latitude-weighted RMSE
 * (era5_data/score.py:92-105) and ACC (era5_data/score.py:123-135) per plane, p = var*13 + level for the 65
 * upper-air planes, 65 + var for the 4 surface planes; anomalies are taken against the scalar statistics mean of
 * the plane, as the reference does.  normalised != 0: out_* are the model's normalised outputs (normBackData is
 * applied on the fly); 0: physical fields.  tgt_* are physical.  lat_weights[lat]: the weights of
 * latitude_weighting_factor_torch (host-computed, caller-owned device array).  ws_acc: double[276] scratch.
 * rmse / acc: float[69]. */
int pangu_scores(const float* out_upper, const float* out_surface, const float* tgt_upper, const float* tgt_surface,
                 const float* surface_mean, const float* surface_std, const float* upper_mean, const float* upper_std,
                 const float* lat_weights, double* ws_acc, float* rmse, float* acc, int lat, int lon, int normalised,
                 void* stream);

/* Generic nn.Linear forward used by the stand-alone module API (EarthAttention3D.linear2,
 * Mlp.linear1/2 outside the fused block path) and by the unit tests of the tcgen05 GEMM engine:
 * out = a16 [M,K] * w16 [N,K]^T + bias.  gelu == 0: fp32 out32 and 16-bit out16 (both required),
 * N % 192 == 0.  gelu != 0: exact-erf GELU applied, 16-bit out16 only, N % 256 == 0.  K % 64 == 0.
 * gelu == 0 with out32 == NULL: 16-bit out16 only, N % 256 == 0 (pre-activation recompute of the backward). */
int pangu_linear(const void* a16, const void* w16, const float* bias, float* out32, void* out16,
                 int M, int N, int K, int gelu, int fp16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward pass.  The reference obtains these through autograd over
 * models/layers.py (loss.backward(), models/pangu_sample.py:69); here every gradient is a hand-written
 * kernel.  Parameter gradients are ACCUMULATED into caller-provided fp32 buffers (the .grad semantics).
 * ------------------------------------------------------------------------------------------------ */

/* fp32 [rows, cols] -> 16-bit TRANSPOSED copy [cols_pad, rows_pad] (zero padded): the B operands
 * (W^T) of the dgrad GEMMs, from the (out, in) weights of the state_dict. */
int pangu_cast16_t(const float* src, void* dst, int rows, int cols, int rows_pad, int cols_pad, int fp16, void* stream);

/* Data-gradient GEMM  out[M, N] = a16[M, K] * wt16[N, K]^T (+ bias),  wt16 = transposed weight copy:
 *   kind 0: out32 = (resid32 ? resid32 : 0) + acc, identity rows                         (N % 192 == 0)
 *   kind 1: out16 row-major                                                                (N % 256 == 0)
 *   kind 2: a16 rows are in WINDOW order (roll state `roll`): out32[token] = resid32[token] + acc,
 *           pad rows dropped -- backward of pad + roll + window partition (models/layers.py:188-221)
 *   kind 3: a16 rows natural; out16 rows scattered to WINDOW order (pad rows are not written) --
 *           backward of window reverse + un-roll + crop (models/layers.py:227-243)
 * K % 64 == 0.  (Z, H, W) is the token grid of the row maps (kinds 2, 3). */
int pangu_dgrad(const void* a16, const void* wt16, const float* bias, const float* resid32, float* out32, void* out16,
                int M, int N, int K, int kind, int Z, int H, int W, int roll, int fp16, void* stream);

/* Weight-gradient GEMM  dw[n, k_off + k] += alpha * sum_m dy16[m, n] * x16[m, k]  for n < N, k < K
 * (autograd of F.linear w.r.t. the weight).  dy16 [M, ld_dy], x16 [M, ld_x] 16-bit; dw fp32 [N, ldw]. */
int pangu_wgrad(const void* dy16, int ld_dy, const void* x16, int ld_x, float* dw, int ldw, int k_off,
                int M, int N, int K, float alpha, int fp16, void* stream);

/* Bias gradient: out[n] += alpha * sum_m src16[m, n], n < n_valid (N columns scanned, N % 8 == 0). */
int pangu_colsum16(const void* src16, int ld, float* out, int M, int N, int n_valid, float alpha, int fp16, void* stream);

/* LayerNorm backward (nn.LayerNorm, eps 1e-5; models/layers.py:141-142, 429, 472).  y: pre-norm values,
 * g: fp32 gradient w.r.t. the normalised output [rows, C], scale: DropPath factor of the branch.
 *   mode 0: y [rows, C], dx16 [rows, C]                                   (block norm1 / norm2; C 192 | 384)
 *   mode 1: UpSample.norm: rows = high-res tokens, y = linear1 output [T2, 4C] before the pixel shuffle,
 *           dx16 [T2, 4C] in the same layout (cropped positions are not written)        (C 192)
 *   mode 2: DownSample.norm: rows = low-res tokens, y = high-res stream x [T, C/4] fp32 (2x2 merge + pad
 *           recomputed), dx32 [T, C/4] is ADDED to                                      (C 768)
 * dgamma / dbeta [C] are accumulated into, multiplied by palpha (1 / loss scale; 1 for bf16).  dbias (nullable, mode 0):
 * [C] += palpha * sum_rows dx, the bias gradient of the linear layer that produced y (saves a pass over dx). */
int pangu_layernorm_bwd(const float* y, const float* g, const float* gamma, void* dx16, float* dx32,
                        float* dgamma, float* dbeta, float* dbias, int rows, int C, int mode, int Z, int H, int W,
                        float scale, float palpha, int fp16, void* stream);

/* GELU backward in place: dh16[m, n] *= gelu'(pre16[m, n]) (exact erf form, nn.GELU; models/layers.py:261), [M, N] 16-bit.
 * dbias (nullable): [N] += alpha * column sums of the result, the bias gradient of Mlp.linear1. */
int pangu_gelu_bwd(void* dh16, const void* pre16, int M, int N, float* dbias, float alpha, int fp16, void* stream);

/* Backward of pangu_window_attention (autograd of models/layers.py:368-415).  datt16w: [Tp, C] gradient of the
 * merged-head attention output in WINDOW order (pad rows zero); dqkv16: [Tp, 3C] window order, column
 * s*C + head*32 + d, gradient w.r.t. the un-scaled linear1 output; dbias [types, heads, 144, 144] += palpha * dS (nullable);
 * dbqkv [3C] += palpha * column sums of dqkv, the gradient of attention.linear1.bias (nullable). */
int pangu_window_attention_bwd(const void* qkv16, const void* datt16w, const float* earth_bias, void* dqkv16,
                               float* dbias, float* dbqkv, int Z, int H, int W, int C, int heads, int roll, float palpha,
                               int fp16, void* stream);

/* Backward of the un-patchify + crop of PatchRecovery_pretrain (models/layers.py:519-545): output-field
 * gradients * scale (loss scale for fp16 operands) -> dy_upper [7*Hh*Ww, 192] (features >= 160 zero),
 * dy_surface [Hh*Ww, 128] (features >= 64 zero). */
int pangu_recover_grad_gather(const float* d_upper, const float* d_surface, void* dy_upper, void* dy_surface,
                              int lat, int lon, float scale, int fp16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANGU_B200_H */
