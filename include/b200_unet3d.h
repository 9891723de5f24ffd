/* b200_unet3d.h — C ABI of the B200-native (sm_100a) 3D U-Net hot path.
 *
 * The reference (qwertyhgb/Prostate-Cancer-Multimodal-Segmentation) has no FFI: every op on its hot path is a
 * stock torch.nn call.  Each entry point below replaces one of those call sites (cited per function, paths
 * relative to the reference root) and is what a ctypes binding in the reference would bind — see INTEGRATION.md.
 *
 * Conventions
 *   - Activations are NDHWC bf16 "views": a base pointer to channel 0 of voxel (0,0,0,0), extents, and `ld`, the
 *     number of bf16 elements between consecutive voxels (ld >= c; ld > c addresses one half of a concat buffer).
 *   - All device memory is owned by the caller; nothing is allocated or freed here.  All work is enqueued on the
 *     given CUDA stream (a cudaStream_t passed as void*), no host synchronisation, CUDA-graph capturable.
 *   - Every function returns 0 on success, else a b200_status; b200_last_error() gives the thread-local message.
 *   - There is no CPU path and no cuDNN/cuBLAS path behind any of these.
 */
#ifndef B200_UNET3D_H
#define B200_UNET3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B200_OK = 0,
    B200_ERR_BAD_ARG = 1,
    B200_ERR_UNSUPPORTED_SHAPE = 2,
    B200_ERR_CUDA = 3,
    B200_ERR_DRIVER = 4
} b200_status;

typedef struct {
    void* ptr;   /* bf16* */
    int64_t n, d, h, w, c;
    int64_t ld;  /* voxel pitch in elements */
} b200_act;

/* conv3d fprop epilogue modes */
enum { B200_EPI_PLAIN = 0, B200_EPI_BIAS_STATS = 1, B200_EPI_AFFINE_RELU = 2, B200_EPI_BIAS = 3 };

const char* b200_last_error(void);
int b200_abi_version(void);
/* Programmatic dependent launch for every kernel of the library (default on): a kernel's CTAs may become resident while
   the previous kernel of the stream drains, and wait (griddepcontrol.wait) for its completion before touching global
   memory; under stream capture the launches become programmatic graph edges.  on = 1 / 0 sets the process-wide
   switch, on < 0 only queries; returns the previous setting.  No reference counterpart (torch launches serialise). */
int b200_set_pdl(int on);
/* The 64-column full-resolution convolutions (depth-marching kernels) on CTA pairs with tcgen05.mma.cta_group::2
   (csrc/dmarch2.cu; default on) or on single-CTA MMAs with a multicast weight stream (csrc/dmarch.cu).  Same results up
   to the fp32 accumulation order.  on < 0 only queries; returns the previous setting. */
int b200_set_dmarch_pair_mma(int on);
/* number of SMs of the current device (grid sizing), or <0 on error */
int b200_sm_count(void);

/* ---- layout packing ------------------------------------------------------------------------------------ */
/* (N,C,D,H,W) fp32 -> NDHWC bf16 view, channels c..out->c-1 zero (input of UNet3D.forward, models/unet3d.py:247) */
int b200_pack_input(const float* x_ncdhw, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w,
                    const b200_act* out, void* stream);
/* Conv3d weight (Cout,Cin,3,3,3) fp32 -> w_packed bf16 [27][Cout][cin_pad]; packed tap index t = kd*9 + kw*3 + kh
 * (kh fastest).  The one layout serves fprop (K-major B operand) and dgrad (MN-major B operand).
 * models/unet3d.py:29,35 */
int b200_pack_conv_weight(const float* w, int cout, int cin, int cin_pad, void* w_packed, void* stream);
/* ConvTranspose3d weight (Cin,Cout,2,2,2) fp32 -> w_fwd bf16 [8*Cout][Cin], w_dgrad bf16 [8][Cin][Cout],
 * bias (Cout) -> bias8 fp32 [8*Cout].  models/unet3d.py:120 */
int b200_pack_convt_weight(const float* w, const float* bias, int cin, int cout, void* w_fwd, void* w_dgrad,
                           float* bias8, void* stream);

/* First conv (Cin = 5 modalities): the 27*Cin-wide receptive field is expanded once into rows of k_pad bf16
 * (column k = c*27 + tap, i.e. the flattened (Cin,3,3,3) weight order) so that the layer runs as a plain GEMM:
 * b200_im2col_input + b200_pack_rows + b200_conv1_fprop / b200_conv1_wgrad.  models/unet3d.py:29 (inc.conv.0) */
int b200_im2col_input(const float* x_ncdhw, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w,
                      const b200_act* out, void* stream);
/* fp32 [rows][k] -> bf16 [rows][k_pad], zero padded */
int b200_pack_rows(const float* w, int rows, int k, int k_pad, void* out, void* stream);
/* pointwise (1 tap) forms of b200_conv3d_fprop / b200_conv3d_wgrad: y = x . w_rows^T (+ epilogue); dw (Cout,k_real) += */
int b200_conv1_fprop(const b200_act* x, const void* w_rows, const float* bias, const b200_act* y,
                     float* stats_partial, int mode, const float* scale, const float* shift, void* stream);
int b200_conv1_wgrad(const b200_act* x, const b200_act* dy, float* dw, int k_real, void* stream);

/* ---- 3x3x3 convolution, padding 1 (nn.Conv3d, models/unet3d.py:29,35), tcgen05 implicit GEMM ------------ */
/* number of 128-voxel output bricks of a volume */
int64_t b200_conv3d_mtiles(int64_t n, int64_t d, int64_t h, int64_t w);
/* rows of stats_partial that fprop(BIAS_STATS) writes for this problem (one per persistent CTA, or per block of the
 * split-K finalize pass when a workspace will be passed), <0 on error; ntaps = 27 for b200_conv3d_fprop, 1 for
 * b200_conv1_fprop */
int b200_conv3d_stat_rows(int64_t n, int64_t d, int64_t h, int64_t w, int64_t cout, int ntaps, int with_workspace);
/* Split-K for the deep levels (a handful of 128-voxel bricks, K = 27 * Cin large): bytes of the caller-owned fp32
 * workspace [splits][voxels][out_cols] that b200_conv3d_fprop / _dgrad use for this problem, 0 when the problem does
 * not split.  Scratch only: no initial state needed, nothing kept between calls; the partial tiles are added in a fixed
 * order (results are run-to-run identical).  Without a workspace (NULL) the same problem runs un-split. */
int64_t b200_conv3d_workspace_bytes(int64_t n, int64_t d, int64_t h, int64_t w, int64_t out_cols);
/* mode BIAS_STATS : y = bf16(conv + bias); stats_partial[row][Cout][2] = (sum, sum of squares) of stored y
 * mode AFFINE_RELU: y = relu(conv * scale + shift)   (eval-mode BatchNorm + bias folded)
 * mode BIAS / PLAIN likewise without statistics. */
int b200_conv3d_fprop(const b200_act* x, const void* w_packed, const float* bias, const b200_act* y,
                      float* stats_partial, int mode, const float* scale, const float* shift, void* workspace,
                      int64_t workspace_bytes, void* stream);
int b200_conv3d_dgrad(const b200_act* dy, const void* w_packed, const b200_act* dx, void* workspace,
                      int64_t workspace_bytes, void* stream);
/* The same input gradient for the layers the depth-marching CTA-pair kernel takes (dx of 64 or 32 channels at >= 8 x 16
 * in-plane; b200_conv3d_dgrad_kmajor_supported), from the TRANSPOSED packed weights w_packed_t = [27][Cin][Cout]
 * (b200_transpose_taps(w_packed, 27, Cout, Cin, ...)): the two CTAs of a pair split the weight tile by rows, which needs
 * the GEMM's K (here Cout) contiguous.  Autograd of the nn.Conv3d call sites models/unet3d.py:29,35. */
int b200_conv3d_dgrad_kmajor_supported(int64_t n, int64_t d, int64_t h, int64_t w, int64_t cin);
int b200_conv3d_dgrad_kmajor(const b200_act* dy, const void* w_packed_t, const b200_act* dx, void* stream);
/* bf16 [taps][rows][cols] -> [taps][cols][rows] */
int b200_transpose_taps(const void* src, int64_t taps, int64_t rows, int64_t cols, void* dst, void* stream);
/* dw += weight gradient; x->c may exceed cin_real (zero padded channels).
 * packed_layout = 0: dw is torch's (Cout, cin_real, 3,3,3);  1: dw is [27][Cout][cin_real] in the packed tap order
 * of b200_pack_conv_weight (the engine's physical parameter layout: contiguous, coalesced accumulation). */
int b200_conv3d_wgrad(const b200_act* x, const b200_act* dy, float* dw, int cin_real, int packed_layout,
                      void* stream);

/* ---- ConvTranspose3d k=2 s=2 (models/unet3d.py:120,134) + F.pad + cat written in place (:143-156) -------- */
/* y: view (upper channel half of the concat buffer) with the skip's extents; output voxel (2d+i+pad_d, ...) */
int b200_convt2x_fwd(const b200_act* x, const void* w_fwd, const float* bias8, const b200_act* y, int pad_d,
                     int pad_h, int pad_w, void* stream);
int b200_convt2x_dgrad(const b200_act* dy, int pad_d, int pad_h, int pad_w, const void* w_dgrad,
                       const b200_act* dx, void* stream);
/* dw fp32 (Cin, Cout, 2,2,2) += */
int b200_convt2x_wgrad(const b200_act* x, const b200_act* dy, int pad_d, int pad_h, int pad_w, float* dw,
                       void* stream);

/* ---- BatchNorm3d + ReLU (models/unet3d.py:31-33,37-39) -------------------------------------------------- */
/* reduce conv-epilogue partials; train-mode batch statistics, running-stat update (momentum, unbiased var),
 * scale = gamma*rstd, shift = beta - mean*scale.  running_* may be NULL.  num_batches_tracked (int64, nullable) is
 * incremented by one in the same launch (nn.BatchNorm3d's counter). */
int b200_bn_finalize(const float* stats_partial, int64_t rows, int64_t count, int c, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* mean, float* rstd, float* scale, float* shift,
                     void* stream);
/* eval mode: scale = gamma/sqrt(rv+eps), shift = beta + (conv_bias - rm)*scale */
int b200_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      const float* conv_bias, float eps, int c, float* scale, float* shift, void* stream);
/* out = relu(y*scale + shift) */
int b200_bn_apply_relu(const b200_act* y, const float* scale, const float* shift, const b200_act* out, void* stream);
/* backward, pass 1: partial[blk][c][2] = (sum dy_m, sum dy_m*xhat), dy_m = dout * (y*scale+shift > 0).
 * returns number of partial rows written through *nblk (<= b200_bn_bwd_max_blocks()). */
int b200_bn_bwd_max_blocks(void);
int b200_bn_bwd_reduce(const b200_act* dout, const b200_act* y, const float* scale, const float* shift,
                       const float* mean, const float* rstd, float* partial, int* nblk, void* stream);
/* dgamma += sum dy_m*xhat ; dbeta += sum dy_m ; coef[c][2] = (sum dy_m / count, sum dy_m*xhat / count) */
int b200_bn_bwd_finalize(const float* partial, int nblk, int c, int64_t count, float* dgamma, float* dbeta,
                         float* coef, void* stream);
/* pass 2: dy = gamma*rstd*(dy_m - coef0 - xhat*coef1) (bf16) ; dbias[c] += sum of stored dy */
int b200_bn_bwd_apply(const b200_act* dout, const b200_act* y, const float* scale, const float* shift,
                      const float* mean, const float* rstd, const float* gamma, const float* coef,
                      const b200_act* dy, float* dbias, void* stream);

/* ---- fused forms: these passes are HBM-bound, so they are made cheaper by never writing what can be recomputed ---- */
/* encoder block (models/unet3d.py:37-39 followed by :80): out = relu(y*scale + shift) and pooled = MaxPool3d(2)(out)
 * in one pass (= b200_bn_apply_relu + b200_maxpool3d_fwd without re-reading out). */
int b200_bn_apply_relu_pool(const b200_act* y, const float* scale, const float* shift, const b200_act* out,
                            const b200_act* pooled, void* stream);
/* BatchNorm backward of the network's last BatchNorm whose incoming gradient is the head's input gradient
 * dout = dlogits . w: equals b200_head_bwd + b200_bn_bwd_reduce / _apply (3 streams instead of 7); the reduce pass
 * also accumulates the head's dw (ncls, C) += dlogits^T relu(bn(y)) and db (ncls) += sum dlogits. */
int b200_bn_bwd_reduce_head(const float* dlogits, const float* w, int ncls, const b200_act* y, const float* scale,
                            const float* shift, const float* mean, const float* rstd, float* partial, int* nblk,
                            float* dw, float* db, void* stream);
int b200_bn_bwd_apply_head(const float* dlogits, const float* w, int ncls, const b200_act* y, const float* scale,
                           const float* shift, const float* mean, const float* rstd, const float* coef,
                           const b200_act* dy, float* dbias, void* stream);

/* ---- MaxPool3d(2) (models/unet3d.py:80) ---------------------------------------------------------------- */
int b200_maxpool3d_fwd(const b200_act* x, const b200_act* y, void* stream);
/* dx = dskip (may be NULL) + scatter(dy) to the first maximum in d,h,w scan order; voxels not covered by a
 * window (odd extents) get dskip only. */
int b200_maxpool3d_bwd(const b200_act* x, const b200_act* y, const b200_act* dy, const b200_act* dskip,
                       const b200_act* dx, void* stream);

/* ---- head Conv3d 1x1x1 (models/unet3d.py:222,295) ------------------------------------------------------ */
/* logits fp32 (N,ncls,D,H,W) = w (ncls,C) . x + b ; if probs != NULL also sigmoid(logits) (predict, :298-318) */
int b200_head_fwd(const b200_act* x, const float* w, const float* b, int ncls, float* logits, float* probs,
                  void* stream);
/* dx bf16 = dlogits . w ; dw += dlogits^T x ; db += sum dlogits */
int b200_head_bwd(const b200_act* x, const float* w, int ncls, const float* dlogits, const b200_act* dx, float* dw,
                  float* db, void* stream);

/* ---- sigmoid-BCE + Dice (utils/losses.py:44-92,107-152) ------------------------------------------------ */
/* sums[4] = (sum softplus(z)-z*t, sum sigmoid(z)*t, sum sigmoid(z), sum t) over all elements;
 * loss[0] = bce_w * sums0/n + dice_w * (1 - (2*sums1+smooth)/(sums2+sums3+smooth)).  workspace: >= 4*1024 floats */
int b200_loss_fwd(const float* logits, const float* target, int64_t n, float bce_w, float dice_w, float smooth,
                  float* workspace, float* sums, float* loss, void* stream);
/* dlogits = gout[0] * dL/dz using sums from the forward */
int b200_loss_bwd(const float* logits, const float* target, int64_t n, float bce_w, float dice_w, float smooth,
                  const float* sums, const float* gout, float* dlogits, void* stream);

/* ---- optimizer (torch.optim.Adam, utils/trainer.py:113-117,192) ---------------------------------------- */
/* fused over a flat fp32 buffer: g = grad*grad_scale + wd*p ; Adam moments ; bias-corrected update.
 * step is the 1-based step count.  If found_inf != NULL and *found_inf != 0 the update is skipped.
 * bf16_shadow (nullable): bf16 copy of the updated parameters written in the same pass.
 * dyn_scalars (nullable): device float[3] = (lr / (1 - beta1^step), sqrt(1 - beta2^step), grad_scale) read by the kernel
 * instead of the values derived from lr / step / grad_scale, so that one captured launch (CUDA graph) serves every step. */
int b200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                   double grad_scale,
                   const float* found_inf, void* bf16_shadow, const float* dyn_scalars, void* stream);
/* out[i] = bf16(x[i]) — operand shadow of the flat parameter buffer when a foreign optimizer updated it */
int b200_cast_bf16(const float* x, int64_t n, void* out, void* stream);
/* sum of squares of a flat fp32 buffer -> out[0] (+=) ; nonfinite flag -> out[1] (clip_grad_norm_/GradScaler) */
int b200_sumsq(const float* x, int64_t n, float* out, void* stream);

/* ---- misc bandwidth helpers ---------------------------------------------------------------------------- */
int b200_fill_zero(const b200_act* v, void* stream);
/* out[c] += sum over voxels of v[.,c]  (ConvTranspose3d bias gradient) */
int b200_channel_sum(const b200_act* v, float* out, void* stream);
/* NDHWC bf16 view -> (N,C,D,H,W) fp32 (debug / per-layer parity taps) */
int b200_unpack_act(const b200_act* v, float* out_ncdhw, void* stream);

/* ---- "next" rows: input pipeline and validation metrics on the device -------------------------- */
/* nvol contiguous fp32 volumes (d_in,h_in,w_in) -> (d_out,h_out,w_out) with the semantics of the reference's
 * sitk.ResampleImageFilter call (script/data_loader.py:240-283 image: linear; :395-409 label: nearest then > 0):
 * output index i reads the continuous input index i*in/out per axis, the upper neighbour is clamped to the last index,
 * points at or beyond size-0.5 get the default value 0.  nearest != 0: floor(index+0.5).  binarize != 0: out = out > 0. */
int b200_resample3d(const float* in, int64_t nvol, int64_t d_in, int64_t h_in, int64_t w_in, float* out,
                    int64_t d_out, int64_t h_out, int64_t w_out, int nearest, int binarize, void* stream);
/* in place, per volume: (x - min) / (max - min), a constant volume becomes zeros (script/predict.py:69-75).
 * workspace: >= 8 * nvol bytes */
int b200_minmax_normalize(float* x, int64_t nvol, int64_t voxels_per_volume, void* workspace, void* stream);
/* counts[s][3] (int64, +=) = (|P & T|, |P|, |T|) with P = score > threshold, T = label > 0.5, per sample: the integer
 * sums behind calculate_dice_score / calculate_iou (script/validate_model.py:24-95) */
int b200_seg_counts(const float* score, const float* label, int64_t nsamples, int64_t voxels_per_sample,
                    float threshold, int64_t* counts, void* stream);

/* First conv of the 5-modality network straight from the fp32 (N,5,D,H,W) input: the im2col rows are built in shared
 * memory inside the GEMM kernels (never in HBM).  w_rows = b200_pack_rows of the (Cout,5,3,3,3) weight, k_pad = 144;
 * modes as b200_conv3d_fprop; dw is fp32 [Cout][135] (+=).  models/unet3d.py:194 (inc = DoubleConv3D(5, 64)), :29.
 * b200_conv1_direct_supported(c, cout, w) != 0 tells whether this form exists for the channel counts (c == 5) and row
 * length w; other thin inputs use b200_im2col_input + b200_conv1_fprop / b200_conv1_wgrad.  BIAS_STATS needs
 * b200_conv1_direct_stat_rows(...) rows of stats_partial. */
int b200_conv1_direct_supported(int64_t c, int64_t cout, int64_t w);
int b200_conv1_direct_stat_rows(int64_t n, int64_t d, int64_t h, int64_t w, int64_t cout);
int b200_conv1_direct_fprop(const float* x, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w, const void* w_rows,
                            const float* bias, const b200_act* y, float* stats_partial, int mode, const float* scale,
                            const float* shift, void* stream);
int b200_conv1_direct_wgrad(const float* x, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w, const b200_act* dy,
                            float* dw, void* stream);

/* Forward of the same first conv as a depth-marching kernel (csrc/conv1_march.cu; models/unet3d.py:194, :29): a CTA
 * walks an 8 x 16 brick column along depth and builds ONE 45-row slice image per input slice, which feeds three output
 * slices; two epilogue warpgroups.  For 5 input channels and Cout <= 64 (b200_conv1_march_supported); w_slices =
 * b200_pack_conv1_slices of the (Cout,5,3,3,3) weight: bf16 [3 kd][Cout][64], row k = c*9 + kh*3 + kw.  Modes and
 * arguments as b200_conv1_direct_fprop; BIAS_STATS needs b200_conv1_march_stat_rows(...) rows of stats_partial. */
int b200_conv1_march_supported(int64_t c, int64_t cout);
int b200_conv1_march_stat_rows(int64_t n, int64_t d, int64_t h, int64_t w, int64_t cout);
int b200_pack_conv1_slices(const float* w, int64_t cout, int64_t cin, void* out, void* stream);
int b200_conv1_march_fprop(const float* x, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w, const void* w_slices,
                           const float* bias, const b200_act* y, float* stats_partial, int mode, const float* scale,
                           const float* shift, void* stream);
/* ... and its weight gradient by the same march (autograd of models/unet3d.py:29 for inc): the slice images are the
 * K-major operand, the 8 x 16 dy bricks the other; dw is fp32 (Cout,5,3,3,3) (+=), accumulated in TMEM over every
 * slice a CTA visits and added once. */
int b200_conv1_march_wgrad(const float* x, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w, const b200_act* dy,
                           float* dw, void* stream);

/* per-channel sum over the box [d0,d0+bd) x [h0,h0+bh) x [w0,w0+bw) of every sample, added to out[c] (fp32): the
 * ConvTranspose3d bias gradient when F.pad (models/unet3d.py:149-151) put a zero border around the upsampled map */
int b200_channel_sum_box(const b200_act* v, int d0, int h0, int w0, int bd, int bh, int bw, float* out, void* stream);

/* ---- sliding-window inference (BASELINE configs[3]; script/predict.py:152-172 predicts whole volumes) ------------ */
/* origins: device int32 [nwin][4] = (volume, d0, h0, w0).  gather: x (n,c,d,h,w) fp32 -> out (nwin,c,wd,wh,ww) fp32.
 * accumulate: acc (n,k,d,h,w) += window logits (nwin,k,wd,wh,ww), windows added in list order per voxel (no atomics),
 * for volumes [v_lo, v_lo+v_cnt).  finalize: acc[(nk),d,h,w] /= cover[d]*cover[D+h]*cover[D+H+w] (device int32, the
 * per-axis number of covering windows), then probs = sigmoid, mask = probs > threshold (either may be null). */
int b200_window_gather(const float* x, int64_t n, int64_t c, int64_t d, int64_t h, int64_t w, const int32_t* origins,
                       int nwin, int64_t wd, int64_t wh, int64_t ww, float* out, void* stream);
int b200_window_accumulate(const float* logits, const int32_t* origins, int nwin, int64_t k, int64_t wd, int64_t wh,
                           int64_t ww, float* acc, int64_t n, int64_t d, int64_t h, int64_t w, int v_lo, int v_cnt,
                           void* stream);
int b200_window_finalize(float* acc, const int32_t* cover, int64_t nk, int64_t d, int64_t h, int64_t w, float threshold,
                         float* probs, float* mask, void* stream);

/* routing queries (instrumentation): 3x3x3 conv fprop / dgrad -> 0 igemm_kernel, 1 dmarch_kernel (64 output columns on
 * 8 x 16 bricks), 2 igemm_pair_kernel (tiles of >= 128 columns: CTA pairs, cta_group::2); weight gradient -> 0 wgrad_kernel, 1 wgrad_halo_kernel (w >= 8 and h >= 16) */
int b200_conv3d_kernel_id(int64_t n, int64_t d, int64_t h, int64_t w, int64_t out_cols);
int b200_conv3d_wgrad_kernel_id(int64_t h, int64_t w);

/* The tcgen05 micro-probes and kernel ablation switches used during development are NOT part of this ABI: they are
 * compiled only into the development library (libb200unet3d_dev.so, -DB200_DEV) and declared in csrc/dev_api.h. */

#ifdef __cplusplus
}
#endif
#endif
