/* change3d_b200 — C ABI of the B200-native X3D change-detection hot path.
 *
 * The reference (zhuduowang/Change3D) has no FFI/plugin layer: its hot path is torch.nn modules
 * (model/x3d.py, model/change_decoder.py, model/trainer.py).  This header is the boundary a
 * maintainer binds instead of those modules' forward/backward: plain device pointers + sizes,
 * a cudaStream_t passed as void*, int status return (0 = ok, C3D_ERR_* otherwise), no C++
 * exceptions, no torch types.  All tensors are fp32, activations NDHWC ([N,T,H,W,Cs], channel
 * stride Cs a multiple of 4; pad lanes must be zero).  Every launch is stream-ordered and
 * re-entrant across streams; buffers are borrowed for the duration of the launch only.
 *
 * Each entry cites the reference code it replaces.
 */
#ifndef CHANGE3D_B200_H
#define CHANGE3D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Stays 1 while the library and its only binder (change3d_b200/_lib.py) ship together; descriptors only ever grow by
 * appending fields (c3d_gemm_desc.flags was appended in round 1), entry points are only added. */
#define C3D_ABI_VERSION 1

/* status codes */
#define C3D_STATUS_OK 0
#define C3D_STATUS_BAD_ARGUMENT 1
#define C3D_STATUS_CUDA_ERROR 2
#define C3D_STATUS_SHARED_MEMORY 3

int c3d_version(void);

/* ---- operand of the pointwise-GEMM family --------------------------------------------------
 * A GEMM row is one NDHWC pixel.  `mode` fuses a prologue into the global->shared staging:
 *   0 none | 1 relu(bn(A)) | 2 swish(gate * bn(A)) | 3 BN-backward: scale*(A - c1 - yhat(A2)*c2)
 *   | 4 |A - A2|
 * `map` selects the row addressing:
 *   0 dense (img_stride between images) | 1 stride-2 spatial subsample
 *   (2 and 3 were the ConvTranspose2d gather maps of round 1; the transposed convolutions are dense GEMMs +
 *   c3d_convt_col2im / c3d_convt_im2col now and the maps are rejected)
 * bnp = float[4][ld] (mean, rstd, gamma*rstd, beta) from c3d_bn_finalize; coef = float[2][ld].
 */
typedef struct c3d_operand {
  const float* A;
  const float* A2;
  const float* bnp;
  const float* coef;
  const float* gate;          /* [samples][ld] or NULL */
  int mode, map;
  int ld;                     /* channel stride of A / A2 */
  int OH, OW;                 /* GEMM-row grid per image */
  int IH, IW;                 /* source grid per image */
  long long img_stride;       /* elements between images of A */
  long long img_stride2;      /* ... of A2 (0 = same as A) */
  int frames_per_sample;      /* gate row = image / frames_per_sample */
  int seg0, nseg;             /* reserved (gather maps of round 1): must be 0 */
} c3d_operand;

/* Y[M x Ns] = epilogue( prologue(A)[M x K] * Wt[K x N] ),  Wt[red][out] = W[red*w_sr + out*w_so]
 * epi: 0 store (+ per-channel sum / sum-of-squares into stats[2][Ns], double, atomically added)
 *      1 Y += relu(acc) in place, previous Y saved to Y2                 (Encoder.enhance, model/trainer.py:88-108)
 *      2 Swish/SE/BN_b backward: du = acc * swish'(gate*bn(E1)); stats[sample][2][Ns] += (du, du*yhat)
 *      3 Y = acc + E1 (+ E2 upsampled x2 at even pixels)                 (residual / shortcut gradient join)
 *      4 reserved (ConvTranspose2d scatter of round 1; rejected)
 * Replaces nn.Conv3d 1x1x1 conv_a / conv_c / branch1_conv forward and dgrad (model/x3d.py:173-175,214-216,301-311)
 * and the decoder / enhance 1x1 Conv2d. */
typedef struct c3d_gemm_desc {
  c3d_operand a;
  const float* W;
  long long w_sr, w_so, w_cls_stride;   /* w_cls_stride: reserved, ignored */
  int Kred;                   /* logical reduction length (0 = a's staged width) */
  int N, Ns;                  /* logical / strided output channels */
  long long M;                /* GEMM rows */
  float* Y;
  long long out_img_stride;   /* 0 = dense */
  int epi;
  double* stats;              /* may be NULL */
  const float* E1;
  long long e1_img_stride;
  const float* E2;
  const float* ebnp;
  const float* egate;
  const float* bias;          /* reserved, ignored */
  float* Y2;
  long long rows_per_sample;  /* epi 2: rows per batch sample */
  int flags;                  /* C3D_GEMM_* */
} c3d_gemm_desc;

/* W is a model parameter: no kernel earlier in this training / inference step writes it, so the launch may
 * stage it ahead of stream order (programmatic dependent launch, overlapping the previous kernel's tail).
 * Leave it clear for weights produced on the stream (re-laid-out copies, freshly loaded checkpoints ...). */
#define C3D_GEMM_W_CONSTANT 1

int c3d_pw_gemm(const c3d_gemm_desc* desc, void* cuda_stream);

/* dW[n*dw_sn + k*dw_sk] += sum_rows P[row][n] * Q[row][k]   (fp32 atomics; dW must be zeroed by the caller)
 * Replaces the weight gradient of the same convolutions (autograd of model/x3d.py:173-175,214-216,301-311;
 * model/change_decoder.py:30-45; model/trainer.py:57-69). */
typedef struct c3d_wgrad_desc {
  c3d_operand p, q;
  long long M;
  float* dW;
  long long dw_sn, dw_sk;
  int N, K;
} c3d_wgrad_desc;

int c3d_pw_wgrad(const c3d_wgrad_desc* desc, void* cuda_stream);

/* ---- BatchNorm3d statistics -> parameter block -------------------------------------------------
 * nn.BatchNorm3d(eps=1e-5, momentum=0.1) (model/x3d.py:94-98,176-180,203-208,217-221,296-298).
 * stats = double[groups][2][Cs] (sum, sum of squares) accumulated by the producing kernel.
 * training != 0: batch statistics, running stats updated in place (unbiased variance);
 * training == 0: running statistics.  Writes bnp = float[4][Cs]; pad lanes (c >= C) are zeroed. */
int c3d_bn_finalize(const double* stats, int groups, long long count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, int C, int Cs, float momentum, float eps,
                    int training, float* bnp, void* cuda_stream);

/* BN_b finalize + SqueezeExcitation gate (fvcore SqueezeExcitation, call site model/x3d.py:194-202):
 * stats = double[N][2][Cs] per-sample sums from c3d_dw_conv_fwd; gate[n][c] = sigmoid(W2 relu(W1 pooled + b1) + b2),
 * pooled[n][c] = bn(mean_{T,H,W} y).  Saves zhat_mean[N][Cs], hidden[N][R], gate[N][Cs] for backward. */
int c3d_bn_se_finalize(const double* stats, int N, long long count_per_sample, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, int C, int Cs, float momentum, float eps, int training,
                       const float* w1, const float* b1, const float* w2, const float* b2, int R,
                       float* bnp, float* zhat_mean, float* hidden, float* gate, void* cuda_stream);

/* Y = relu( bnA(A) + [bnB(B) | B] ) elementwise over [M x Cs]; bnpB / B may be NULL.
 * ResBlock fusion + activation (model/x3d.py:326-327) and the stem's BN+ReLU (model/x3d.py:94-99). */
int c3d_bn_add_relu(const float* A, const float* bnpA, const float* B, const float* bnpB, float* Y,
                    long long M, int Cs, void* cuda_stream);

/* ---- depthwise 3x3x3 Conv3d (conv_b, model/x3d.py:184-193), stride (1,s,s), pad 1 --------------
 * input = relu(bn_a(Xraw)) applied on the fly (zero padding in the activated domain).
 * stats = double[N][2][Cs] per-sample (sum, sum of squares) of the raw output. w = torch layout [C][1][3][3][3]. */
int c3d_dw_conv_fwd(const float* X, const float* bnp_a, const float* w, float* Y, double* stats,
                    int N, int T, int IH, int IW, int C, int Cs, int stride, void* cuda_stream);

/* ---- stem (model/x3d.py:23-106): frame assembly (model/trainer.py:154-162) + Conv3d 1x3x3 (3->24)
 * + depthwise 5x1x1 temporal conv; raw output + BN statistics (double[2][24]).
 * Frame f of sample n, channel ci is the H*W plane at frame_ptr[f] + n*stride_n[f] + ci*stride_c[f]
 * (host arrays of T entries), so [pre, perception frames (stride_n = 0), post] is read in place.
 * Y = [B][T][H][W][24], T in 3..5. w_xy = conv.conv_t.weight [24][3][1][3][3], w_t = conv.conv_xy.weight [24][1][5][1][1]. */
int c3d_stem_fwd(const float* const* frame_ptr, const long long* stride_n, const long long* stride_c,
                 const float* w_xy, const float* w_t, float* Y, double* stats, int B, int T, int H, int W,
                 void* cuda_stream);

/* ---- decoder head (model/change_decoder.py:53-55,76-79): Conv2d 3x3 (C -> ncls, pad 1, no bias) + optional sigmoid */
int c3d_dec_head_fwd(const float* X, const float* w, float* Y, int B, int H, int W, int C, int ncls,
                     int apply_sigmoid, void* cuda_stream);

/* ============================== backward ============================== */

/* ReLU backward of a ResBlock / stem output fused with the BN backward reductions of the layer(s) below:
 * d_pre = dOut * (out > 0); stats_c[2][Cs] += (sum d_pre, sum d_pre*yhat_c); when the shortcut is normalised
 * (branch1_norm, model/x3d.py:296-298,312) stats_1 likewise with y_1 / bnp_1.  (autograd of model/x3d.py:326-327)
 * out = NULL: the mask is recomputed as (y_c - mean) * scale + beta > 0 from bnp_c (valid when `out` is exactly
 * relu(bn_c(y_c)), i.e. no shortcut: the stem).  d_pre = NULL: statistics only, nothing is stored. */
int c3d_relu_bwd_stats(const float* dOut, const float* out, const float* y_c, const float* bnp_c, const float* y_1,
                       const float* bnp_1, float* d_pre, double* stats_c, double* stats_1, long long M, int Cs,
                       void* cuda_stream);

/* BatchNorm backward coefficients: coef[2][Cs] = (mean d, mean d*yhat); dgamma = sum d*yhat; dbeta = sum d. */
int c3d_bn_bwd_finalize(const double* stats, int groups, long long count, int C, int Cs, float* coef, float* dgamma,
                        float* dbeta, void* cuda_stream);

/* SqueezeExcitation backward (+ BN_b backward coefficients).  stats = double[N][2][Cs] per-sample (sum du, sum du*zhat)
 * from c3d_pw_gemm epi 2.  gate == NULL: plain BN_b block.  Outputs coef[2][Cs], dgamma/dbeta[C], dpool[N][Cs]
 * (= dL/dpooled / (T*H*W)), and the SE parameter gradients dw1[R][C], db1[R], dw2[C][R], db2[C]. */
int c3d_se_bn_bwd_finalize(const double* stats, int N, long long count_per_sample, const float* bnp, const float* gamma,
                           const float* beta, const float* gate, const float* hidden, const float* zhat_mean,
                           const float* w1, const float* w2, int C, int Cs, int R, float* coef, float* dgamma,
                           float* dbeta, float* dpool, float* dw1, float* db1, float* dw2, float* db2, void* cuda_stream);

/* conv_b backward (autograd of model/x3d.py:184-193 with BN_b/SE on its output and ReLU+BN_a on its input):
 * dy_b = scale_b*(du*gate + dpool - c1 - zhat*c2); dr = conv_transpose(dy_b, w) * (bn_a(y_a) > 0);
 * dW[C][27] += (fp32 atomics, caller zeroes); stats_a[2][Cs] += (sum dr, sum dr*yhat_a).
 * `du` is scratch: the stride-2 and fallback paths overwrite it with dy_b in an elementwise pre-pass (the stride-1
 * row-streaming kernel applies that transform while it stages the rows and leaves du untouched). */
int c3d_dw_conv_bwd(float* du, const float* y_b, const float* bnp_b, const float* gate, const float* dpool,
                    const float* coef_b, const float* y_a, const float* bnp_a, const float* w, float* dr, float* dW,
                    double* stats_a, int N, int T, int IH, int IW, int C, int Cs, int stride, void* cuda_stream);

/* out[c] += sum_rows X[row][c]  (ConvTranspose2d bias gradient, model/change_decoder.py:32,38,44) */
int c3d_colsum(const float* X, long long M, int Cs, float* out, void* cuda_stream);

/* Stem backward (autograd of model/x3d.py:70-99 and of the frame assembly model/trainer.py:154-162).
 * y_raw = saved raw stem output, bnp[4][24] from c3d_bn_finalize, coef from c3d_bn_bwd_finalize.
 * relu_mask = 0: d_pre is the gradient AFTER the ReLU mask (d_pre of c3d_relu_bwd_stats).
 * relu_mask = 1: d_pre is the gradient w.r.t. the stem's ReLU output; the mask out > 0 is recomputed from y_raw and
 *                bnp with the forward's own expression, so the masked copy never exists in memory (pair it with
 *                c3d_relu_bwd_stats(out = NULL, d_pre = NULL) for the statistics).
 * dw_xy[24][27], dw_t[24][5], dperception[3][P][H][W] are accumulated with fp32 atomics (caller zeroes);
 * dperception may be NULL. */
int c3d_stem_bwd(const float* const* frame_ptr, const long long* stride_n, const long long* stride_c,
                 const float* d_pre, const float* y_raw, const float* bnp, const float* coef, const float* w_xy,
                 const float* w_t, float* dw_xy, float* dw_t, float* dperception, int B, int T, int H, int W,
                 int relu_mask, void* cuda_stream);

/* Decoder head backward (autograd of model/change_decoder.py:76-79): dpred/pred NCHW, X/dX NHWC, dW[ncls][C][3][3]
 * accumulated with fp32 atomics (caller zeroes). */
int c3d_dec_head_bwd(const float* dpred, const float* pred, const float* X, const float* w, float* dX, float* dW,
                     int B, int H, int W, int C, int ncls, int is_sigmoid, void* cuda_stream);

/* ConvTranspose2d(k=4, s=2, p=1) of the decoder up-blocks (model/change_decoder.py:30-45,71-73) split into a dense
 * GEMM and a gather: U = t x W with U[b][j][i][ky][kx][co] (c3d_pw_gemm, N = 16*cout), then
 * out[b][y][x][co] = bias[co] + skip[b][y][x][co] + the 4 entries of U with 2j-1+ky = y, 2i-1+kx = x.
 * skip may be NULL; skip_img_stride = elements between images of skip.  out is dense (B, 2h, 2w, cout). */
int c3d_convt_col2im(const float* U, const float* skip, long long skip_img_stride, const float* bias, float* out,
                     int B, int h, int w, int cout, void* cuda_stream);

/* Its mirror for the backward: V[b][j][i][ky][kx][co] = d_out[b][2j-1+ky][2i-1+kx][co] (zero outside), so that
 * d t = V x W^T and d W = t^T x V are dense (c3d_pw_gemm / c3d_pw_wgrad). */
int c3d_convt_im2col(const float* d_out, float* V, int B, int h, int w, int cout, void* cuda_stream);

/* torch.optim.Adam step (scripts/train_BCD.py:284-290: L2 weight decay added to the gradient) on flat, 16-byte
 * aligned fp32 buffers; g is multiplied by grad_scale first (1/world_size after a sum all-reduce), then clamped to
 * [-grad_clip, grad_clip] when grad_clip > 0 (clip_gradient, model/utils.py:481-491; scripts/train_CC.py:141-144). */
int c3d_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, float grad_clip, void* cuda_stream);

/* ============================== losses and metrics ==============================
 * Every forward is one streaming pass: partial sums go to a small device workspace `ws` (C3D_LOSS_WS_BYTES,
 * zeroed ONCE by the caller; the kernels leave it zeroed), the last CTA writes the loss and the backward
 * coefficients to `out` (device, float[C3D_LOSS_OUT_FLOATS]); nothing is read back to the host.
 * `gout` = device pointer to the upstream scalar gradient (NULL = 1), multiplied by the host float `gscale`. */
#define C3D_LOSS_WS_BYTES 128
#define C3D_LOSS_OUT_FLOATS 8

/* BCEDiceLoss (model/utils.py:154-169): out[0] = BCE(pred, target) + 1 - dice, out[4] = BCE, out[5] = dice,
 * out[1..3] = coefficients for the backward.  pred = probabilities (after the head's sigmoid), target in {0,1},
 * n elements.  When cm != NULL the same pass adds the 2x2 confusion matrix of (target, pred > 0.5) to
 * cm[2*gt + pr] (long long[4]) -- the `torch.where(output > 0.5, ...)` mask + ConfuseMatrixMeter.update_cm of
 * scripts/train_BCD.py:203-225 / utils/metric_tool.py:111-128 without the device->host copy. */
int c3d_bce_dice_fwd(const float* pred, const float* target, long long n, void* ws, float* out, long long* cm,
                     void* cuda_stream);
/* dpred[i] = g * dLoss/dpred[i]  (autograd of model/utils.py:154-169; BCE gradient clamped like aten: / max(p(1-p), 1e-12)) */
int c3d_bce_dice_bwd(const float* pred, const float* target, const float* out, const float* gout, float gscale,
                     float* dpred, long long n, void* cuda_stream);

/* CrossEntropyLoss2d (model/utils.py:171-178): NLLLoss(log_softmax(logits, dim 1), target, ignore_index, 'mean').
 * logits (B, C, HW) with batch_stride elements between samples (0 = C*HW; lets `mask[:, 1:]`-style views pass),
 * C <= 16; target int64 (B, HW).  out[0] = loss (nan when every pixel is ignored, like torch), out[1] = 1/count.
 * Optional: argmax_out int64 (B, HW) = torch.argmax(logits, 1) (scripts/train_SCD.py:243-244);
 * cm long long[C*C] += histogram of (target, argmax) over pixels with 0 <= target < C. */
int c3d_ce2d_fwd(const float* logits, const long long* target, int B, int C, long long HW, long long batch_stride,
                 long long ignore_index, void* ws, float* out, long long* argmax_out, long long* cm,
                 void* cuda_stream);
/* dlogits dense (B, C, HW) = g * (softmax - onehot) / count on counted pixels, 0 on ignored ones. */
int c3d_ce2d_bwd(const float* logits, const long long* target, int B, int C, long long HW, long long batch_stride,
                 long long ignore_index, const float* out, const float* gout, float gscale, float* dlogits,
                 void* cuda_stream);

/* ChangeSimilarity (model/utils.py:180-203): CosineEmbeddingLoss(margin 0, 'mean') between softmax(x1) and
 * softmax(x2) per pixel, target +1 where label_change == 0 and -1 elsewhere.  x1/x2 (B, C, HW) with their own
 * batch strides (the scripts pass `pre_mask[:, 1:]`, scripts/train_SCD.py:228); label_change int64 (B, HW).
 * out[0] = loss.  Backward writes dense dx1, dx2 (B, C, HW). */
int c3d_change_similarity_fwd(const float* x1, const float* x2, const long long* label_change, int B, int C,
                              long long HW, long long batch_stride1, long long batch_stride2, void* ws, float* out,
                              void* cuda_stream);
int c3d_change_similarity_bwd(const float* x1, const float* x2, const long long* label_change, int B, int C,
                              long long HW, long long batch_stride1, long long batch_stride2, const float* gout,
                              float gscale, float* dx1, float* dx2, void* cuda_stream);

/* get_confuse_matrix (utils/metric_tool.py:111-128): cm[num_classes*gt + pred] += 1 over the n entries with
 * 0 <= gt < num_classes (gt float32 when gt_is_float, else int64; pred int64).  num_classes <= 16. */
int c3d_confusion_matrix(const void* gt, int gt_is_float, const long long* pred, long long n, int num_classes,
                         long long* cm, void* cuda_stream);

/* ---- attention core of the captioning head (model/caption_decoder.py:393-423: the scaled-dot-product part of the
 * nn.MultiheadAttention calls; 8 heads x 24 channels, <= 52 target tokens, 256 memory tokens) and of the cached decode
 * step of the caption search (scripts/train_CC.py:209-322).  One CTA per (batch, head).
 * Element (position l, batch b, head h, channel c) of an operand is at ptr + l * ls + b * bs + h * hd + c, so the
 * (L, B, E) projections and slices of a packed (L, B, 3E) in-projection are addressed in place.
 *   S = scale * Q K^T;  causal: key j visible to query i iff j <= i + (Lk - Lq);  P = softmax rows;
 *   dropout: P * keep / (1 - p) with keep (B, nh, Lq, Lk) uint8 (NULL = none);  O = P V.
 * P (B, nh, Lq, Lk), when not NULL, receives the probabilities before dropout (needed by the backward).
 * Limits: Lq <= 64, Lk <= 256, hd <= 64. */
typedef struct c3d_attn_desc {
  const float* q; const float* k; const float* v;
  long long q_ls, q_bs, k_ls, k_bs, v_ls, v_bs;
  float* o;
  long long o_ls, o_bs;
  float* P;
  const unsigned char* keep;
  float keep_scale;
  float scale;
  int B, nh, hd, Lq, Lk, causal;
} c3d_attn_desc;
int c3d_attention_fwd(const c3d_attn_desc* d, void* cuda_stream);
/* dq / dk / dv are written with the strides of q / k / v (dense overwrite of this head's channels). */
int c3d_attention_bwd(const c3d_attn_desc* d, const float* dO, long long do_ls, long long do_bs, float* dq, float* dk,
                      float* dv, void* cuda_stream);

/* Input pipeline (data/transforms.py:82-154,166-206 and the SCD / BDA variants :210-612): normalize -> scale ->
 * random_crop_resize -> random_flip -> random_exchange -> to_tensor of B raw pairs in one launch.
 * img (B, Hs, Ws, 6) uint8 HWC [pre RGB | post RGB] (or float32, already normalised, when img_is_float);
 * label (B, Hs, Ws, L) uint8 or NULL; params int[B][8] = {do_crop, x1, y1, flip_rows, flip_cols, exchange_images,
 * exchange_labels01, 0} (the host draws them in the reference's order).  Outputs: pre / post (B, 3, H, W) float32 NCHW;
 * label_out (B, L, H, W): float32 ceil(l / 255) when label_mode == 0 (BCD), int64 class ids when 1 (SCD / BDA).
 * Bilinear / nearest sampling follow cv2.resize (INTER_LINEAR on the normalised float image, INTER_NEAREST on labels).
 * A crop is only valid when (Hs, Ws) == (H, W) (the reference scales first, then crops the scaled image). */
int c3d_augment_pairs(const void* img, int img_is_float, const unsigned char* label, const int* params, int B, int Hs,
                      int Ws, int H, int W, int L, int label_mode, float mean, float std, float* pre, float* post,
                      void* label_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
