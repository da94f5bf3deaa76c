/*
 * littlegan_b200 - C-ABI of the B200-native LittleGAN hot path.
 *
 * The reference (IXarea/LittleGAN) has no native ABI: every arithmetic op on its hot path is
 * a TensorFlow-1.15 library kernel reached through Keras layers.  Each entry point below
 * therefore cites the reference call site whose TF op it replaces.  Conventions:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every call is asynchronous on `stream`, allocates nothing and transfers no ownership;
 *   - return 0 on success, <0 on error (message via lg_last_error(), thread-local);
 *   - activations are NHWC; `dtype` selects the activation storage type (LG_F32 / LG_BF16);
 *     weights, biases, gradients of weights, statistics and losses are always fp32/fp64.
 *
 * Convolution family.  One 5x5 geometry, stride s in {1,2}, TF 'SAME' padding (pad = 1 for
 * s=2, 2 for s=1), links a "big" map [N,Hb,Wb,A] and a "small" map [N,Hb/s,Wb/s,B] through a
 * kernel W[5,5,A,B] (fp32):
 *     fprop : small[n,i,j,b] = sum_{ky,kx,a} big[n, s*i+ky-pad, s*j+kx-pad, a] * W[ky,kx,a,b]
 *     dgrad : big[n,Y,X,a]   = sum_{s*i+ky-pad==Y, s*j+kx-pad==X, b} small[n,i,j,b] * W[ky,kx,a,b]
 *     wgrad : dW[ky,kx,a,b] += sum_{n,i,j} big[n, s*i+ky-pad, s*j+kx-pad, a] * small[n,i,j,b]
 * tf.layers.Conv2D (kernel HWIO = [5,5,in=A,out=B], model.py:15) is fprop; its input gradient is
 * dgrad.  tf.layers.Conv2DTranspose (kernel [5,5,out=A,in=B], model.py:39,86) is dgrad; its input
 * gradient is fprop.  Both share wgrad with (big, small) = (x, dy) resp. (dy, x).
 */
#ifndef LITTLEGAN_B200_H_
#define LITTLEGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_F32 0
#define LG_BF16 1

#define LG_ACT_NONE 0
#define LG_ACT_TANH 1
#define LG_ACT_SIGMOID 2

#define LG_OK 0
#define LG_ERR_INVALID (-1)
#define LG_ERR_CUDA (-2)
#define LG_ERR_UNSUPPORTED (-3)

/* ABI version of this header (bumped on any signature change). */
int lg_abi_version(void);
/* Last error message of the calling thread ("" if none). */
const char* lg_last_error(void);
/* 1 if the tcgen05/TMA kernels can run on the current device (sm_100), else 0. */
int lg_tensor_core_path_available(void);
/* 1 if the tcgen05 path covers this geometry (op: 0 fprop, 1 dgrad, 2 wgrad), else 0. */
#define LG_OP_FPROP 0
#define LG_OP_DGRAD 1
#define LG_OP_WGRAD 2
int lg_conv2d_tc_supported(int op, int N, int Hb, int Wb, int A, int B, int stride);
/* CTA pairs (tcgen05 cta_group::2: the two SMs of a TPC share one 256-row MMA tile and each fetches half
 * of the weight tile) in the generic conv kernels.  mode 0 = never, 1 = when the launch has enough tiles
 * to fill every TPC (default; LG_TC_PAIRS in the environment sets the initial mode), 2 = whenever the
 * geometry allows (tests).  Returns the previous mode.  Process-wide, not thread-safe. */
int lg_set_cta_pairs(int mode);

/* ---- convolution family (replaces TF Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter,
 *      model.py:15,39,86 and their autodiff from eager_trainer.py:145,149,163) -------------- */

/* Optional fused epilogue of a BACKWARD conv launch (tensor-core path): the launch produces g = dL/da for
 * the layer below, a = LeakyReLU(InstanceNorm(z)) (instance.py:105-128).  With this descriptor the kernel
 * reads the matching element of z, stores dy = g * LeakyReLU'(gamma*xhat + beta) instead of g, and adds
 * (sum dy, sum dy*xhat) per sample into `red` - pass 1 of the norm backward without a separate kernel.
 * Follow with lg_instnorm_act_bwd_apply(..., dy_ready = 1).  lg_conv2d_norm_bwd_supported says whether a
 * geometry accepts it. */
typedef struct lg_norm_bwd {
  const void* z;        /* pre-norm output of the layer below (LG_BF16), shape of this launch's output */
  const double* stats;  /* [N][2] sum / sum of squares of z per sample */
  const float* gamma;   /* [1] */
  const float* beta;    /* [1] */
  double* red;          /* [N][2], ACCUMULATED (caller zeroes) */
  float eps;            /* added to the standard deviation */
  float alpha;          /* LeakyReLU slope */
} lg_norm_bwd_t;
int lg_conv2d_norm_bwd_supported(int op, int N, int Hb, int Wb, int A, int B, int stride);

/* fprop.  bias: [B] or NULL.  stats: double[N][2] (sum, sum of squares of the written values,
 * per sample, ACCUMULATED - caller zeroes) or NULL.  use_tc: 1 = tcgen05 path (requires LG_BF16,
 * wpack from lg_pack_conv_weights), 0 = fp32-accumulate SIMT path reading W fp32.  norm_bwd: NULL or
 * the fused epilogue above (excludes stats). */
int lg_conv2d_fprop(const void* big, const float* W, const void* wpack, const float* bias,
                    void* small_out, double* stats, int N, int Hb, int Wb, int A, int B,
                    int stride, int dtype, int use_tc, const lg_norm_bwd_t* norm_bwd, void* stream);

/* dgrad (== Conv2DTranspose forward).  bias: [A] or NULL; act: LG_ACT_NONE / LG_ACT_TANH applied
 * after the bias; stats as above, taken on the pre-activation values; norm_bwd as above (excludes
 * stats and act). */
int lg_conv2d_dgrad(const void* small, const float* W, const void* wpack, const float* bias,
                    void* big_out, double* stats, int N, int Hb, int Wb, int A, int B,
                    int stride, int act, int dtype, int use_tc, const lg_norm_bwd_t* norm_bwd,
                    void* stream);

/* dgrad on the row-streaming tensor-core kernel (csrc/tc_rowdgrad.cu) for the decoder's last transposed conv
 * (A = 32 output channels, B = 64, stride 2, 128-pixel-wide result; model.py:39): every input row is fetched
 * once into a shared-memory ring and all 25 taps read it in place.  wpack: from lg_pack_rowdgrad_weights
 * (W [5,5,A,B] fp32; call with W or wpack NULL for the size in bytes; refresh after every weight update).
 * bias / stats as for lg_conv2d_dgrad, LG_BF16 only, no activation. */
int lg_conv2d_dgrad_rows_supported(int N, int Hb, int Wb, int A, int B, int stride);
int lg_pack_rowdgrad_weights(const float* W, void* wpack, int A, int B, void* stream);
int lg_conv2d_dgrad_rows(const void* small, const void* wpack, const float* bias, void* big_out, double* stats,
                         int N, int Hb, int Wb, int A, int B, int stride, void* stream);

/* dgrad with an RGB result (A = 3, 128-pixel-wide result: stride 1 / B = 32 = the generator's final
 * Conv2DTranspose, model.py:86; stride 2 / B = 64 = the input gradient of the encoder's first Conv2D, model.py:15) on the row-streaming tensor-core kernel (csrc/tc_rowdeconv.cu).  Same arithmetic as
 * lg_conv2d_dgrad (LG_BF16); big_out_pad8, if not NULL, also receives the image with 8-channel zero-padded
 * pixels [N,Hb,Wb,8] - the layout lg_conv2d_fprop_rows fetches. */
int lg_conv2d_dgrad_rgb_supported(int N, int Hb, int Wb, int A, int B, int stride);
int lg_conv2d_dgrad_rgb(const void* small, const float* W, const float* bias, void* big_out, void* big_out_pad8,
                        double* stats, int N, int Hb, int Wb, int A, int B, int stride, int act, void* stream);

/* wgrad.  dW [5,5,A,B] fp32 is ACCUMULATED into (caller zeroes). */
int lg_conv2d_wgrad(const void* big, const void* small, float* dW, int N, int Hb, int Wb,
                    int A, int B, int stride, int dtype, int use_tc, void* stream);

/* wgrad on the tcgen05 path when `big` is stored with A_big >= A channels (zero padded so that TMA can
 * fetch it: the 3-channel image layers use A_big = 16).  dW stays [5,5,A,B]; bf16 activations only. */
int lg_conv2d_wgrad_padded(const void* big, const void* small, float* dW, int N, int Hb, int Wb,
                           int A_big, int A, int B, int stride, void* stream);

/* fprop on the row-streaming tensor-core kernel (csrc/tc_rowconv.cu) for maps that are 128 pixels wide with
 * few channels: every input row is fetched once into a shared-memory ring and all 25 taps read it in place.
 * `big` is [N,Hb,Wb,A_big] LG_BF16, A_big = 8 for A <= 8 (RGB images zero-padded to 16-byte pixels, see
 * lg_pad_channels) or A_big = A = 32.  wpack: the bf16 weight operand made by lg_pack_rowconv_weights from
 * W[5,5,A,B] (call it with W or wpack NULL to get the size in bytes; refresh after every weight update).
 * Same contract as lg_conv2d_fprop otherwise (bias, stats, norm_bwd).
 * lg_conv2d_fprop_rows_supported: 1 if the geometry is covered. */
int lg_conv2d_fprop_rows_supported(int N, int Hb, int Wb, int A_big, int A, int B, int stride);
int lg_pack_rowconv_weights(const float* W, void* wpack, int A_big, int A, int B, int stride, void* stream);
int lg_conv2d_fprop_rows(const void* big, int A_big, const void* wpack, const float* bias, void* small_out,
                         double* stats, int N, int Hb, int Wb, int A, int B, int stride,
                         const lg_norm_bwd_t* norm_bwd, void* stream);

/* dst[row, 0:Cpad] = (src[row, 0:C], 0, ..., 0): channel padding of an NHWC tensor (rows = N*H*W). */
int lg_pad_channels(const void* src, void* dst, int64_t rows, int C, int Cpad, int dtype,
                    void* stream);

/* Conv2DTranspose spelled out (thin aliases; same arithmetic as above). */
int lg_conv2d_transpose_fprop(const void* x_small, const float* W, const void* wpack,
                              const float* bias, void* y_big, double* stats, int N, int Hb, int Wb,
                              int A_out, int B_in, int stride, int act, int dtype, int use_tc,
                              void* stream);
int lg_conv2d_transpose_dgrad(const void* dy_big, const float* W, const void* wpack,
                              void* dx_small, int N, int Hb, int Wb, int A_out, int B_in,
                              int stride, int dtype, int use_tc, void* stream);

/* bf16 operand copies of one kernel W[25][A][B] for the tcgen05 path:
 *   wpack = [ Wt : 25 x Apad x Bpad (b contiguous) | Wf : 25 x Bpad x Apad (a contiguous) ],
 * Apad = round_up(A,16), Bpad = round_up(B,16), zero padded.  Returns bytes needed when
 * wpack == NULL. */
int64_t lg_pack_conv_weights(const float* W, void* wpack, int A, int B, void* stream);
/* The same for n <= LG_PACK_MAX layers in one launch (arrays of host pointers / sizes). */
#define LG_PACK_MAX 12
int lg_pack_conv_weights_multi(const float* const* W, void* const* wpack, const int* A, const int* B, int n,
                               void* stream);

/* db[c] += sum_rows g[row, c]  (bias gradient; rows = N*H*W).  fp32 accumulate. */
int lg_bias_grad(const void* g, float* db, int64_t rows, int C, int dtype, void* stream);

/* ---- InstanceNormalization(axis=None) + LeakyReLU (instance.py:105-128, model.py:21-24,46-50,
 *      100-102,130-131) ----------------------------------------------------------------------- */

/* Per-sample sum / sum of squares of pre(z) over M elements, pre = leaky(alpha_pre)
 * (alpha_pre = 1 -> identity); stats double[N][2] ACCUMULATED. */
int lg_rowstats(const void* z, double* stats, int N, int64_t M, float alpha_pre, int dtype,
                void* stream);

/* out = leaky_post( gamma * (pre(z) - mu) / (sigma + eps) + beta ) [+ skip]
 * mu, sigma from stats (population variance), gamma/beta: device scalars.  z (and dz below) are
 * z_dtype, out / skip / g are dtype; (z_dtype, dtype) in {(f32,f32), (bf16,bf16), (f32,bf16)} - the
 * last one is the fp32 dense head feeding a bf16 decoder (model.py:98-102,129-132). */
int lg_instnorm_act_fwd(const void* z, const double* stats, const float* gamma, const float* beta,
                        const void* skip, void* out, int N, int64_t M, float eps, float alpha_pre,
                        float alpha_post, int z_dtype, int dtype, void* stream);

/* Backward, pass 1: red double[N][2] += (sum dy, sum dy*xhat), dy = g * leaky_post'(y). */
int lg_instnorm_act_bwd_reduce(const void* g, const void* z, const double* stats,
                               const float* gamma, const float* beta, double* red, int N,
                               int64_t M, float eps, float alpha_pre, float alpha_post, int z_dtype,
                               int dtype, void* stream);
/* Backward, pass 2: dz = pre'(z) * (gamma/s) * (dy - mean(dy) - xhat*(s/sigma)*mean(dy*xhat));
 * dgamma += sum_n red[n][1], dbeta += sum_n red[n][0] (added once, by block 0), either may be
 * NULL.  dy_ready: `g` already holds dy (written by a conv launch with a lg_norm_bwd_t epilogue).
 * dbias (or NULL): float[C] += per-channel sums of dz (the bias gradient of the conv that produced z,
 * channels innermost); fused only when lg_instnorm_bias_grad_fusable(M, C, z_dtype), else
 * LG_ERR_UNSUPPORTED (use lg_bias_grad). */
int lg_instnorm_bias_grad_fusable(int64_t M, int C, int z_dtype);
int lg_instnorm_act_bwd_apply(const void* g, const void* z, const double* stats, const double* red,
                              const float* gamma, const float* beta, void* dz, float* dgamma,
                              float* dbeta, float* dbias, int C, int dy_ready, int N, int64_t M,
                              float eps, float alpha_pre, float alpha_post, int z_dtype, int dtype,
                              void* stream);

/* ---- Dense (tf.layers.Dense, model.py:62,63,83,120): C[M,N] (+)= op(A)[M,K] * op(B)[K,N] ----
 * A is activation-typed (a_dtype), B fp32, C c_dtype.  transA: A stored [K,M]; transB: B stored
 * [N,K].  accumulate: C += (fp32 C only; also used for split-K, caller zeroes).  bias [N] (or NULL)
 * is added in the epilogue when accumulate == 0. */
int lg_gemm(const void* A, const float* Bm, const float* bias, void* C, int M, int N, int K,
            int transA, int transB, int accumulate, int a_dtype, int c_dtype, void* stream);
/* Two Dense layers sharing one input - the discriminator heads (model.py:62-63,70-72):
 *   out_h[N,U_h] = act(feat[N,F] @ W_h[F,U_h] + b_h),  h = 0,1;  feat is `dtype`, everything else fp32.
 * One pass over feat.  Needs U0 + U1 <= 48 and F % 64 == 0 (else LG_ERR_UNSUPPORTED: use lg_gemm).
 * `workspace`: lg_dense_heads_workspace_bytes(N) bytes, zero-filled ONCE by the caller; every call leaves it
 * zeroed again (calls sharing a workspace must be stream-ordered). */
int64_t lg_dense_heads_workspace_bytes(int N);
int lg_dense_heads_fwd(const void* feat, const float* W0, const float* b0, int U0, const float* W1,
                       const float* b1, int U1, float* out0, float* out1, int N, int F, int act, int dtype,
                       void* workspace, void* stream);
/* Backward of the pair given dl_h = d loss / d(pre-activation) [N,U_h] fp32 (one of them may be NULL = 0):
 *   dfeat[N,F] (`dtype`, written; may be NULL) = dl0 @ W0^T + dl1 @ W1^T
 *   dW_h[F,U_h] += feat^T @ dl_h ; db_h[U_h] += column sums of dl_h      (each may be NULL = skip) */
int lg_dense_heads_bwd(const void* feat, const float* dl0, const float* dl1, const float* W0, int U0,
                       const float* W1, int U1, void* dfeat, float* dW0, float* dW1, float* db0, float* db1,
                       int N, int F, int dtype, void* stream);
/* x[r,c] = act(x[r,c] + bias[c]) in place, fp32. */
int lg_bias_act(float* x, const float* bias, int rows, int cols, int act, void* stream);

/* ---- losses (eager_trainer.py:85-102, Keras binary_crossentropy, utils.py:47-48) ----------- */

/* p [rows,cols] probabilities (post-sigmoid).  target: device [rows,cols] or NULL -> constant
 * target_const.  loss_accum[0] += weight * mean(bce).  dlogit (may be NULL) =
 * weight/(rows*cols) * dBCE/dp * p*(1-p)   (gradient w.r.t. the pre-sigmoid logits). */
int lg_bce_sigmoid(const float* p, const float* target, float target_const, int rows, int cols,
                   float weight, float* loss_accum, float* dlogit, void* stream);
/* n <= LG_BCE_MAX such terms in one launch (same arithmetic per item: p [n] probabilities, target tensor or constant,
 * weight, loss accumulator (+=), optional d(loss)/d(logit) output). */
#define LG_BCE_MAX 8
typedef struct lg_bce_item {
  const float* p;
  const float* target;   /* NULL: use target_const */
  float* loss_accum;     /* may be NULL */
  float* dlogit;         /* may be NULL */
  float target_const;
  float weight;
  int n;                 /* rows * cols */
  int pad_;
} lg_bce_item_t;
int lg_bce_sigmoid_multi(const lg_bce_item_t* items, int n, void* stream);

/* y = tanh output image, t = target image (both activation-typed, n elements).
 * loss_accum[0] += weight * mean|t - y|  (if loss_accum != NULL);
 * dpre (may be NULL) = (g_in + weight/n * sign(y - t)) * (1 - y^2), g_in may be NULL. */
int lg_l1_tanh_bwd(const void* y, const void* t, const void* g_in, void* dpre, int64_t n,
                   float weight, float* loss_accum, int dtype, void* stream);

/* ---- optimiser (tf.compat.v1.train.AdamOptimizer, eager_trainer.py:28-30,164-168) ---------- */

/* state: double[4] = {t, beta1^t, beta2^t, lr_t}; advances t by one and recomputes
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t).  Called once per apply_gradients. */
int lg_adam_advance(double* state, double lr, double beta1, double beta2, void* stream);
/* TF Adam on a flat range: optional value clip of g to [-clip, clip] (clip <= 0: none),
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps). */
int lg_adam_apply(float* p, const float* g, float* m, float* v, int64_t n, const double* state,
                  float beta1, float beta2, float eps, float clip, void* stream);

/* ---- input augmentation of the train step (eager_trainer.py:127-131: tf.image.random_flip_left_right,
 *      random_brightness(0.02), random_contrast(0.75, 1.003), random_hue(0.03) and + 0.1 * N(0, 0.2)) -----------
 * x: [N,H,W,3] fp32 (H*W and W multiples of 4).  params: float[4 + 4 N] = { brightness delta, contrast factor, hue
 * delta (turns), step ; per image: mean_r, mean_g, mean_b, flip (0/1) }.  state: uint64[2] = { seed, step } on the
 * device.  lg_augment_prepare writes the per-image channel means and, with draw != 0, this step's draws from
 * Philox4x32-10(seed, step) (draw == 0: the caller has filled the three scalars and the flips).
 * lg_augment_apply writes out = hue(contrast(brightness(flip(x)))) + noise in out_dtype; noise: a [N,H,W,3] fp32
 * tensor added as is, or NULL = N(0, noise_std) from Philox(seed, step), in which case it also advances state[1]. */
int lg_augment_prepare(const float* x, int N, int H, int W, void* state, float* params, float max_brightness,
                       float contrast_lo, float contrast_hi, float max_hue, int draw, void* stream);
int lg_augment_apply(const float* x, const float* params, const float* noise, void* state, float noise_std,
                     void* out, int N, int H, int W, int out_dtype, void* stream);

/* ---- casts ------------------------------------------------------------------------------- */
int lg_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);
/* dst = src * scale + shift in fp32 arithmetic (dtypes LG_F32 / LG_BF16 each side): the in-step copies and the
 * adjuster's input condition (cond + 1) * 0.5 (eager_trainer.py:155). */
int lg_scale_shift(const void* src, void* dst, int64_t n, float scale, float shift, int src_dtype, int dst_dtype,
                   void* stream);
/* out[0..n) ~ N(0, 1) fp32 (the generator noise, eager_trainer.py:125 tf.random.normal) from Philox4x32-10.
 * state: device uint64[3] = {seed, step, 0}; the launch advances `step` on the device (CUDA-graph replayable);
 * out 16-byte aligned. */
int lg_normal_fill(float* out, int64_t n, void* state, void* stream);
/* data_rescale (utils.py:51-52, applied by dataset.py:29 to every decoded image): dst[i] = src[i] / 127.5 - 1 for
 * n decoded image bytes; dst fp32 or bf16 (dst_dtype).  Both pointers 16-byte aligned.  Replaces the host-side
 * tf.cast + tf.divide + tf.subtract of the input pipeline: the batch crosses PCIe as bytes. */
int lg_u8_rescale(const void* src, void* dst, int64_t n, int dst_dtype, void* stream);

/* ---- Inception-2015 pool_3 forward (fid.py:36-106: create_inception_graph / get_activations) -- */

/* One conv + folded batch-norm (+ ReLU) unit of the graph, NHWC.  x: N x H x W pixels of x_stride channels, the unit
 * reads channels [x_off, x_off + Cin).  W fp32 [kh,kw,Cin,Cout] (TF HWIO).  y: N x Ho x Wo pixels of y_stride
 * channels, the unit writes channels [y_off, y_off + Cout) - the slice of the block's concat output - with
 * Ho = (H + 2 ph - kh) / stride + 1.  y = relu(scale[co] * conv + shift[co]); scale = gamma / sqrt(var + eps),
 * shift = beta - mean * scale.  dtype: storage type of x and y (LG_F32 / LG_BF16), accumulation fp32. */
int lg_conv2d_bn_relu(const void* x, const float* W, const void* wpack, const float* scale, const float* shift, void* y,
                      int N, int H, int Wd, int Cin, int x_stride, int x_off, int kh, int kw, int stride, int ph, int pw,
                      int Cout, int y_stride, int y_off, int relu, int dtype, void* stream);
/* bf16 tensor-core operand of a unit's kernel (W fp32 [kh,kw,Cin,Cout] -> wpack), made once per weight set.  With W or
 * wpack NULL: the size of wpack in bytes, 0 when the geometry has no tensor-core form (Cin % 8 or Cout % 16 != 0; such
 * units run the SIMT path of lg_conv2d_bn_relu from W).  lg_conv2d_bn_relu takes the tcgen05 path when dtype is
 * LG_BF16, wpack is given and the channel strides / offsets are multiples of 8; otherwise W is required. */
int lg_pack_conv_bn_weights(const float* W, void* wpack, int Cin, int kh, int kw, int Cout, void* stream);
/* k x k pooling with zero padding `pad` into a channel slice.  mode 0: max; 1: mean over the in-bounds taps (TF
 * AvgPool SAME); 2: mean with the padding counted. */
int lg_pool2d(const void* x, void* y, int N, int H, int W, int C, int x_stride, int x_off, int k, int stride, int pad,
              int mode, int y_stride, int y_off, int dtype, void* stream);
/* pool_3: y fp32 [N,C] = mean over the HW positions of x [N,HW,C]. */
int lg_global_avgpool(const void* x, float* y, int N, int HW, int C, int dtype, void* stream);
/* The graph's input stage: TF-1.x ResizeBilinear (align_corners = false) of x [N,H,W,C] (uint8 if src_is_u8, else
 * fp32, values 0..255) to y [N,Ho,Wo,Cpad], then (v - sub) * mul; channels [C, Cpad) of y are zero (Cpad = 8 makes an
 * RGB pixel one 16-byte run, the form the tensor-core conv unit reads). */
int lg_resize_bilinear_norm(const void* x, void* y, int N, int H, int W, int C, int Cpad, int Ho, int Wo, float sub,
                            float mul, int src_is_u8, int dtype, void* stream);

/* ---- FID statistics (fid.py:169-188: np.mean / np.cov in fp64) ----------------------------- */

/* X [n,d] fp32 features.  S1 double[d] += sum_r (x_r - shift); S2 double[d,d], UPPER triangle
 * only (j >= i) += sum_r (x_r - shift)(x_r - shift)^T.  shift double[d] (may be NULL = 0). */
int lg_fid_accumulate(const float* X, const double* shift, double* S1, double* S2, int64_t n,
                      int d, void* stream);
/* mu = shift + S1/n ; sigma = (S2 - S1 S1^T / n) / (n-1)  (unbiased, np.cov default); reads the upper
 * triangle of S2 and writes the full, exactly symmetric sigma. */
int lg_fid_finalize(const double* S1, const double* S2, const double* shift, double* mu,
                    double* sigma, int64_t n, int d, void* stream);

/* ---- Frechet distance (fid.py:112-163: scipy.linalg.sqrtm of sigma1 sigma2 in fp64) ------------------------
 * Tr sqrtm(sigma1 sigma2) = Tr (R sigma2 R)^(1/2) with R = sigma1^(1/2) (same spectrum, symmetric PSD), both
 * square roots by the coupled Newton-Schulz iteration - fp64 matrix products only.  The host side
 * (littlegan_b200/fid.py) drives the iteration through these three entry points; all matrices are row-major
 * double[n,n] in device memory. */

/* C = alpha * A @ B + diag * I.  C must not alias A or B. */
int lg_dgemm(const double* A, const double* B, double* C, int n, double alpha, double diag, void* stream);
/* out3[0] = trace(A), out3[1] = ||A||_F^2, out3[2] = max_ij |A_ij - A_ji|  (device doubles, written by the launch). */
int lg_dmat_stats(const double* A, int n, double* out3, void* stream);
/* dst = alpha * S + diag * I with S = src, or (src + src^T) / 2 when `symmetrise` (then dst != src), or 0 when
 * src is NULL. */
int lg_dmat_scale_shift(const double* src, double* dst, int n, double alpha, double diag, int symmetrise,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LITTLEGAN_B200_H_ */
