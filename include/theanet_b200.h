/* theanet_b200 -- C ABI of the B200 (sm_100a) kernels behind theanet's CNN-training hot path.
 *
 * The reference (rakeshvar/theanet) has no FFI: its "operator API" is the set of Theano ops each
 * layer asks for.  Every entry point below replaces one such call site (cited as
 * <reference file>:<line>) and is what theanet_b200/_C.py binds with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative TN_ERR_* code; tn_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - all pointers are DEVICE pointers owned by the caller unless the name ends in _host; the
 *     library never allocates or frees caller memory and never synchronises;
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous and legal inside
 *     CUDA-graph capture;
 *   - activations are NCHW float32, conv filters OIHW, dense weights (n_in, n_out) row-major --
 *     the layouts theanet's .pkl exposes (theanet/neuralnet.py:298-301);
 *   - `ctl` is a device int32[TN_CTL_WORDS] control block the host refreshes once per step (so a
 *     captured graph can be replayed): see TN_CTL_*.
 */
#ifndef THEANET_B200_H
#define THEANET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TN_VERSION 100

enum {
  TN_OK = 0,
  TN_ERR_ARG = -1,
  TN_ERR_SHAPE = -2,
  TN_ERR_ALIGN = -3,
  TN_ERR_CUDA = -4,
  TN_ERR_UNSUPPORTED = -5
};

/* activation kinds -- theanet/layer/layer.py:27-39 */
enum {
  TN_ACT_LINEAR = 0,
  TN_ACT_RELU = 1,        /* max(0,x) */
  TN_ACT_LEAKY = 2,       /* reluNN: max(0,x) + (min(0,x)*NN)/100, NN in act_nn */
  TN_ACT_TANH = 3,
  TN_ACT_SCALED_TANH = 4, /* 1.7*tanh(2x/3) */
  TN_ACT_SIGMOID = 5,
  TN_ACT_SOFTPLUS = 6
};

/* control block words (device int32[TN_CTL_WORDS]) */
enum {
  TN_CTL_STEP = 0,     /* Philox step counter */
  TN_CTL_SAMPLE0 = 1,  /* global index (within the minibatch stream) of this shard's first sample */
  TN_CTL_ROW0 = 2,     /* first corpus row of this shard when slicing x[i*B:(i+1)*B] */
  TN_CTL_LR_BITS = 3,  /* float32 learning rate, bit pattern */
  TN_CTL_WORDS = 8
};

/* Philox stream purposes (counter word 3) -- see DESIGN.md "Randomness" */
enum { TN_RNG_DROPOUT = 0, TN_RNG_FLIP = 1, TN_RNG_NOISE = 2, TN_RNG_SCALARS = 3, TN_RNG_COLOR = 4, TN_RNG_AUX = 5 };

int tn_version(void);
const char *tn_last_error(void);
/* number of kernels this library has launched (or recorded into a graph capture) so far */
uint64_t tn_launch_count(void);
/* refuses anything that is not compute capability 10.x */
int tn_device_check(int device);
/* Write the per-step words of the control block (TN_CTL_STEP, _SAMPLE0, _ROW0, _LR_BITS) on
 * `stream`.  The values are kernel arguments, i.e. captured when the call is enqueued: the host may
 * run ahead of the device without racing a step that is still queued. */
int tn_set_ctl(int32_t *ctl, int step, int sample0, int row0, int lr_bits, void *stream);

/* ---- randomness (replaces theano RandomStreams: dropout.py:10-12, inlayers.py:72-141) -------- */
/* words[s*n + j] = Philox4x32-10 word j of sample (sample0+s); used by tests to pin the stream */
int tn_philox_words(uint32_t *words, int n_samples, int n_per_sample, uint64_t seed, int purpose,
                    int step, int sample0, void *stream);

/* ---- ElasticLayer (theanet/layer/inlayers.py:29-163) -------------------------------------- */
typedef struct tn_elastic_prm {
  int h;              /* img_sz (square) */
  int sigma;          /* gaussian half-width; table is (2*sigma+1)^2 */
  float translation;  /* 0 = off */
  float magnitude;    /* 0 = off */
  float log_zoom;     /* float32(ln zoom); 0 = off */
  float angle_rad;    /* float32(angle*pi/180); 0 = off */
  int zoom_on;        /* zoom != 1 */
  int nearest;        /* 1: iround gather (inlayers.py:124-127), 0: bilinear (:129-137) */
  double clip_hi;     /* h - 1 - .001 (inlayers.py:121-122) */
  int step_offset;    /* the draws are those of step ctl[TN_CTL_STEP] + step_offset: 1 computes the
                         NEXT step's field while the current step is still running */
  int reserved;
} tn_elastic_prm;

/* noise[2*h*h] ~ N(0,1) float32, Box-Muller on the (seed, step) Philox stream (inlayers.py:94) */
int tn_elastic_noise(float *noise, int h, uint64_t seed, const int32_t *ctl, void *stream);
/* The per-minibatch sampling grid (inlayers.py:77-122).  u_inj: 8 injected uniforms or NULL (then
 * drawn from the (seed, step) stream).  noise: the N(0,1) field, or NULL to draw it in-kernel
 * from the same stream tn_elastic_noise uses.  filt: the (2*sigma+1)^2 table (inlayers.py:87-91).
 * Outputs: target[2*h*h] float64 before clipping (debugout, may be NULL), tyx[2*h*h] float64
 * clipped coordinates (may be NULL), gidx[h*h] int32 gather base (row*h+col), gfrac[2*h*h]
 * float32 (fy, fx) -- written only when !nearest. */
int tn_elastic_field(const tn_elastic_prm *prm_host, const float *noise, const float *u_inj,
                     const float *filt, uint64_t seed, const int32_t *ctl, double *target,
                     double *tyx, int32_t *gidx, float *gfrac, void *stream);
/* Batch loader + warp (inlayers.py:63-64,124-142): out[b] = flip(gather(invert(src[row(b)]))).
 * row(b) = idx ? idx[b] : ctl[ROW0] + b.  mode 0 = no gather (identity grid), 1 = nearest,
 * 2 = bilinear.  pflip > 0 draws the per-pixel Bernoulli mask from the Philox stream unless
 * flip_inj (B*C*h*h float32 0/1) is given. */
int tn_elastic_warp(const float *corpus, const int32_t *idx, const int32_t *ctl, int B, int C,
                    int h, int invert, int mode, const int32_t *gidx, const float *gfrac,
                    double pflip, const float *flip_inj, uint64_t seed, float *out, void *stream);

/* ---- ConvLayer (theanet/layer/convpool.py:14-95; nnet.conv2d :54-56, filter flipped) -------- */
/* out[b,m,i,j] = act(bias[m] + sum_{c,u,v} xpad[b,c,i+u,j+v] * W[m,c,f-1-u,f-1-v]); stride 1;
 * pad_lo zeros before, whatever is needed after (valid: 0, same: f-1-(f-1)/2). */
int tn_conv2d_fprop(const float *x, const float *W, const float *bias, float *out, int B, int C,
                    int S, int M, int f, int pad_lo, int out_sz, int act, int act_nn,
                    void *stream);
/* dx[b,c,y,x] = sum_{m,u,v} gz[...] * W[...] (gradient wrt the layer input; tt.grad layer.py:83).
 * If x_in != NULL the result is multiplied by act'(x_in) of the PREVIOUS layer's activation
 * (act_prev / nn_prev), i.e. it directly yields dL/dz of a conv layer feeding this one. */
int tn_conv2d_dgrad(const float *gz, const float *W, float *dx, const float *x_in, int B, int C,
                    int S, int M, int f, int pad_lo, int out_sz, int act_prev, int nn_prev,
                    void *stream);
size_t tn_conv2d_wgrad_workspace_bytes(int B, int C, int S, int M, int f);
/* dW (OIHW) and db from the layer input x and dL/dz; deterministic two-stage reduction through
 * `workspace` (>= tn_conv2d_wgrad_workspace_bytes). */
int tn_conv2d_wgrad(const float *x, const float *gz, float *dW, float *db, void *workspace, int B,
                    int C, int S, int M, int f, int pad_lo, int out_sz, void *stream);

/* ---- strided ConvLayer (convpool.py:54-56,69-70): nnet.conv2d(subsample=(s,s)) is the stride-1
 * convolution sampled on the lattice (i*s, j*s).  out[p,i,j] = x[p,i*s,j*s]; the gradient puts
 * g[p,i,j] back on the lattice and zeros elsewhere, after which the stride-1 wgrad / dgrad apply. */
int tn_subsample2d(const float *x, float *out, int planes, int S, int stride, int out_sz,
                   void *stream);
int tn_upsample2d_zero(const float *g, float *out, int planes, int S, int stride, int out_sz,
                       void *stream);

/* ---- MeanLayer (theanet/layer/convpool.py:129-144): out[plane] = mean over the S x S map --------- */
int tn_meanpool_fwd(const float *x, float *out, int planes, int S, void *stream);
/* dx[plane,i] = dout[plane] / (S*S) * act'(x[plane,i]); x = output of the layer below (NULL or
 * act = linear: no activation derivative), so it yields dL/dz of a ConvLayer below directly. */
int tn_meanpool_bwd(const float *dout, const float *x, float *dx, int planes, int S, int act,
                    int act_nn, void *stream);

/* ---- ColorLayer (theanet/layer/color.py:9-52): random white balance + two gamma curves per
 * (sample, map): out = maxval * (1 - (1 - clip(x/maxval * e1, 0, 1)^e2)^e3), e1 = exp(log_balance*u1),
 * e2 = exp(log_gamma*u2), e3 = exp(log_gamma*u3), u ~ U(-1,1) from the (seed, step, global sample)
 * stream (block = map) or from u_inj (B*C*3 float32).  log_* = float32(ln balance / gamma). */
int tn_color_jitter(const float *x, float *out, int B, int C, int S, float log_balance,
                    float log_gamma, float maxval, uint64_t seed, const int32_t *ctl,
                    const float *u_inj, void *stream);

/* ---- auxiliary-input layers (theanet/layer/auxiliary.py:14-160) ------------------------------
 * LocationInfo input: aux (B,2,2) -> loc (B,2).  train != 0: aux[b,0,:]*u + aux[b,1,:]*(1-u) with
 * u ~ U(0,1) per sample from the (seed, step, global sample) stream or u_inj (B) (:25-28);
 * train == 0: mean over axis 1 (:31).  Both times boost (:33).  The two small dense layers that
 * follow (:36-53) and every gradient use tn_dense_*. */
int tn_aux_location_mix(const float *aux, float *loc, int B, float boost, int train, uint64_t seed,
                        const int32_t *ctl, const float *u_inj, void *stream);
/* out (B, na+nc) = [a (B,na) | c (B,nc)]  (AuxConcatLayer, :80) */
int tn_concat_cols(const float *a, int na, const float *c, int nc, float *out, int B, void *stream);
/* out (B,n) = src[:, off:off+n] of a (B,stride) matrix (gradient of the concatenation) */
int tn_slice_cols(const float *src, int stride, int off, int n, float *out, int B, void *stream);
/* dst[i] += src[i]  (SoftAuxLayer scores: hidden + cross term, :133-134) */
int tn_add_inplace(float *dst, const float *src, int64_t n, void *stream);

/* ---- ConvLayer + PoolLayer fused (small channel counts: theanet's shipped networks) ----------
 * One image per CTA stays resident in shared memory; f must be 3 or 5.  pool = 0: no PoolLayer
 * follows (pooled / pool_out_sz ignored; dtop is dL/da).  Replaces nnet.conv2d + pool_2d
 * (convpool.py:54-56,106-107) and their gradients (tt.grad, layer.py:83). */
/* a = act(conv(x) + b) (B,M,out,out); pooled = maxpool(a) (B,M,pool_out,pool_out) */
int tn_convpool_fprop(const float *x, const float *W, const float *bias, float *a, float *pooled,
                      int B, int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                      int act_nn, int pool, int pool_out_sz, void *stream);
size_t tn_convpool_bwd_weights_workspace_bytes(int B, int C, int M, int f);
/* dW, db from (x, a, pooled, dtop = dL/dpooled): dL/dz = [a == pooled(window)] * dtop * act'(a)
 * is rebuilt in shared memory (every tied maximum receives the gradient) */
int tn_convpool_bwd_weights(const float *x, const float *a, const float *pooled,
                            const float *dtop, float *dW, float *db, void *workspace, int B,
                            int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                            int act_nn, int pool, int pool_out_sz, void *stream);
/* dx = dL/d(layer input); if below != NULL (the output of a ConvLayer feeding this one directly)
 * dx is multiplied by that layer's act' */
int tn_convpool_bwd_data(const float *a, const float *pooled, const float *dtop, const float *W,
                         float *dx, const float *below, int B, int C, int S, int M, int f,
                         int pad_lo, int out_sz, int act, int act_nn, int pool, int pool_out_sz,
                         int act_below, int nn_below, void *stream);
/* Second-generation path for the shipped geometry (filter 3x3, mode 'valid', pool 2, ReLU-family
 * or linear activation; conv_small.cu): groups of images per CTA, and ONE backward launch that
 * rebuilds dL/dz once and produces dW, db and (dx != NULL) dL/d(layer input), with the cross-CTA
 * sum of the weight gradient folded in (two-level ticket, fixed order: deterministic).
 * tn_convpool_fprop takes this path by itself when tn_convpool_small_supported() says so.
 * tn_convpool_fprop_train is the training-time forward: `a` may be NULL (the un-pooled activations
 * are then never written) and `tie` (B*M*pool_out^2 bytes), if given, records per pooled cell which
 * window elements equal the maximum (bit 2*dy+dx) -- Theano's tie-duplicating MaxPoolGrad needs
 * exactly that.  tn_convpool_bwd takes either `tie` or `a` (then it compares a with pooled).
 * `workspace` (>= tn_convpool_bwd_workspace_bytes) must be zero-filled once before its first
 * use; the kernel leaves its ticket counters at zero. */
int tn_convpool_small_supported(int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                                int pool, int pool_out_sz);
int tn_convpool_fprop_train(const float *x, const float *W, const float *bias, float *a,
                            float *pooled, uint8_t *tie, int B, int C, int S, int M, int f,
                            int pad_lo, int out_sz, int act, int act_nn, int pool, int pool_out_sz,
                            void *stream);
size_t tn_convpool_bwd_workspace_bytes(int B, int C, int S, int M, int f, int pad_lo, int out_sz,
                                       int act, int pool, int pool_out_sz, int need_dx);
int tn_convpool_bwd(const float *x, const float *a, const uint8_t *tie, const float *pooled,
                    const float *dtop, const float *W, float *dW, float *db, float *dx,
                    const float *below, void *workspace, int B, int C, int S, int M, int f,
                    int pad_lo, int out_sz, int act, int act_nn, int pool, int pool_out_sz,
                    int act_below, int nn_below, void *stream);

/* Debug aid (tools/phase_times.py): when buf != NULL, thread 0 of every CTA of tn_convpool_bwd
 * stores clock64() at its phase boundaries into buf[cta * 64 + slot] (grid <= 1024 CTAs). */
int tn_convpool_debug_timestamps(long long *buf);

/* ---- ConvLayer on tcgen05 tensor cores: bf16 implicit GEMM, NHWC activations (conv_tc.cu) -----
 * For wide layers (C % 64 == 0, M % 64 == 0, mode 'same', output width a power of two <= 128):
 * config C4 of BASELINE.json.  Activation tensors are NHWC bfloat16 (void*); filters are repacked
 * from the float32 OIHW master copy with tn_conv2d_tc_pack_weights (dgrad = 0: [M][f*f][C],
 * flipped, for fprop; dgrad = 1: [C][f*f][M] for dgrad).  Accumulation is float32. */
int tn_conv2d_tc_supported(int C, int S, int M, int f, int out_sz);
int tn_nchw_f32_to_nhwc_bf16(const float *x, void *y, int B, int C, int H, int W, void *stream);
int tn_nhwc_bf16_to_nchw_f32(const void *x, float *y, int B, int C, int H, int W, void *stream);
int tn_conv2d_tc_pack_weights(const float *W, void *Wp, int M, int C, int f, int dgrad,
                              void *stream);
/* a = act(conv(x) + b) (B,out,out,M) bf16; pooled (may be NULL) = 2x2 max-pool of a, fused
 * (needs an even out_sz <= 16); act must be linear / relu / reluNN */
int tn_conv2d_tc_fprop(const void *x, const void *Wp, const float *bias, void *a, void *pooled,
                       int B, int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                       int act_nn, void *stream);
/* dx (B,S,S,C) bf16 from gz (B,out,out,M) bf16 */
int tn_conv2d_tc_dgrad(const void *gz, const void *Wp_dgrad, void *dx, int B, int C, int S, int M,
                       int f, int pad_lo, int out_sz, void *stream);
/* First layer with few input channels (C*f*f <= 64): xcol (B,S,S,64) bf16 with
 * xcol[..,k] = xpad[c, y+u-pad, x+v-pad], k = (c*f+u)*f+v; the layer then runs as a 1x1
 * tensor-core convolution over 64 "channels" (tn_conv2d_tc_fprop / _wgrad with f = 1, C = 64)
 * with filters packed by tn_conv2d_tc_pack_weights_im2col ([M][64]) and the float32 [M][64]
 * weight gradient scattered back to OIHW by tn_conv2d_tc_unpack_wgrad_im2col. */
int tn_im2col_bf16(const float *x, void *xcol, int B, int C, int S, int f, int pad_lo, void *stream);
int tn_conv2d_tc_pack_weights_im2col(const float *W, void *Wcol, int M, int C, int f, void *stream);
int tn_conv2d_tc_unpack_wgrad_im2col(const float *dWcol, float *dW, int M, int C, int f,
                                     void *stream);
/* 2x2 max-pool of an NHWC bf16 tensor (B,S,S,M) -> (B,S/2,S/2,M) */
int tn_maxpool2_nhwc_bf16(const void *a, void *pooled, int B, int S, int M, void *stream);
/* gz = [a == pooled(2x2 window)] * dtop * act'(a), NHWC bf16 (every tied maximum receives the
 * gradient).  dtop: float32 NCHW (B,M,S/2,S/2) if dtop_nchw_f32 else bf16 NHWC; pooled == NULL:
 * no pool layer, dtop has the shape of a. */
int tn_poolbwd_nhwc_bf16(const void *a, const void *pooled, const void *dtop, int dtop_nchw_f32,
                         void *gz, int B, int S, int M, int act, int act_nn, void *stream);
size_t tn_conv2d_tc_wgrad_workspace_bytes(int B, int C, int M, int f, int out_sz);
/* dW (OIHW float32), db from x (B,S,S,C) and gz (B,out,out,M), both bf16; split-K partials in
 * `workspace`, reduced in a fixed order (deterministic) */
int tn_conv2d_tc_wgrad(const void *x, const void *gz, float *dW, float *db, void *workspace, int B,
                       int C, int S, int M, int f, int pad_lo, int out_sz, void *stream);

/* ---- PoolLayer (theanet/layer/convpool.py:97-127; pool_2d max, stride = window) ------------- */
/* planes = B*C; out_sz = ceil(S/p) (ignore_border=False) or S/p */
int tn_maxpool_fwd(const float *x, float *out, int planes, int S, int p, int out_sz,
                   void *stream);
/* dx = (x == out[window]) ? dout[window] : 0  -- every tied maximum receives the gradient
 * (Theano MaxPoolGrad) -- then times act'(x) of the layer that produced x (fused dL/dz). */
int tn_maxpool_bwd(const float *dout, const float *x, const float *out, float *dx, int planes,
                   int S, int p, int out_sz, int act, int act_nn, void *stream);

/* ---- HiddenLayer / DropOutLayer (theanet/layer/hidden.py:11-54, dropout.py:9-31) ------------ */
/* out = act(x.W + b) * mask * out_scale.  mask ~ Bernoulli(pkeep) per (global sample, unit) from
 * the (seed, step) stream (pkeep >= 1: none); mask_inj (B*n_out float32) overrides the stream.
 * out_scale = 1-pdrop for the test twin (hidden.py:50-55), 1 otherwise. */
int tn_dense_fwd(const float *x, const float *W, const float *bias, float *out, int B, int n_in,
                 int n_out, int act, int act_nn, double pkeep, uint64_t seed, const int32_t *ctl,
                 const float *mask_inj, float out_scale, void *stream);
/* dx = g.W^T; if prev_out != NULL it is turned into dL/dz of the previous dense layer:
 * dx *= mask_prev * act_prev'(prev_out) (prev_out = that layer's stored, masked output). */
int tn_dense_bwd_data(const float *g, const float *W, float *dx, int B, int n_in, int n_out,
                      const float *prev_out, int act_prev, int nn_prev, double pkeep_prev,
                      uint64_t seed_prev, const int32_t *ctl, const float *mask_inj_prev,
                      void *stream);
/* dW = x^T.g, db = column sums of g */
int tn_dense_bwd_weights(const float *x, const float *g, float *dW, float *db, int B, int n_in,
                         int n_out, void *stream);
/* The same two products with a bound on the SMs (= CTAs of the cluster split-K tensor-core kernel)
 * each may occupy, 0 = all: the two gradients of a dense layer are independent, and a caller that
 * launches them on two streams with max_sms summing to <= 148 gets them side by side instead of
 * two kernels that each want the whole GPU and take turns (NeuralNet: 1/3 for dW, 2/3 for dX). */
int tn_dense_bwd_data_sm(const float *g, const float *W, float *dx, int B, int n_in, int n_out,
                         const float *prev_out, int act_prev, int nn_prev, double pkeep_prev,
                         uint64_t seed_prev, const int32_t *ctl, const float *mask_inj_prev,
                         int max_sms, void *stream);
int tn_dense_bwd_weights_sm(const float *x, const float *g, float *dW, float *db, int B, int n_in,
                            int n_out, int max_sms, void *stream);
/* Dense-path selector (process-wide debugging / benchmarking knob): 0 = auto (tcgen05 tensor cores
 * with 3xTF32 error compensation whenever n_in, n_out are multiples of 4 and n_out > 32, CUDA-core
 * kernels otherwise; the K range of a product is split over a thread-block cluster whose CTAs
 * exchange their partial tiles through distributed shared memory, fixed order: deterministic),
 * 1 = CUDA cores only, 2 = tensor cores with a single TF32 pass, 3 = as 0, 4 = 3xTF32 on the
 * first-generation kernel (one CTA per tile, fp32 promotion of every 64-deep k-block). */
int tn_set_dense_mode(int mode);
/* Debug aid (tools/gemm_phase_times.py): when buf != NULL, every CTA of the cluster split-K kernel
 * stores clock64() at its phase boundaries into buf[cta * 16 + slot]. */
int tn_dense_debug_timestamps(long long *buf);
/* standalone dropout / test-time scaling: out = x * mask * scale (also used on gradients) */
int tn_dropout_apply(const float *x, float *out, int B, int n, double pkeep, uint64_t seed,
                     const int32_t *ctl, const float *mask_inj, float scale, void *stream);
/* dL/dz = g * act'(a) for a layer whose stored output is a (conv directly followed by dense) */
int tn_act_bwd(const float *g, const float *a, float *gz, int64_t n, int act, int act_nn,
               void *stream);

/* ---- SoftmaxLayer + NLL (theanet/layer/outlayers.py:50-51,69-80,83-102) ---------------------- */
/* labels: y[row(b)] with row(b) as in tn_elastic_warp.  logprob = log softmax(z);
 * rowloss[b] = -logprob[b,y_b]; g = (softmax - onehot) * inv_global_batch. */
int tn_softmax_nll_fwd_bwd(const float *z, const int32_t *y, const int32_t *idx,
                           const int32_t *ctl, int B, int n, float inv_global_batch,
                           float *logprob, float *g, float *rowloss, void *stream);
/* test twin: logprob, preds = argmax (first maximum, int64), stats[0] = mean(pred != y),
 * stats[1] = mean(p[y])  (outlayers.py:69-80).  stats is float[2 + 2*B]; the tail is scratch. */
int tn_softmax_test_stats(const float *z, const int32_t *y, const int32_t *idx,
                          const int32_t *ctl, int B, int n, float *logprob, int64_t *preds,
                          float *stats, void *stream);

/* ---- the other output layers and losses (theanet/layer/outlayers.py:38-64,105-147) -------------
 * kind says how the scores z = x.w + b become (features, logprob, probs):
 *   TN_OUT_SOFTMAX  SoftmaxLayer   features = logprob = log softmax(z)                 (:83-102)
 *   TN_OUT_EXPLOSS  ExpLossLayer   o = z - mean_j z; features = o; logprob = log softmax(o) (:105-126)
 *   TN_OUT_HINGE    HingeLayer     features = logprob = probs = z                       (:129-147)
 * loss is the per-row term of the cost (cost = sum_b rowloss[b] / global batch):
 *   TN_LOSS_NLL -lp[y] (:50-51); TN_LOSS_NLLSQ lp[y]^2 (:41-42); TN_LOSS_NLLTRUNC
 *   max(0, log_threshold - lp[y]) (:44-48, 'nllNN' -> threshold NN/100); TN_LOSS_EXP exp(-o[y])
 *   (:38-39); TN_LOSS_HINGE mean_j max(0, z_j + 1 - z_y) over ALL j (:62-64).
 * SOFTMAX takes NLL / NLLSQ / NLLTRUNC, EXPLOSS takes EXP, HINGE takes HINGE (as in the reference's
 * constructors); anything else is TN_ERR_UNSUPPORTED.  g = dL/dz * inv_global_batch; logprob may
 * be NULL. */
enum { TN_OUT_SOFTMAX = 0, TN_OUT_EXPLOSS = 1, TN_OUT_HINGE = 2 };
enum { TN_LOSS_NLL = 0, TN_LOSS_NLLSQ = 1, TN_LOSS_NLLTRUNC = 2, TN_LOSS_EXP = 3, TN_LOSS_HINGE = 4 };
int tn_output_loss_fwd_bwd(const float *z, const int32_t *y, const int32_t *idx,
                           const int32_t *ctl, int B, int n, int kind, int loss,
                           float log_threshold, float inv_global_batch, float *features,
                           float *logprob, float *g, float *rowloss, void *stream);
/* test twin for any kind (outlayers.py:66-80): preds = argmax of the scores (first maximum),
 * stats[0] = mean(pred != y), stats[1] = mean(probs[y]) -- probs = the raw scores for HINGE
 * (:138).  features / logprob / preds may be NULL; stats is float[2 + 2*B]. */
int tn_output_test_stats(const float *z, const int32_t *y, const int32_t *idx, const int32_t *ctl,
                         int B, int n, int kind, float *features, float *logprob, int64_t *preds,
                         float *stats, void *stream);

/* ---- classifier head: HiddenLayer -> SoftmaxLayer with n_in <= 1024, n_out <= 32, fused ---------
 * (hidden.py:30-32 + outlayers.py:50-51,83-102 and their gradients in two launches) */
int tn_softmax_head_supported(int n_in, int n_out);
/* scores = h.W + b; logprob, g, rowloss as tn_softmax_nll_fwd_bwd; dh = g.W^T (may be NULL) and,
 * if below != 0, dh *= mask_below * act_below'(h): dL/dz of the layer that produced h */
int tn_softmax_head_fwd_bwd(const float *h, const float *W, const float *bias, const int32_t *y,
                            const int32_t *idx, const int32_t *ctl, int B, int n_in, int n_out,
                            float inv_global_batch, float *logprob, float *g, float *rowloss,
                            float *dh, int below, int act_below, int nn_below,
                            double pkeep_below, uint64_t seed_below, const float *mask_inj_below,
                            void *stream);
size_t tn_softmax_head_workspace_bytes(int B, int n_in, int n_out);
/* dW = h^T.g, db = column sums of g.  `workspace` must be zero-filled ONCE by the caller before
 * the first call (it holds the completion tickets, which every launch leaves at zero).  If rowloss
 * != NULL the launch also writes nll_sum[0] = sum_b rowloss[b] (what tn_reduce_rowloss computes). */
int tn_softmax_head_bwd_weights(const float *h, const float *g, float *dW, float *db,
                                void *workspace, int B, int n_in, int n_out, const float *rowloss,
                                float *nll_sum, void *stream);

/* ---- Layer.get_updates / get_wtcost (theanet/layer/layer.py:70-117) ------------------------- */
typedef struct tn_param_seg {
  int64_t offset;   /* element offset into the flat theta / velocity / gradient buffers */
  int64_t size;
  int32_t ndim;     /* 1, 2 (rows x cols, norm over axis 0) or 4 (rows = out kernels) */
  int32_t rows;
  int32_t cols;
  float momentum;
  float rate;       /* 0 freezes the tensor (layer.py:74-75) */
  float maxnorm;    /* 0 = off */
  float l1;
  float l2;
} tn_param_seg;

size_t tn_update_workspace_bytes(int nseg, int64_t total);
/* One step for all tensors: g' = g*grad_scale + L1*sgn(theta) + 2*L2*theta;
 * v' = m*v + (1-m)*g'; theta' = theta - rate*lr*v (OLD v); maxnorm on theta'.
 * cost_out[0] = (sum(rowloss[0..n_rowloss)) or nll_sum[0]) * nll_scale + L1/L2 weight cost of the
 * PRE-update theta.  segs_host is copied into the launch (<= 64 segments). lr comes from ctl.
 * `workspace` (tn_update_workspace_bytes) must be zero-filled ONCE by the caller before the first
 * call: its tail is the completion ticket of the cost reduction, left at zero by every launch. */
int tn_sgd_momentum_maxnorm_update(float *theta, float *vel, const float *grad,
                                   const tn_param_seg *segs_host, int nseg, int64_t total,
                                   const int32_t *ctl, float grad_scale, const float *nll_sum,
                                   float nll_scale, float *cost_out, void *workspace,
                                   void *stream);
/* Data parallel, fused: the same update with the gradient all-reduce folded in.  peer_grads[r] /
 * peer_flags[r] (host arrays of `world` device pointers) are rank r's flat gradient buffer
 * (total + 4 floats; [total] is its NLL partial sum) and flag array (int[32], zero-filled once:
 * words 0..7 receive the ranks' tokens, word 8 counts this rank's executions of the kernel),
 * mapped into this process with tn_ipc_open_handle (entry [rank] is the local buffer).  The kernel
 * signals and waits for all ranks, sums the buffers in rank order while it updates, and needs the
 * caller to alternate between two gradient buffers from step to step (see update.cu).
 * peer_end (the offset of a segment, or 0 / total = everything): only flat indices below it are
 * summed over the ranks here; the rest of the buffer and its NLL slot must already hold the sum
 * over all ranks (NeuralNet all-reduces the big dense-layer gradients with NCCL while the conv
 * layers are still back-propagating and leaves the few late conv gradients to this kernel). */
int tn_allreduce_sgd_update(float *theta, float *vel, const float *const *peer_grads,
                            int *const *peer_flags, int world, int rank,
                            const tn_param_seg *segs_host, int nseg, int64_t total,
                            int64_t peer_end, const int32_t *ctl, float grad_scale, float nll_scale,
                            float *cost_out, void *workspace, void *stream);
/* Two-shot all-reduce (sum) of peer_bufs[*][offset, offset + count) over peer memory, in place: every
 * rank reduces 1/world of the range from all ranks (rank order: bit-identical totals everywhere),
 * then gathers the other ranks' totals.  peer_bufs / peer_flags as above, except that the flag arrays
 * must hold int[32] (words 16..25 belong to this kernel).  Like tn_allreduce_sgd_update it relies on
 * the caller alternating between two buffers from step to step.  offset, count: multiples of 4. */
int tn_peer_allreduce(float *const *peer_bufs, int *const *peer_flags, int world, int rank,
                      int64_t offset, int64_t count, void *stream);
/* peer-mappable device memory (cudaMalloc, zero-filled) and CUDA-IPC handles (64 bytes) */
int tn_peer_alloc(size_t bytes, void **ptr);
int tn_peer_free(void *ptr);
int tn_ipc_get_handle(void *ptr, void *handle_out);
int tn_ipc_open_handle(const void *handle, void **ptr);
int tn_ipc_close_handle(void *ptr);
/* nll_sum[0] = sum_b rowloss[b] (fixed order, deterministic) */
int tn_reduce_rowloss(const float *rowloss, int B, float *nll_sum, void *stream);

#ifdef __cplusplus
}
#endif
#endif
