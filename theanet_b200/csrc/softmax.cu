// SoftmaxLayer + negative log-likelihood (theanet/layer/outlayers.py:50-51,69-80,83-102).
// One warp per sample row: row max, log-sum-exp, log-probabilities, the NLL term and its gradient
// (softmax - onehot)/B in one pass with warp shuffles.  The test variant returns the first-maximum
// argmax (int64, as Theano's argmax) and the two error statistics.
#include "common.cuh"

namespace tn {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int label_of(const int32_t *y, const int32_t *idx, const int32_t *ctl,
                                        int b) {
  const int64_t row = idx ? (int64_t)idx[b] : (int64_t)ctl[TN_CTL_ROW0] + b;
  return y[row];
}

__global__ void softmax_nll_kernel(const float *__restrict__ z, const int32_t *__restrict__ y,
                                   const int32_t *__restrict__ idx,
                                   const int32_t *__restrict__ ctl, int B, int n, float inv_bg,
                                   float *__restrict__ logprob, float *__restrict__ g,
                                   float *__restrict__ rowloss) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float *zr = z + (size_t)b * n;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, zr[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(zr[j] - m);
  s = warp_sum(s);
  const float ls = logf(s);
  const int label = label_of(y, idx, ctl, b);
  for (int j = lane; j < n; j += 32) {
    const float lp = (zr[j] - m) - ls;
    logprob[(size_t)b * n + j] = lp;
    float p = expf(lp);
    if (j == label) {
      p -= 1.f;
      rowloss[b] = -lp;
    }
    g[(size_t)b * n + j] = p * inv_bg;
  }
}

__global__ void softmax_test_kernel(const float *__restrict__ z, const int32_t *__restrict__ y,
                                    const int32_t *__restrict__ idx,
                                    const int32_t *__restrict__ ctl, int B, int n,
                                    float *__restrict__ logprob, int64_t *__restrict__ preds,
                                    float *__restrict__ rowstat) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float *zr = z + (size_t)b * n;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    const float v = zr[j];
    if (v > m) { m = v; am = j; }  // first maximum within the lane's strided subsequence
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(zr[j] - m);
  s = warp_sum(s);
  const float ls = logf(s);
  const int label = label_of(y, idx, ctl, b);
  for (int j = lane; j < n; j += 32) {
    const float lp = (zr[j] - m) - ls;
    if (logprob) logprob[(size_t)b * n + j] = lp;
    if (j == label) rowstat[B + b] = expf(lp);
  }
  if (lane == 0) {
    if (preds) preds[b] = am;
    rowstat[b] = am != label ? 1.f : 0.f;
  }
}

// ---- the other output layers / losses (theanet/layer/outlayers.py:38-64,105-147) ---------------
// kind: how scores z become (features, logprob, probs); loss: the per-row term of the cost.
//   SOFTMAX  features = logprob = log softmax(z)                          (:83-102)
//   EXPLOSS  o = z - mean_j z; features = o; logprob = log softmax(o)      (:105-126)
//   HINGE    features = logprob = probs = z                                (:129-147)
//   nll  -lp[y] (:50-51)    nllsq  lp[y]^2 (:41-42)    nllNN  max(0, log(thr) - lp[y]) (:44-48)
//   exp  exp(-o[y]) (:38-39)    hinge  mean_j max(0, z_j + 1 - z_y), the j = y term included (:62-64)
// g = dL/dz * inv_bg.  maximum(0, a) passes the gradient where a >= 0 (Theano: eq(out, a)).
__global__ void out_loss_kernel(const float *__restrict__ z, const int32_t *__restrict__ y,
                                const int32_t *__restrict__ idx, const int32_t *__restrict__ ctl,
                                int B, int n, int kind, int loss, float log_thr, float inv_bg,
                                float *__restrict__ feat, float *__restrict__ logprob,
                                float *__restrict__ g, float *__restrict__ rowloss) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float *zr = z + (size_t)b * n;
  const size_t row = (size_t)b * n;
  const int label = label_of(y, idx, ctl, b);
  const float zy = zr[label];
  if (kind == TN_OUT_HINGE) {
    float cnt = 0.f, sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float a = zr[j] + 1.f - zy;
      if (a >= 0.f) { cnt += 1.f; sum += a; }
    }
    cnt = warp_sum(cnt);
    sum = warp_sum(sum);
    const float sc = inv_bg / (float)n;
    for (int j = lane; j < n; j += 32) {
      const float zj = zr[j];
      feat[row + j] = zj;
      if (logprob) logprob[row + j] = zj;
      float d = (zj + 1.f - zy >= 0.f) ? 1.f : 0.f;
      if (j == label) d -= cnt;
      g[row + j] = d * sc;
    }
    if (lane == 0) rowloss[b] = sum / (float)n;
    return;
  }
  float m = -INFINITY, tot = 0.f;
  for (int j = lane; j < n; j += 32) { m = fmaxf(m, zr[j]); tot += zr[j]; }
  m = warp_max(m);
  const float mean = kind == TN_OUT_EXPLOSS ? warp_sum(tot) / (float)n : 0.f;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(zr[j] - m);
  s = warp_sum(s);
  const float ls = logf(s);
  const float lpy = (zy - m) - ls;
  // d(loss)/d(lp[y]) for the softmax losses
  float dl = -1.f, lossv = -lpy;
  if (loss == TN_LOSS_NLLSQ) { dl = 2.f * lpy; lossv = lpy * lpy; }
  else if (loss == TN_LOSS_NLLTRUNC) {
    const float a = log_thr - lpy;
    dl = a >= 0.f ? -1.f : 0.f;
    lossv = fmaxf(0.f, a);
  }
  const float e = expf(-(zy - mean));                 // exp loss
  if (loss == TN_LOSS_EXP) lossv = e;
  for (int j = lane; j < n; j += 32) {
    const float lp = (zr[j] - m) - ls;                 // softmax is shift invariant: same for o
    feat[row + j] = kind == TN_OUT_EXPLOSS ? zr[j] - mean : lp;
    if (logprob) logprob[row + j] = lp;
    float d;
    if (loss == TN_LOSS_EXP) d = -e * ((j == label ? 1.f : 0.f) - 1.f / (float)n);
    else d = dl * ((j == label ? 1.f : 0.f) - expf(lp));
    g[row + j] = d * inv_bg;
  }
  if (lane == 0) rowloss[b] = lossv;
}

// test twin (outlayers.py:69-80): first-maximum argmax of the scores, mean(pred != y) and
// mean(probs[y]) with probs = softmax (SOFTMAX, EXPLOSS) or the raw scores (HINGE, :138)
__global__ void out_test_kernel(const float *__restrict__ z, const int32_t *__restrict__ y,
                                const int32_t *__restrict__ idx, const int32_t *__restrict__ ctl,
                                int B, int n, int kind, float *__restrict__ feat,
                                float *__restrict__ logprob, int64_t *__restrict__ preds,
                                float *__restrict__ rowstat) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float *zr = z + (size_t)b * n;
  const size_t row = (size_t)b * n;
  float m = -INFINITY, tot = 0.f;
  int am = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    const float v = zr[j];
    tot += v;
    if (v > m) { m = v; am = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  const float mean = kind == TN_OUT_EXPLOSS ? warp_sum(tot) / (float)n : 0.f;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(zr[j] - m);
  s = warp_sum(s);
  const float ls = logf(s);
  const int label = label_of(y, idx, ctl, b);
  for (int j = lane; j < n; j += 32) {
    const float lp = (zr[j] - m) - ls;
    const float f = kind == TN_OUT_HINGE ? zr[j] : kind == TN_OUT_EXPLOSS ? zr[j] - mean : lp;
    if (feat) feat[row + j] = f;
    if (logprob) logprob[row + j] = kind == TN_OUT_HINGE ? zr[j] : lp;
    if (j == label) rowstat[B + b] = kind == TN_OUT_HINGE ? zr[j] : expf(lp);
  }
  if (lane == 0) {
    if (preds) preds[b] = am;
    rowstat[b] = am != label ? 1.f : 0.f;
  }
}

// single CTA, fixed-order tree: out[k] = scale * sum_b in[k*B + b]
__global__ void reduce_rows_kernel(const float *__restrict__ in, int B, int nrows, float scale,
                                   float *__restrict__ out) {
  __shared__ float red[256];
  for (int k = 0; k < nrows; ++k) {
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += 256) s += in[(size_t)k * B + b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = red[0] * scale;
    __syncthreads();
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_softmax_nll_fwd_bwd(const float *z, const int32_t *y, const int32_t *idx,
                                      const int32_t *ctl, int B, int n, float inv_global_batch,
                                      float *logprob, float *g, float *rowloss, void *stream) {
  TN_REQUIRE(z && y && logprob && g && rowloss && (idx || ctl), TN_ERR_ARG,
             "tn_softmax_nll_fwd_bwd: null argument");
  TN_REQUIRE(B > 0 && n > 0, TN_ERR_SHAPE, "tn_softmax_nll_fwd_bwd: bad shape");
  const int wpb = 8;
  softmax_nll_kernel<<<ceil_div(B, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      z, y, idx, ctl, B, n, inv_global_batch, logprob, g, rowloss);
  TN_LAUNCH_CHECK("tn_softmax_nll_fwd_bwd");
  return TN_OK;
}

extern "C" int tn_softmax_test_stats(const float *z, const int32_t *y, const int32_t *idx,
                                     const int32_t *ctl, int B, int n, float *logprob,
                                     int64_t *preds, float *stats, void *stream) {
  // stats: float[2 + 2*B]; [0] = mean(pred != y), [1] = mean(p[y]); the tail is scratch
  TN_REQUIRE(z && y && stats && (idx || ctl), TN_ERR_ARG, "tn_softmax_test_stats: null argument");
  TN_REQUIRE(B > 0 && n > 0, TN_ERR_SHAPE, "tn_softmax_test_stats: bad shape");
  const int wpb = 8;
  cudaStream_t st = (cudaStream_t)stream;
  softmax_test_kernel<<<ceil_div(B, wpb), wpb * 32, 0, st>>>(z, y, idx, ctl, B, n, logprob, preds,
                                                            stats + 2);
  TN_LAUNCH_CHECK("tn_softmax_test_stats");
  reduce_rows_kernel<<<1, 256, 0, st>>>(stats + 2, B, 2, 1.f / (float)B, stats);
  TN_LAUNCH_CHECK("tn_softmax_test_stats(reduce)");
  return TN_OK;
}

extern "C" int tn_output_loss_fwd_bwd(const float *z, const int32_t *y, const int32_t *idx,
                                      const int32_t *ctl, int B, int n, int kind, int loss,
                                      float log_threshold, float inv_global_batch,
                                      float *features, float *logprob, float *g, float *rowloss,
                                      void *stream) {
  TN_REQUIRE(z && y && features && g && rowloss && (idx || ctl), TN_ERR_ARG,
             "tn_output_loss_fwd_bwd: null argument");
  TN_REQUIRE(B > 0 && n > 0, TN_ERR_SHAPE, "tn_output_loss_fwd_bwd: bad shape");
  const bool ok = (kind == TN_OUT_SOFTMAX && (loss == TN_LOSS_NLL || loss == TN_LOSS_NLLSQ ||
                                              loss == TN_LOSS_NLLTRUNC)) ||
                  (kind == TN_OUT_EXPLOSS && loss == TN_LOSS_EXP) ||
                  (kind == TN_OUT_HINGE && loss == TN_LOSS_HINGE);
  TN_REQUIRE(ok, TN_ERR_UNSUPPORTED, "tn_output_loss_fwd_bwd: kind %d does not take loss %d", kind,
             loss);
  const int wpb = 8;
  out_loss_kernel<<<ceil_div(B, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      z, y, idx, ctl, B, n, kind, loss, log_threshold, inv_global_batch, features, logprob, g,
      rowloss);
  TN_LAUNCH_CHECK("tn_output_loss_fwd_bwd");
  return TN_OK;
}

extern "C" int tn_output_test_stats(const float *z, const int32_t *y, const int32_t *idx,
                                    const int32_t *ctl, int B, int n, int kind, float *features,
                                    float *logprob, int64_t *preds, float *stats, void *stream) {
  TN_REQUIRE(z && y && stats && (idx || ctl), TN_ERR_ARG, "tn_output_test_stats: null argument");
  TN_REQUIRE(B > 0 && n > 0, TN_ERR_SHAPE, "tn_output_test_stats: bad shape");
  TN_REQUIRE(kind >= TN_OUT_SOFTMAX && kind <= TN_OUT_HINGE, TN_ERR_UNSUPPORTED,
             "tn_output_test_stats: unknown kind %d", kind);
  const int wpb = 8;
  cudaStream_t st = (cudaStream_t)stream;
  out_test_kernel<<<ceil_div(B, wpb), wpb * 32, 0, st>>>(z, y, idx, ctl, B, n, kind, features,
                                                         logprob, preds, stats + 2);
  TN_LAUNCH_CHECK("tn_output_test_stats");
  reduce_rows_kernel<<<1, 256, 0, st>>>(stats + 2, B, 2, 1.f / (float)B, stats);
  TN_LAUNCH_CHECK("tn_output_test_stats(reduce)");
  return TN_OK;
}

extern "C" int tn_reduce_rowloss(const float *rowloss, int B, float *nll_sum, void *stream) {
  TN_REQUIRE(rowloss && nll_sum && B > 0, TN_ERR_ARG, "tn_reduce_rowloss: bad argument");
  reduce_rows_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rowloss, B, 1, 1.f, nll_sum);
  TN_LAUNCH_CHECK("tn_reduce_rowloss");
  return TN_OK;
}
