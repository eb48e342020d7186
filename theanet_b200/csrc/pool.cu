// PoolLayer kernels (theanet/layer/convpool.py:97-127): max pooling with window = stride = p, no
// padding, out = ceil(S/p) (partial edge windows kept) or floor(S/p) (ignore_border).  Pure HBM
// kernels.  The backward pass reproduces Theano's MaxPoolGrad -- every element equal to its
// window's maximum receives the gradient -- and fuses the activation derivative of the layer that
// produced the pooled tensor, so it emits dL/dz of that layer directly.
#include "common.cuh"

namespace tn {

template <int P>
__global__ void maxpool_fwd_kernel(const float *__restrict__ x, float *__restrict__ out,
                                   int64_t total, int S, int p_rt, int O) {
  const int p = P > 0 ? P : p_rt;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int oj = (int)(t % O);
    const int64_t r = t / O;
    const int oi = (int)(r % O);
    const int64_t plane = r / O;
    const float *src = x + plane * S * S;
    const int y0 = oi * p, x0 = oj * p;
    const int y1 = min(y0 + p, S), x1 = min(x0 + p, S);
    float m = -INFINITY;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) m = fmaxf(m, src[yy * S + xx]);
    out[t] = m;
  }
}

__global__ void maxpool_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ x,
                                   const float *__restrict__ out, float *__restrict__ dx,
                                   int64_t total, int S, int p, int O, int act, float act_nn) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(t % S);
    const int64_t r = t / S;
    const int yy = (int)(r % S);
    const int64_t plane = r / S;
    const int oi = yy / p, oj = xx / p;
    float g = 0.f;
    if (oi < O && oj < O) {  // rows/cols beyond O*p exist only with ignore_border
      const int64_t o = (plane * O + oi) * O + oj;
      const float a = x[t];
      if (a == out[o]) g = dout[o] * act_bwd_from_out(a, act, act_nn);
    }
    dx[t] = g;
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_maxpool_fwd(const float *x, float *out, int planes, int S, int p, int out_sz,
                              void *stream) {
  TN_REQUIRE(x && out, TN_ERR_ARG, "tn_maxpool_fwd: null argument");
  TN_REQUIRE(planes > 0 && S > 0 && p > 0 && out_sz > 0 && out_sz * p < S + p, TN_ERR_SHAPE,
             "tn_maxpool_fwd: bad shape planes=%d S=%d p=%d out=%d", planes, S, p, out_sz);
  const int64_t total = (int64_t)planes * out_sz * out_sz;
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (p == 2) maxpool_fwd_kernel<2><<<blocks, threads, 0, st>>>(x, out, total, S, p, out_sz);
  else if (p == 3) maxpool_fwd_kernel<3><<<blocks, threads, 0, st>>>(x, out, total, S, p, out_sz);
  else maxpool_fwd_kernel<0><<<blocks, threads, 0, st>>>(x, out, total, S, p, out_sz);
  TN_LAUNCH_CHECK("tn_maxpool_fwd");
  return TN_OK;
}

extern "C" int tn_maxpool_bwd(const float *dout, const float *x, const float *out, float *dx,
                              int planes, int S, int p, int out_sz, int act, int act_nn,
                              void *stream) {
  TN_REQUIRE(dout && x && out && dx, TN_ERR_ARG, "tn_maxpool_bwd: null argument");
  TN_REQUIRE(planes > 0 && S > 0 && p > 0 && out_sz > 0 && out_sz * p < S + p, TN_ERR_SHAPE,
             "tn_maxpool_bwd: bad shape");
  const int64_t total = (int64_t)planes * S * S;
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 32);
  maxpool_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(dout, x, out, dx, total, S, p,
                                                                   out_sz, act, (float)act_nn);
  TN_LAUNCH_CHECK("tn_maxpool_bwd");
  return TN_OK;
}
