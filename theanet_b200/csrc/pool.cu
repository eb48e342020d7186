// PoolLayer kernels (theanet/layer/convpool.py:97-127): max pooling with window = stride = p, no
// padding, out = ceil(S/p) (partial edge windows kept) or floor(S/p) (ignore_border).  Pure HBM
// kernels.  The backward pass reproduces Theano's MaxPoolGrad -- every element equal to its
// window's maximum receives the gradient -- and fuses the activation derivative of the layer that
// produced the pooled tensor, so it emits dL/dz of that layer directly.
#include "common.cuh"

namespace tn {

struct PoolGeom {
  int S, p, O;
  FastDiv32 divO, divS, divp;
};

// 32-bit indexing (tensors below 2^32 elements) with multiply-high divisions
template <int P>
__global__ void maxpool_fwd_kernel(const float *__restrict__ x, float *__restrict__ out,
                                   uint32_t total, PoolGeom k) {
  const int p = P > 0 ? P : k.p;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t r = k.divO.div(t);
    const int oj = (int)(t - r * k.O);
    const uint32_t plane = k.divO.div(r);
    const int oi = (int)(r - plane * k.O);
    const float *src = x + (size_t)plane * k.S * k.S;
    const int y0 = oi * p, x0 = oj * p;
    const int y1 = min(y0 + p, k.S), x1 = min(x0 + p, k.S);
    float m = -INFINITY;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) m = fmaxf(m, src[yy * k.S + xx]);
    out[t] = m;
  }
}

__global__ void maxpool_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ x,
                                   const float *__restrict__ out, float *__restrict__ dx,
                                   uint32_t total, PoolGeom k, ActK ak) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t r = k.divS.div(t);
    const int xx = (int)(t - r * k.S);
    const uint32_t plane = k.divS.div(r);
    const int yy = (int)(r - plane * k.S);
    const int oi = (int)k.divp.div((uint32_t)yy), oj = (int)k.divp.div((uint32_t)xx);
    float g = 0.f;
    if (oi < k.O && oj < k.O) {  // rows/cols beyond O*p exist only with ignore_border
      const size_t o = ((size_t)plane * k.O + oi) * k.O + oj;
      const float a = x[t];
      if (a == out[o]) g = dout[o] * act_bwd_k(ak, a);
    }
    dx[t] = g;
  }
}

// ---- MeanLayer (theanet/layer/convpool.py:129-144): tt.mean(x, axis=(2,3)) ----------------------
// one warp per (sample, map) plane; fixed-order lane-strided sums + shuffle tree (deterministic)
__global__ void meanpool_fwd_kernel(const float *__restrict__ x, float *__restrict__ out,
                                    int planes, int hw) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < planes; p += gridDim.x * wpb) {
    const float *src = x + (size_t)p * hw;
    float s = 0.f;
    for (int i = lane; i < hw; i += 32) s += src[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[p] = __fdiv_rn(s, (float)hw);       // sum / count, as Theano's mean
  }
}

// dx[plane, i] = dout[plane] / (S*S) * act'(x[plane, i]); x = the output of the layer below
__global__ void meanpool_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ x,
                                    float *__restrict__ dx, uint32_t total, FastDiv32 divhw,
                                    float hw, ActK ak) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t plane = divhw.div(t);
    const float g = __fdiv_rn(dout[plane], hw);
    dx[t] = x ? g * act_bwd_k(ak, x[t]) : g;
  }
}

// ---- strided convolution = stride-1 convolution + subsampling (convpool.py:54-56,69-70) ----------
// out[p,i,j] = x[p, i*s, j*s]
__global__ void subsample2d_kernel(const float *__restrict__ x, float *__restrict__ out,
                                   uint32_t total, int S, int s, int O, FastDiv32 divO) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t r = divO.div(t);
    const int j = (int)(t - r * O);
    const uint32_t plane = divO.div(r);
    const int i = (int)(r - plane * O);
    out[t] = x[((size_t)plane * S + (size_t)i * s) * S + (size_t)j * s];
  }
}

// its gradient: out[p,y,x] = g[p, y/s, x/s] on the sampled lattice, 0 elsewhere
__global__ void upsample2d_zero_kernel(const float *__restrict__ g, float *__restrict__ out,
                                       uint32_t total, int S, int s, int O, FastDiv32 divS,
                                       FastDiv32 divs) {
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const uint32_t r = divS.div(t);
    const uint32_t xx = t - r * S;
    const uint32_t plane = divS.div(r);
    const uint32_t yy = r - plane * S;
    const uint32_t i = divs.div(yy), j = divs.div(xx);
    const bool on = i * s == yy && j * s == xx && (int)i < O && (int)j < O;
    out[t] = on ? g[((size_t)plane * O + i) * O + j] : 0.f;
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_subsample2d(const float *x, float *out, int planes, int S, int stride,
                              int out_sz, void *stream) {
  TN_REQUIRE(x && out, TN_ERR_ARG, "tn_subsample2d: null argument");
  TN_REQUIRE(planes > 0 && S > 0 && stride > 0 && out_sz > 0 && (out_sz - 1) * stride < S,
             TN_ERR_SHAPE, "tn_subsample2d: bad shape S=%d stride=%d out=%d", S, stride, out_sz);
  const int64_t total = (int64_t)planes * out_sz * out_sz;
  TN_REQUIRE((int64_t)planes * S * S < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_subsample2d: tensor too large");
  const int blocks = (int)min64(ceil_div64(total, 256), (int64_t)kNumSM * 32);
  subsample2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, (uint32_t)total, S, stride,
                                                              out_sz, FastDiv32(out_sz));
  TN_LAUNCH_CHECK("tn_subsample2d");
  return TN_OK;
}

extern "C" int tn_upsample2d_zero(const float *g, float *out, int planes, int S, int stride,
                                  int out_sz, void *stream) {
  TN_REQUIRE(g && out, TN_ERR_ARG, "tn_upsample2d_zero: null argument");
  TN_REQUIRE(planes > 0 && S > 0 && stride > 0 && out_sz > 0 && (out_sz - 1) * stride < S,
             TN_ERR_SHAPE, "tn_upsample2d_zero: bad shape S=%d stride=%d out=%d", S, stride, out_sz);
  const int64_t total = (int64_t)planes * S * S;
  TN_REQUIRE(total < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_upsample2d_zero: tensor too large");
  const int blocks = (int)min64(ceil_div64(total, 256), (int64_t)kNumSM * 32);
  upsample2d_zero_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      g, out, (uint32_t)total, S, stride, out_sz, FastDiv32(S), FastDiv32(stride));
  TN_LAUNCH_CHECK("tn_upsample2d_zero");
  return TN_OK;
}

extern "C" int tn_meanpool_fwd(const float *x, float *out, int planes, int S, void *stream) {
  TN_REQUIRE(x && out, TN_ERR_ARG, "tn_meanpool_fwd: null argument");
  TN_REQUIRE(planes > 0 && S > 0, TN_ERR_SHAPE, "tn_meanpool_fwd: bad shape planes=%d S=%d", planes, S);
  const int wpb = 8;
  const int blocks = (int)min64(ceil_div64(planes, wpb), (int64_t)kNumSM * 8);
  meanpool_fwd_kernel<<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x, out, planes, S * S);
  TN_LAUNCH_CHECK("tn_meanpool_fwd");
  return TN_OK;
}

extern "C" int tn_meanpool_bwd(const float *dout, const float *x, float *dx, int planes, int S,
                               int act, int act_nn, void *stream) {
  TN_REQUIRE(dout && dx, TN_ERR_ARG, "tn_meanpool_bwd: null argument");
  TN_REQUIRE(planes > 0 && S > 0, TN_ERR_SHAPE, "tn_meanpool_bwd: bad shape");
  const int64_t total = (int64_t)planes * S * S;
  TN_REQUIRE(total < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_meanpool_bwd: tensor too large");
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 32);
  meanpool_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      dout, x, dx, (uint32_t)total, FastDiv32(S * S), (float)(S * S), make_actk(act, act_nn));
  TN_LAUNCH_CHECK("tn_meanpool_bwd");
  return TN_OK;
}

extern "C" int tn_maxpool_fwd(const float *x, float *out, int planes, int S, int p, int out_sz,
                              void *stream) {
  TN_REQUIRE(x && out, TN_ERR_ARG, "tn_maxpool_fwd: null argument");
  TN_REQUIRE(planes > 0 && S > 0 && p > 0 && out_sz > 0 && out_sz * p < S + p, TN_ERR_SHAPE,
             "tn_maxpool_fwd: bad shape planes=%d S=%d p=%d out=%d", planes, S, p, out_sz);
  const int64_t total = (int64_t)planes * out_sz * out_sz;
  TN_REQUIRE((int64_t)planes * S * S < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_maxpool_fwd: tensor too large");
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 32);
  cudaStream_t st = (cudaStream_t)stream;
  PoolGeom k{S, p, out_sz, FastDiv32(out_sz), FastDiv32(S), FastDiv32(p)};
  if (p == 2) maxpool_fwd_kernel<2><<<blocks, threads, 0, st>>>(x, out, (uint32_t)total, k);
  else if (p == 3) maxpool_fwd_kernel<3><<<blocks, threads, 0, st>>>(x, out, (uint32_t)total, k);
  else maxpool_fwd_kernel<0><<<blocks, threads, 0, st>>>(x, out, (uint32_t)total, k);
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
  TN_REQUIRE(total < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_maxpool_bwd: tensor too large");
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 32);
  PoolGeom k{S, p, out_sz, FastDiv32(out_sz), FastDiv32(S), FastDiv32(p)};
  maxpool_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      dout, x, out, dx, (uint32_t)total, k, make_actk(act, act_nn));
  TN_LAUNCH_CHECK("tn_maxpool_bwd");
  return TN_OK;
}
