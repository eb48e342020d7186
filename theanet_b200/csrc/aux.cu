// Auxiliary-input layers (theanet/layer/auxiliary.py:14-160): the LocationInfo input mix and the
// column plumbing AuxConcatLayer / SoftAuxLayer need.  The two tiny dense layers of LocationInfo
// and all gradients run on the dense kernels (tn_dense_*); these are pure HBM element kernels.
#include "common.cuh"

namespace tn {

// aux (B,2,2) -> loc (B,2).  train: aux[b,0,:]*u + aux[b,1,:]*(1-u), u ~ U(0,1) per sample
// (auxiliary.py:25-28); test: mean over axis 1 (:31).  Then * boost (:33).
__global__ void aux_location_mix_kernel(const float *__restrict__ aux, float *__restrict__ loc,
                                        int B, float boost, int train, uint64_t seed,
                                        const int32_t *__restrict__ ctl,
                                        const float *__restrict__ u_inj) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float4 a = *reinterpret_cast<const float4 *>(aux + 4 * (size_t)b);   // [a00 a01 a10 a11]
  float l0, l1;
  if (train) {
    float u;
    if (u_inj) {
      u = u_inj[b];
    } else {
      const Philox4 r = philox_block(seed, TN_RNG_AUX, (uint32_t)ctl[TN_CTL_STEP],
                                     (uint32_t)(ctl[TN_CTL_SAMPLE0] + b), 0u);
      u = (float)(((double)r.x + 0.5) * (1.0 / 4294967296.0));
    }
    const float v = __fsub_rn(1.f, u);
    l0 = __fadd_rn(__fmul_rn(a.x, u), __fmul_rn(a.z, v));
    l1 = __fadd_rn(__fmul_rn(a.y, u), __fmul_rn(a.w, v));
  } else {
    l0 = __fmul_rn(__fadd_rn(a.x, a.z), 0.5f);
    l1 = __fmul_rn(__fadd_rn(a.y, a.w), 0.5f);
  }
  loc[2 * b] = __fmul_rn(l0, boost);
  loc[2 * b + 1] = __fmul_rn(l1, boost);
}

__global__ void concat_cols_kernel(const float *__restrict__ a, int na, const float *__restrict__ c,
                                   int nc, float *__restrict__ out, int64_t total) {
  const int n = na + nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / n;
    const int j = (int)(t - b * n);
    out[t] = j < na ? a[b * na + j] : c[b * nc + (j - na)];
  }
}

__global__ void slice_cols_kernel(const float *__restrict__ src, int stride, int off, int n,
                                  float *__restrict__ out, int64_t total) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / n;
    const int j = (int)(t - b * n);
    out[t] = src[b * stride + off + j];
  }
}

__global__ void add_inplace_kernel(float *__restrict__ dst, const float *__restrict__ src,
                                   int64_t n) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * blockDim.x)
    dst[t] = __fadd_rn(dst[t], src[t]);
}

static inline int grid_for(int64_t total, int threads) {
  return (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 16);
}

}  // namespace tn

using namespace tn;

extern "C" int tn_aux_location_mix(const float *aux, float *loc, int B, float boost, int train,
                                   uint64_t seed, const int32_t *ctl, const float *u_inj,
                                   void *stream) {
  TN_REQUIRE(aux && loc, TN_ERR_ARG, "tn_aux_location_mix: null argument");
  TN_REQUIRE(B > 0, TN_ERR_SHAPE, "tn_aux_location_mix: bad batch %d", B);
  TN_REQUIRE(!train || u_inj || ctl, TN_ERR_ARG, "tn_aux_location_mix: need ctl or injected draws");
  TN_REQUIRE(((uintptr_t)aux & 15) == 0, TN_ERR_ALIGN, "tn_aux_location_mix: aux must be 16-byte aligned");
  aux_location_mix_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(aux, loc, B, boost, train,
                                                                            seed, ctl, u_inj);
  TN_LAUNCH_CHECK("tn_aux_location_mix");
  return TN_OK;
}

extern "C" int tn_concat_cols(const float *a, int na, const float *c, int nc, float *out, int B,
                              void *stream) {
  TN_REQUIRE(a && c && out, TN_ERR_ARG, "tn_concat_cols: null argument");
  TN_REQUIRE(B > 0 && na > 0 && nc > 0, TN_ERR_SHAPE, "tn_concat_cols: bad shape");
  const int64_t total = (int64_t)B * (na + nc);
  concat_cols_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, na, c, nc, out, total);
  TN_LAUNCH_CHECK("tn_concat_cols");
  return TN_OK;
}

extern "C" int tn_slice_cols(const float *src, int stride, int off, int n, float *out, int B,
                             void *stream) {
  TN_REQUIRE(src && out, TN_ERR_ARG, "tn_slice_cols: null argument");
  TN_REQUIRE(B > 0 && n > 0 && off >= 0 && off + n <= stride, TN_ERR_SHAPE, "tn_slice_cols: bad shape");
  const int64_t total = (int64_t)B * n;
  slice_cols_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, stride, off, n, out, total);
  TN_LAUNCH_CHECK("tn_slice_cols");
  return TN_OK;
}

extern "C" int tn_add_inplace(float *dst, const float *src, int64_t n, void *stream) {
  TN_REQUIRE(dst && src && n > 0, TN_ERR_ARG, "tn_add_inplace: bad argument");
  add_inplace_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
  TN_LAUNCH_CHECK("tn_add_inplace");
  return TN_OK;
}
