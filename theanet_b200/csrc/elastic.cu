// ElasticLayer kernels (theanet/layer/inlayers.py:29-163).
//
//   tn_elastic_noise : N(0,1) field on the (seed, step) Philox stream            (:94)
//   tn_elastic_field : the per-minibatch sampling grid, float64 coordinates       (:77-122)
//   tn_elastic_warp  : batch loader: invert, nearest / bilinear gather, flip noise (:63-64,124-142)
//
// The grid does not depend on the batch (one transform per minibatch), so the two
// field kernels are latency-bound; tn_elastic_warp is the HBM kernel (reads and writes each pixel
// once; the gather stays inside one image, i.e. inside L1/L2).
#include "common.cuh"

namespace tn {

// ---------------------------------------------------------------------------------------------
// the 4 standard normals of Philox block t of the (seed, step) noise stream (Box-Muller in fp64)
__device__ __forceinline__ void noise_block(uint64_t seed, uint32_t step, int t, float (&z)[4]) {
  const Philox4 r = philox_block(seed, TN_RNG_NOISE, step, 0u, (uint32_t)t);
  const double k = 1.0 / 4294967296.0;
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const double u1 = ((double)w[2 * p] + 0.5) * k;
    const double u2 = ((double)w[2 * p + 1] + 0.5) * k;
    const double rad = sqrt(-2.0 * log(u1));
    const double ang = 6.283185307179586 * u2;
    double s, c;
    sincos(ang, &s, &c);
    z[2 * p] = (float)(rad * c);
    z[2 * p + 1] = (float)(rad * s);
  }
}

__global__ void elastic_noise_kernel(float *__restrict__ noise, int n, uint64_t seed,
                                     const int32_t *__restrict__ ctl) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;  // Philox block index: 4 words -> 4 normals
  if (4 * t >= n) return;
  float z[4];
  noise_block(seed, (uint32_t)ctl[TN_CTL_STEP], t, z);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (4 * t + q < n) noise[4 * t + q] = z[q];
}

// ---------------------------------------------------------------------------------------------
struct FieldScalars {
  double t[2];       // translation
  double origin[2];
  double zoomer[2];
  double c, s;
};

__device__ __forceinline__ float two_u_minus_one(float u) {
  return __fsub_rn(__fmul_rn(2.f, u), 1.f);
}

__global__ void elastic_field_kernel(tn_elastic_prm prm, const float *__restrict__ noise,
                                     const float *__restrict__ u_inj,
                                     const float *__restrict__ filt, uint64_t seed,
                                     const int32_t *__restrict__ ctl, double *__restrict__ target,
                                     double *__restrict__ tyx, int32_t *__restrict__ gidx,
                                     float *__restrict__ gfrac) {
  extern __shared__ float sm[];
  const int h = prm.h, hw = h * h;
  const int k = 2 * prm.sigma + 1;
  float *s_filt = sm;                 // k*k
  float *s_el = sm + k * k;           // 2*h*h, magnitude * noise in float32 (inlayers.py:94)
  __shared__ FieldScalars sc;

  if (prm.magnitude != 0.f) {
    for (int i = threadIdx.x; i < k * k; i += blockDim.x) s_filt[i] = filt[i];
    if (noise) {
      for (int i = threadIdx.x; i < 2 * hw; i += blockDim.x)
        s_el[i] = __fmul_rn(prm.magnitude, noise[i]);
    } else {   // draw the field here (every CTA needs all of it): one launch less per step
      const uint32_t step = (uint32_t)(ctl[TN_CTL_STEP] + prm.step_offset);
      for (int t = threadIdx.x; 4 * t < 2 * hw; t += blockDim.x) {
        float z[4];
        noise_block(seed, step, t, z);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * t + q < 2 * hw) s_el[4 * t + q] = __fmul_rn(prm.magnitude, z[q]);
      }
    }
  }
  if (threadIdx.x == 0) {
    float u[8];
    if (u_inj) {
      for (int i = 0; i < 8; ++i) u[i] = u_inj[i];
    } else {
      const uint32_t step = (uint32_t)(ctl[TN_CTL_STEP] + prm.step_offset);
      const Philox4 a = philox_block(seed, TN_RNG_SCALARS, step, 0u, 0u);
      const Philox4 b = philox_block(seed, TN_RNG_SCALARS, step, 0u, 1u);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      for (int i = 0; i < 8; ++i) u[i] = (float)(((double)w[i] + 0.5) * (1.0 / 4294967296.0));
    }
    for (int c = 0; c < 2; ++c) {
      sc.t[c] = (double)__fmul_rn(prm.translation, two_u_minus_one(u[c]));               // :80-82
      sc.origin[c] = __dmul_rn((double)__fadd_rn(0.25f, __fmul_rn(0.5f, u[2 + c])), (double)h);
      const double e = __dmul_rn((double)prm.log_zoom, (double)two_u_minus_one(u[4 + c]));
      sc.zoomer[c] = (double)(float)exp(e);                                               // :106-108
    }
    const float theta = __fmul_rn(prm.angle_rad, two_u_minus_one(u[6]));                  // :112
    sc.c = (double)(float)cos((double)theta);
    sc.s = (double)(float)sin((double)theta);
  }
  __syncthreads();

  // one warp per pixel: lanes split the filter rows, fixed-order shuffle tree combines them
  const int lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= hw) return;
  const int y = pix / h, x = pix % h;
  double T[2] = {(double)y, (double)x};                                                   // :77
  if (prm.translation != 0.f) {
    T[0] = __dadd_rn(T[0], sc.t[0]);
    T[1] = __dadd_rn(T[1], sc.t[1]);
  }
  if (prm.magnitude != 0.f) {                                                             // :85-97
    const int sg = prm.sigma;
    const int j0 = max(0, sg - x), j1 = min(k, h + sg - x);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const float *el = s_el + c * hw;
      double acc = 0.0;
      for (int i = lane; i < k; i += 32) {
        const int yy = y + i - sg;
        if (yy < 0 || yy >= h) continue;
        const float *frow = s_filt + (k - 1 - i) * k;
        const float *erow = el + yy * h + x - sg;
        for (int j = j0; j < j1; ++j)
          acc = __dadd_rn(acc, __dmul_rn((double)erow[j], (double)frow[k - 1 - j]));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
      T[c] = __dadd_rn(T[c], (double)(float)acc);
    }
  }
  if (lane != 0) return;
  if (prm.zoom_on || prm.angle_rad != 0.f) {                                              // :100-118
    T[0] = __dsub_rn(T[0], sc.origin[0]);
    T[1] = __dsub_rn(T[1], sc.origin[1]);
    if (prm.zoom_on) {
      T[0] = __dmul_rn(T[0], sc.zoomer[0]);
      T[1] = __dmul_rn(T[1], sc.zoomer[1]);
    }
    if (prm.angle_rad != 0.f) {  // R^T . target, R = [[c,-s],[s,c]]
      const double t0 = __dadd_rn(__dmul_rn(sc.c, T[0]), __dmul_rn(sc.s, T[1]));
      const double t1 = __dadd_rn(__dmul_rn(-sc.s, T[0]), __dmul_rn(sc.c, T[1]));
      T[0] = t0;
      T[1] = t1;
    }
    T[0] = __dadd_rn(T[0], sc.origin[0]);
    T[1] = __dadd_rn(T[1], sc.origin[1]);
  }
  if (target) {
    target[pix] = T[0];
    target[hw + pix] = T[1];
  }
  const double ty = fmin(fmax(T[0], 0.0), prm.clip_hi);                                   // :121-122
  const double tx = fmin(fmax(T[1], 0.0), prm.clip_hi);
  if (tyx) {
    tyx[pix] = ty;
    tyx[hw + pix] = tx;
  }
  if (prm.nearest) {                       // iround: half away from zero; coordinates are >= 0
    const int vert = (int)floor(ty + 0.5), horz = (int)floor(tx + 0.5);
    gidx[pix] = vert * h + horz;
  } else {
    const int topp = (int)ty, left = (int)tx;                                             // :129-132
    gidx[pix] = topp * h + left;
    gfrac[pix] = (float)(ty - (double)topp);
    gfrac[hw + pix] = (float)(tx - (double)left);
  }
}

// ---------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float warp_sample(const float *__restrict__ src, int pix, int h, int hw,
                                             const int32_t *__restrict__ gidx,
                                             const float *__restrict__ gfrac, bool invert) {
  if (MODE == 0) {
    const float v = src[pix];
    return invert ? __fsub_rn(1.f, v) : v;
  } else if (MODE == 1) {
    const float v = src[gidx[pix]];
    return invert ? __fsub_rn(1.f, v) : v;
  } else {
    const int b = gidx[pix];
    const float fy = gfrac[pix], fx = gfrac[hw + pix];
    float v00 = src[b], v01 = src[b + 1], v10 = src[b + h], v11 = src[b + h + 1];
    if (invert) {
      v00 = __fsub_rn(1.f, v00); v01 = __fsub_rn(1.f, v01);
      v10 = __fsub_rn(1.f, v10); v11 = __fsub_rn(1.f, v11);
    }
    const float gy = __fsub_rn(1.f, fy), gx = __fsub_rn(1.f, fx);
    float r = __fmul_rn(__fmul_rn(v00, gy), gx);                                          // :134-137
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v01, gy), fx));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v10, fy), gx));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v11, fy), fx));
    return r;
  }
}

template <int MODE>
__global__ void elastic_warp_kernel(const float *__restrict__ corpus,
                                    const int32_t *__restrict__ idx,
                                    const int32_t *__restrict__ ctl, int B, int n, int h,
                                    int invert, const int32_t *__restrict__ gidx,
                                    const float *__restrict__ gfrac, int flip_on,
                                    uint32_t flip_thr, const float *__restrict__ flip_inj,
                                    uint64_t seed, float *__restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int hw = h * h;
  const int ngroups = (n + 3) >> 2;
  const int64_t total = (int64_t)B * ngroups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / ngroups), gq = (int)(t % ngroups);
    const int64_t row = idx ? (int64_t)idx[b] : (int64_t)ctl[TN_CTL_ROW0] + b;
    const float *src_img = corpus + row * n;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = 4 * gq + q;
      if (j < n) {
        const int c = j / hw, pix = j - c * hw;
        v[q] = warp_sample<MODE>(src_img + c * hw, pix, h, hw, gidx, gfrac, invert != 0);
      } else {
        v[q] = 0.f;
      }
    }
    if (flip_inj) {                                                                       // :140-142
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * gq + q;
        if (j < n) {
          const float m = flip_inj[(int64_t)b * n + j];
          v[q] = __fadd_rn(__fmul_rn(__fsub_rn(1.f, v[q]), m), __fmul_rn(v[q], __fsub_rn(1.f, m)));
        }
      }
    } else if (flip_on) {
      const Philox4 r = philox_block(seed, TN_RNG_FLIP, (uint32_t)ctl[TN_CTL_STEP],
                                     (uint32_t)(ctl[TN_CTL_SAMPLE0] + b), (uint32_t)gq);
      if (r.x < flip_thr) v[0] = __fsub_rn(1.f, v[0]);
      if (r.y < flip_thr) v[1] = __fsub_rn(1.f, v[1]);
      if (r.z < flip_thr) v[2] = __fsub_rn(1.f, v[2]);
      if (r.w < flip_thr) v[3] = __fsub_rn(1.f, v[3]);
    }
    float *o = out + (int64_t)b * n + 4 * gq;
    if ((n & 3) == 0) {
      *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (4 * gq + q < n) o[q] = v[q];
    }
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_elastic_noise(float *noise, int h, uint64_t seed, const int32_t *ctl,
                                void *stream) {
  TN_REQUIRE(noise && ctl && h > 0, TN_ERR_ARG, "tn_elastic_noise: bad argument");
  const int n = 2 * h * h;
  const int nblk = ceil_div(n, 4);
  elastic_noise_kernel<<<ceil_div(nblk, 128), 128, 0, (cudaStream_t)stream>>>(noise, n, seed, ctl);
  TN_LAUNCH_CHECK("tn_elastic_noise");
  return TN_OK;
}

extern "C" int tn_elastic_field(const tn_elastic_prm *prm, const float *noise, const float *u_inj,
                                const float *filt, uint64_t seed, const int32_t *ctl,
                                double *target, double *tyx, int32_t *gidx, float *gfrac,
                                void *stream) {
  TN_REQUIRE(prm && gidx, TN_ERR_ARG, "tn_elastic_field: null argument");
  TN_REQUIRE(prm->h > 0 && prm->sigma >= 0, TN_ERR_SHAPE, "tn_elastic_field: bad h/sigma");
  TN_REQUIRE(u_inj || ctl, TN_ERR_ARG, "tn_elastic_field: need ctl or injected uniforms");
  TN_REQUIRE(prm->nearest || gfrac, TN_ERR_ARG, "tn_elastic_field: bilinear needs gfrac");
  TN_REQUIRE(prm->magnitude == 0.f || (filt && (noise || ctl)), TN_ERR_ARG,
             "tn_elastic_field: magnitude != 0 needs the filter table and noise (or ctl to draw it)");
  const int k = 2 * prm->sigma + 1;
  const int hw = prm->h * prm->h;
  const size_t smem = prm->magnitude != 0.f ? (size_t)(k * k + 2 * hw) * sizeof(float) : 0;
  TN_REQUIRE(smem <= 200 * 1024, TN_ERR_UNSUPPORTED,
             "tn_elastic_field: img_sz %d / sigma %d exceed the shared-memory field (%zu B)",
             prm->h, prm->sigma, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(elastic_field_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_elastic_field: %s", cudaGetErrorString(e));
  }
  const int threads = 256;  // 8 warps = 8 pixels per CTA
  elastic_field_kernel<<<ceil_div(hw, threads / 32), threads, smem, (cudaStream_t)stream>>>(
      *prm, noise, u_inj, filt, seed, ctl, target, tyx, gidx, gfrac);
  TN_LAUNCH_CHECK("tn_elastic_field");
  return TN_OK;
}

extern "C" int tn_elastic_warp(const float *corpus, const int32_t *idx, const int32_t *ctl, int B,
                               int C, int h, int invert, int mode, const int32_t *gidx,
                               const float *gfrac, double pflip, const float *flip_inj,
                               uint64_t seed, float *out, void *stream) {
  TN_REQUIRE(corpus && out && ctl, TN_ERR_ARG, "tn_elastic_warp: null argument");
  TN_REQUIRE(B > 0 && C > 0 && h > 0, TN_ERR_SHAPE, "tn_elastic_warp: bad shape");
  TN_REQUIRE(mode >= 0 && mode <= 2, TN_ERR_ARG, "tn_elastic_warp: mode %d", mode);
  TN_REQUIRE(mode == 0 || gidx, TN_ERR_ARG, "tn_elastic_warp: gather needs gidx");
  TN_REQUIRE(mode != 2 || gfrac, TN_ERR_ARG, "tn_elastic_warp: bilinear needs gfrac");
  TN_REQUIRE(((uintptr_t)out & 15) == 0, TN_ERR_ALIGN, "tn_elastic_warp: out not 16B aligned");
  const int n = C * h * h;
  const int flip_on = pflip > 0.0;
  const uint32_t thr = pflip >= 1.0 ? 0xffffffffu : bernoulli_threshold(pflip);
  const int64_t total = (int64_t)B * ((n + 3) / 4);
  const int threads = 256;
  const int blocks = (int)min64(ceil_div64(total, threads), (int64_t)kNumSM * 16);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_WARP_LAUNCH(MODE)                                                                  \
  launch_pdl(elastic_warp_kernel<MODE>, dim3(blocks), dim3(threads), 0, st, corpus, idx, ctl, B, n, \
             h, invert, gidx, gfrac, flip_on, thr, flip_inj, seed, out)
  if (mode == 0) TN_WARP_LAUNCH(0);
  else if (mode == 1) TN_WARP_LAUNCH(1);
  else TN_WARP_LAUNCH(2);
#undef TN_WARP_LAUNCH
  TN_LAUNCH_CHECK("tn_elastic_warp");
  return TN_OK;
}
