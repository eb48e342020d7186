// HiddenLayer / SoftmaxLayer matrix products in float32 on the CUDA cores
// (theanet/layer/hidden.py:30-32 tt.dot; backward via tt.grad, layer.py:83).
//
// One register-tiled, double-buffered SGEMM serves the three products of a dense layer through
// its operand-major switches, with the layer's elementwise work fused into the epilogue:
//   fwd  : out = act(x.W + b) * dropout-mask * scale        (A k-contig,  B n-contig)
//   dX   : dx  = (g.W^T) [* mask_prev * act_prev'(prev_out)] (A k-contig,  B k-contig)
//   dW   : dW  = x^T.g                                        (A m-contig,  B n-contig)
// float32 FFMA keeps the fp32 configs (C1-C3, C5) inside the 1e-3 parity budget without any
// tensor-core input rounding; the bf16 config uses the tcgen05 path (gemm_tc.cu).
#include "common.cuh"
#include "dense_small.cuh"
#include "dense_tc.cuh"

namespace tn {

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;
constexpr int LDS_ = BM + 4;  // padded leading dimension of the shared tiles (floats)

struct GemmArgs {
  const float *A, *B;
  float *C;
  int M, N, K, lda, ldb, ldc;
  // epilogue
  const float *bias;      // EPI 0
  const float *aux;       // EPI 1: prev_out (may be null)
  const float *mask_inj;  // EPI 0/1: injected mask (may be null)
  const int32_t *ctl;
  uint64_t seed;
  uint32_t thr;           // Bernoulli threshold
  int mask_on;            // 0: none, 1: philox, 2: injected
  int act;
  float act_nn;
  float scale;
};

template <int AMODE, bool VEC>
__device__ __forceinline__ void load_a(const GemmArgs &g, int m0, int k0, int t, float (&r)[4]) {
  if (AMODE == 0) {  // A[m*lda + k], k contiguous
    const int row = t >> 2, kq = (t & 3) * 4;
    const int m = m0 + row, k = k0 + kq;
    if (VEC) {
      if (m < g.M && k < g.K) {
        const float4 v = *reinterpret_cast<const float4 *>(g.A + (size_t)m * g.lda + k);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else { r[0] = r[1] = r[2] = r[3] = 0.f; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] = (m < g.M && k + q < g.K) ? g.A[(size_t)m * g.lda + k + q] : 0.f;
    }
  } else {  // A[k*lda + m], m contiguous
    const int kk = t >> 4, mq = (t & 15) * 4;
    const int k = k0 + kk, m = m0 + mq;
    if (VEC) {
      if (k < g.K && m < g.M) {
        const float4 v = *reinterpret_cast<const float4 *>(g.A + (size_t)k * g.lda + m);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else { r[0] = r[1] = r[2] = r[3] = 0.f; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] = (k < g.K && m + q < g.M) ? g.A[(size_t)k * g.lda + m + q] : 0.f;
    }
  }
}

template <int AMODE>
__device__ __forceinline__ void store_a(float *As, int t, const float (&r)[4]) {
  if (AMODE == 0) {
    const int row = t >> 2, kq = (t & 3) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) As[(kq + q) * LDS_ + row] = r[q];
  } else {
    const int kk = t >> 4, mq = (t & 15) * 4;
    *reinterpret_cast<float4 *>(As + kk * LDS_ + mq) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

template <int BMODE, bool VEC>
__device__ __forceinline__ void load_b(const GemmArgs &g, int n0, int k0, int t, float (&r)[4]) {
  if (BMODE == 0) {  // B[k*ldb + n], n contiguous
    const int kk = t >> 4, nq = (t & 15) * 4;
    const int k = k0 + kk, n = n0 + nq;
    if (VEC) {
      if (k < g.K && n < g.N) {
        const float4 v = *reinterpret_cast<const float4 *>(g.B + (size_t)k * g.ldb + n);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else { r[0] = r[1] = r[2] = r[3] = 0.f; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] = (k < g.K && n + q < g.N) ? g.B[(size_t)k * g.ldb + n + q] : 0.f;
    }
  } else {  // B[n*ldb + k], k contiguous
    const int row = t >> 2, kq = (t & 3) * 4;
    const int n = n0 + row, k = k0 + kq;
    if (VEC) {
      if (n < g.N && k < g.K) {
        const float4 v = *reinterpret_cast<const float4 *>(g.B + (size_t)n * g.ldb + k);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else { r[0] = r[1] = r[2] = r[3] = 0.f; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] = (n < g.N && k + q < g.K) ? g.B[(size_t)n * g.ldb + k + q] : 0.f;
    }
  }
}

template <int BMODE>
__device__ __forceinline__ void store_b(float *Bs, int t, const float (&r)[4]) {
  if (BMODE == 0) {
    const int kk = t >> 4, nq = (t & 15) * 4;
    *reinterpret_cast<float4 *>(Bs + kk * LDS_ + nq) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    const int row = t >> 2, kq = (t & 3) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) Bs[(kq + q) * LDS_ + row] = r[q];
  }
}

// EPI 0: dense forward, 1: dense backward-data, 2: plain store
template <int AMODE, int BMODE, bool VEC, int EPI>
__global__ void __launch_bounds__(GT) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK * LDS_];
  __shared__ __align__(16) float Bs[2][BK * LDS_];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  const int nk = (g.K + BK - 1) / BK;
  load_a<AMODE, VEC>(g, m0, 0, t, ra);
  load_b<BMODE, VEC>(g, n0, 0, t, rb);
  store_a<AMODE>(As[0], t, ra);
  store_b<BMODE>(Bs[0], t, rb);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_a<AMODE, VEC>(g, m0, (kt + 1) * BK, t, ra);
      load_b<BMODE, VEC>(g, n0, (kt + 1) * BK, t, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[cur][k * LDS_ + ty * 4]);
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[cur][k * LDS_ + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_a<AMODE>(As[cur ^ 1], t, ra);
      store_b<BMODE>(Bs[cur ^ 1], t, rb);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= g.N) return;
  uint32_t step = 0, sample0 = 0;
  if (EPI != 2 && g.mask_on == 1) {
    step = (uint32_t)g.ctl[TN_CTL_STEP];
    sample0 = (uint32_t)g.ctl[TN_CTL_SAMPLE0];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    float mk[4] = {1.f, 1.f, 1.f, 1.f};
    if (EPI != 2) {
      if (g.mask_on == 1) {
        const Philox4 r = philox_block(g.seed, TN_RNG_DROPOUT, step, sample0 + (uint32_t)m,
                                       (uint32_t)(n >> 2));
        mk[0] = r.x < g.thr ? 1.f : 0.f;
        mk[1] = r.y < g.thr ? 1.f : 0.f;
        mk[2] = r.z < g.thr ? 1.f : 0.f;
        mk[3] = r.w < g.thr ? 1.f : 0.f;
      } else if (g.mask_on == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) mk[j] = g.mask_inj[(size_t)m * g.N + n + j];
      }
    }
    if (EPI == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < g.N) {
          const float a = act_fwd(v[j] + g.bias[n + j], g.act, g.act_nn);
          v[j] = g.mask_on ? a * mk[j] : a;
          if (g.scale != 1.f) v[j] *= g.scale;
        }
      }
    } else if (EPI == 1) {
      if (g.aux) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < g.N) {
            const float d = act_bwd_from_out(g.aux[(size_t)m * g.N + n + j], g.act, g.act_nn);
            v[j] = (g.mask_on ? v[j] * mk[j] : v[j]) * d;
          }
        }
      }
    }
    float *c = g.C + (size_t)m * g.ldc + n;
    if (VEC && n + 3 < g.N) {
      *reinterpret_cast<float4 *>(c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) c[j] = v[j];
    }
  }
}

// db[n] = sum_b g[b,n]; 32 columns x 32 row-slices per CTA, slices combined in a fixed order
__global__ void __launch_bounds__(1024) colsum_kernel(const float *__restrict__ g,
                                                      float *__restrict__ db, int B, int N) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int b = ty; b < B; b += 32) s += g[(size_t)b * N + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float r = red[0][tx];
#pragma unroll
    for (int q = 1; q < 32; ++q) r += red[q][tx];
    db[n] = r;
  }
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

int g_dense_mode = 0;
static bool use_tc(int n_in, int n_out, const void *a, const void *b, const void *c) {
  return g_dense_mode != 1 && dense_tc_ok(n_in, n_out, a, b, c);
}
// 0: one TF32 pass, 1: 3xTF32 cluster split-K (default), 2: first-generation 3xTF32 kernel
static int tc_split() { return g_dense_mode == 2 ? 0 : (g_dense_mode == 4 ? 2 : 1); }

template <int AMODE, int BMODE, int EPI>
static int launch_gemm(const GemmArgs &g, bool vec, const char *name, cudaStream_t st) {
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM));
  if (vec) sgemm_kernel<AMODE, BMODE, true, EPI><<<grid, GT, 0, st>>>(g);
  else sgemm_kernel<AMODE, BMODE, false, EPI><<<grid, GT, 0, st>>>(g);
  TN_LAUNCH_CHECK(name);
  return TN_OK;
}

}  // namespace tn

using namespace tn;

static int mask_mode(double pkeep, const float *mask_inj) {
  if (mask_inj) return 2;
  return pkeep < 1.0 ? 1 : 0;
}

extern "C" int tn_dense_fwd(const float *x, const float *W, const float *bias, float *out, int B,
                            int n_in, int n_out, int act, int act_nn, double pkeep, uint64_t seed,
                            const int32_t *ctl, const float *mask_inj, float out_scale,
                            void *stream) {
  TN_REQUIRE(x && W && bias && out, TN_ERR_ARG, "tn_dense_fwd: null argument");
  TN_REQUIRE(B > 0 && n_in > 0 && n_out > 0, TN_ERR_SHAPE, "tn_dense_fwd: bad shape");
  TN_REQUIRE(mask_mode(pkeep, mask_inj) != 1 || ctl, TN_ERR_ARG, "tn_dense_fwd: dropout needs ctl");
  if (dense_small_ok(n_in, n_out)) {
    SmallArgs a{x, W, bias, nullptr, mask_inj, out, ctl, seed, bernoulli_threshold(pkeep),
                mask_mode(pkeep, mask_inj), act, (float)act_nn, out_scale, B, n_in, n_out};
    return dense_fwd_small(a, (cudaStream_t)stream);
  }
  if (use_tc(n_in, n_out, x, W, out))
    return dense_tc_fwd(x, W, bias, out, B, n_in, n_out, act, (float)act_nn,
                        mask_mode(pkeep, mask_inj), bernoulli_threshold(pkeep), seed, ctl,
                        mask_inj, out_scale, tc_split(), 0, (cudaStream_t)stream);
  GemmArgs g{};
  g.A = x; g.B = W; g.C = out;
  g.M = B; g.N = n_out; g.K = n_in; g.lda = n_in; g.ldb = n_out; g.ldc = n_out;
  g.bias = bias; g.mask_inj = mask_inj; g.ctl = ctl; g.seed = seed;
  g.mask_on = mask_mode(pkeep, mask_inj);
  g.thr = bernoulli_threshold(pkeep);
  g.act = act; g.act_nn = (float)act_nn; g.scale = out_scale;
  const bool vec = n_in % 4 == 0 && n_out % 4 == 0 && aligned16(x) && aligned16(W) && aligned16(out);
  return launch_gemm<0, 0, 0>(g, vec, "tn_dense_fwd", (cudaStream_t)stream);
}

extern "C" int tn_dense_bwd_data_sm(const float *gr, const float *W, float *dx, int B, int n_in,
                                    int n_out, const float *prev_out, int act_prev, int nn_prev,
                                    double pkeep_prev, uint64_t seed_prev, const int32_t *ctl,
                                    const float *mask_inj_prev, int max_sms, void *stream) {
  TN_REQUIRE(gr && W && dx, TN_ERR_ARG, "tn_dense_bwd_data: null argument");
  TN_REQUIRE(B > 0 && n_in > 0 && n_out > 0, TN_ERR_SHAPE, "tn_dense_bwd_data: bad shape");
  if (dense_small_ok(n_in, n_out)) {
    const int mo = prev_out ? mask_mode(pkeep_prev, mask_inj_prev) : 0;
    TN_REQUIRE(mo != 1 || ctl, TN_ERR_ARG, "tn_dense_bwd_data: dropout needs ctl");
    SmallArgs a{gr, W, nullptr, prev_out, mask_inj_prev, dx, ctl, seed_prev,
                bernoulli_threshold(pkeep_prev), mo, act_prev, (float)nn_prev, 1.f, B, n_in, n_out};
    return dense_bwd_data_small(a, (cudaStream_t)stream);
  }
  if (use_tc(n_in, n_out, gr, W, dx)) {
    const int mo = prev_out ? mask_mode(pkeep_prev, mask_inj_prev) : 0;
    TN_REQUIRE(mo != 1 || ctl, TN_ERR_ARG, "tn_dense_bwd_data: dropout needs ctl");
    return dense_tc_bwd_data(gr, W, dx, B, n_in, n_out, prev_out, act_prev, (float)nn_prev, mo,
                             bernoulli_threshold(pkeep_prev), seed_prev, ctl, mask_inj_prev,
                             tc_split(), max_sms, (cudaStream_t)stream);
  }
  GemmArgs g{};
  g.A = gr; g.B = W; g.C = dx;
  g.M = B; g.N = n_in; g.K = n_out; g.lda = n_out; g.ldb = n_out; g.ldc = n_in;
  g.aux = prev_out; g.mask_inj = mask_inj_prev; g.ctl = ctl; g.seed = seed_prev;
  g.mask_on = prev_out ? mask_mode(pkeep_prev, mask_inj_prev) : 0;
  TN_REQUIRE(g.mask_on != 1 || ctl, TN_ERR_ARG, "tn_dense_bwd_data: dropout needs ctl");
  g.thr = bernoulli_threshold(pkeep_prev);
  g.act = act_prev; g.act_nn = (float)nn_prev; g.scale = 1.f;
  const bool vec = n_in % 4 == 0 && n_out % 4 == 0 && aligned16(gr) && aligned16(W) && aligned16(dx);
  return launch_gemm<0, 1, 1>(g, vec, "tn_dense_bwd_data", (cudaStream_t)stream);
}

extern "C" int tn_dense_bwd_data(const float *gr, const float *W, float *dx, int B, int n_in,
                                 int n_out, const float *prev_out, int act_prev, int nn_prev,
                                 double pkeep_prev, uint64_t seed_prev, const int32_t *ctl,
                                 const float *mask_inj_prev, void *stream) {
  return tn_dense_bwd_data_sm(gr, W, dx, B, n_in, n_out, prev_out, act_prev, nn_prev, pkeep_prev,
                              seed_prev, ctl, mask_inj_prev, 0, stream);
}

extern "C" int tn_dense_bwd_weights_sm(const float *x, const float *gr, float *dW, float *db, int B,
                                       int n_in, int n_out, int max_sms, void *stream) {
  TN_REQUIRE(x && gr && dW && db, TN_ERR_ARG, "tn_dense_bwd_weights: null argument");
  TN_REQUIRE(B > 0 && n_in > 0 && n_out > 0, TN_ERR_SHAPE, "tn_dense_bwd_weights: bad shape");
  if (dense_small_ok(n_in, n_out))
    return dense_bwd_weights_small(x, gr, dW, db, B, n_in, n_out, (cudaStream_t)stream);
  if (use_tc(n_in, n_out, x, gr, dW)) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dense_tc_bwd_weights(x, gr, dW, B, n_in, n_out, tc_split(), max_sms, st);
    if (rc) return rc;
    launch_pdl(colsum_kernel, dim3(ceil_div(n_out, 32)), dim3(1024), 0, st, gr, db, B, n_out);
    TN_LAUNCH_CHECK("tn_dense_bwd_weights(db)");
    return TN_OK;
  }
  GemmArgs g{};
  g.A = x; g.B = gr; g.C = dW;
  g.M = n_in; g.N = n_out; g.K = B; g.lda = n_in; g.ldb = n_out; g.ldc = n_out;
  g.scale = 1.f;
  const bool vec = n_in % 4 == 0 && n_out % 4 == 0 && aligned16(x) && aligned16(gr) && aligned16(dW);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_gemm<1, 0, 2>(g, vec, "tn_dense_bwd_weights", st);
  if (rc) return rc;
  launch_pdl(colsum_kernel, dim3(ceil_div(n_out, 32)), dim3(1024), 0, st, gr, db, B, n_out);
  TN_LAUNCH_CHECK("tn_dense_bwd_weights(db)");
  return TN_OK;
}

extern "C" int tn_dense_bwd_weights(const float *x, const float *gr, float *dW, float *db, int B,
                                    int n_in, int n_out, void *stream) {
  return tn_dense_bwd_weights_sm(x, gr, dW, db, B, n_in, n_out, 0, stream);
}

// Debug aid (tools/gemm_phase_times.py): when buf != NULL, every CTA of the cluster split-K kernel
// stores clock64() at its phase boundaries into buf[cta * 16 + slot]
extern "C" int tn_dense_debug_timestamps(long long *buf) {
  dense_tc_set_debug(buf);
  return TN_OK;
}

extern "C" int tn_set_dense_mode(int mode) {
  TN_REQUIRE(mode >= 0 && mode <= 4, TN_ERR_ARG, "tn_set_dense_mode: mode must be 0..4");
  g_dense_mode = mode;
  return TN_OK;
}
