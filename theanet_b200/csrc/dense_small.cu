// Dense layers with a narrow output (n_out <= 32): the SoftmaxLayer of the shipped networks is
// 500 -> 10 (params/mnist.prms:36).  As GEMMs these are N=10 / K=10 problems -- below any MMA
// tile -- and pure bandwidth work: each kernel streams the wide operand once, keeps the weight
// matrix in shared memory and reduces with warp shuffles.  (theanet/layer/hidden.py:30-32.)
#include "common.cuh"
#include "dense_small.cuh"

namespace tn {

// out[b, n] = epi(sum_k x[b,k] W[k,n]); one warp per row, lanes split k, W in smem [k][ld]
__global__ void __launch_bounds__(256) dense_fwd_small_kernel(SmallArgs a) {
  extern __shared__ float ws[];
  const int ld = a.n_out | 1;
  for (int t = threadIdx.x; t < a.n_in * a.n_out; t += blockDim.x) {
    const int k = t / a.n_out, n = t - k * a.n_out;
    ws[k * ld + n] = a.W[t];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  uint32_t step = 0, sample0 = 0;
  if (a.mask_on == 1) {
    step = (uint32_t)a.ctl[TN_CTL_STEP];
    sample0 = (uint32_t)a.ctl[TN_CTL_SAMPLE0];
  }
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.B; b += warps) {
    float acc[kSmallN];
#pragma unroll
    for (int n = 0; n < kSmallN; ++n) acc[n] = 0.f;
    const float *xr = a.x + (size_t)b * a.n_in;
    for (int k = lane; k < a.n_in; k += 32) {
      const float xv = xr[k];
      const float *wr = ws + k * ld;
#pragma unroll
      for (int n = 0; n < kSmallN; ++n)
        if (n < a.n_out) acc[n] = fmaf(xv, wr[n], acc[n]);
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < kSmallN; ++n) {
      if (n < a.n_out) {
        float v = acc[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == n) mine = v;
      }
    }
    if (lane < a.n_out) {
      float v = act_fwd(mine + a.bias[lane], a.act, a.act_nn);
      if (a.mask_on == 1) {
        const Philox4 r = philox_block(a.seed, TN_RNG_DROPOUT, step, sample0 + (uint32_t)b,
                                       (uint32_t)(lane >> 2));
        v *= philox_word(r, lane & 3) < a.thr ? 1.f : 0.f;
      } else if (a.mask_on == 2) {
        v *= a.mask_inj[(size_t)b * a.n_out + lane];
      }
      if (a.scale != 1.f) v *= a.scale;
      a.out[(size_t)b * a.n_out + lane] = v;
    }
  }
}

// dx[b, i] = (sum_j g[b,j] W[i,j]) [* mask * act'(prev_out)]; one warp per row, each lane owns
// quads of 4 consecutive i; W transposed in smem: wt[j][ldi], i contiguous
__global__ void __launch_bounds__(256) dense_bwd_data_small_kernel(SmallArgs a) {
  extern __shared__ __align__(16) float wt[];
  const int ldi = (a.n_in + 3) & ~3;
  for (int t = threadIdx.x; t < a.n_out * ldi; t += blockDim.x) {
    const int j = t / ldi, i = t - j * ldi;
    wt[t] = i < a.n_in ? a.W[(size_t)i * a.n_out + j] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nquads = ldi >> 2;
  uint32_t step = 0, sample0 = 0;
  if (a.mask_on == 1) {
    step = (uint32_t)a.ctl[TN_CTL_STEP];
    sample0 = (uint32_t)a.ctl[TN_CTL_SAMPLE0];
  }
  const bool vec = (a.n_in & 3) == 0;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.B; b += warps) {
    const float gv = lane < a.n_out ? a.x[(size_t)b * a.n_out + lane] : 0.f;
    // warp-uniform trip count: every lane takes part in the shuffles, idle lanes redo the last quad
    for (int qb = 0; qb < nquads; qb += 32) {
      const bool live = qb + lane < nquads;
      const int q = live ? qb + lane : nquads - 1;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < a.n_out; ++j) {
        const float gj = __shfl_sync(0xffffffffu, gv, j);
        const float4 w = *reinterpret_cast<const float4 *>(wt + j * ldi + 4 * q);
        acc.x = fmaf(gj, w.x, acc.x);
        acc.y = fmaf(gj, w.y, acc.y);
        acc.z = fmaf(gj, w.z, acc.z);
        acc.w = fmaf(gj, w.w, acc.w);
      }
      float v[4] = {acc.x, acc.y, acc.z, acc.w};
      const int i0 = 4 * q;
      if (a.aux) {
        float mk[4] = {1.f, 1.f, 1.f, 1.f};
        if (a.mask_on == 1) {
          const Philox4 r = philox_block(a.seed, TN_RNG_DROPOUT, step, sample0 + (uint32_t)b,
                                         (uint32_t)q);
          mk[0] = r.x < a.thr ? 1.f : 0.f;
          mk[1] = r.y < a.thr ? 1.f : 0.f;
          mk[2] = r.z < a.thr ? 1.f : 0.f;
          mk[3] = r.w < a.thr ? 1.f : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (i0 + r < a.n_in) {
            const size_t o = (size_t)b * a.n_in + i0 + r;
            if (a.mask_on == 2) mk[r] = a.mask_inj[o];
            const float d = act_bwd_from_out(a.aux[o], a.act, a.act_nn);
            v[r] = (a.mask_on ? v[r] * mk[r] : v[r]) * d;
          }
        }
      }
      if (!live) continue;
      float *o = a.out + (size_t)b * a.n_in + i0;
      if (vec) {
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (i0 + r < a.n_in) o[r] = v[r];
      }
    }
  }
}

// dW[i, j] = sum_b x[b,i] g[b,j], db[j] = sum_b g[b,j].  One CTA per tile of 32 inputs i (lanes);
// its 16 warps split the batch, then combine through shared memory in a fixed order.
__global__ void __launch_bounds__(512) dense_bwd_weights_small_kernel(
    const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ dW,
    float *__restrict__ db, int B, int n_in, int n_out) {
  extern __shared__ float red[];  // [16 warps][n_out][33]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc[kSmallN];
#pragma unroll
  for (int n = 0; n < kSmallN; ++n) acc[n] = 0.f;
  float dbacc = 0.f;
  const bool ok = i < n_in;
  for (int b = w; b < B; b += 16) {
    const float xv = ok ? x[(size_t)b * n_in + i] : 0.f;
    const float gv = lane < n_out ? g[(size_t)b * n_out + lane] : 0.f;
    dbacc += gv;
#pragma unroll
    for (int n = 0; n < kSmallN; ++n)
      if (n < n_out) acc[n] = fmaf(xv, __shfl_sync(0xffffffffu, gv, n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < kSmallN; ++n)
    if (n < n_out) red[(w * n_out + n) * 33 + lane] = acc[n];
  __syncthreads();
  // thread t -> (n, lane) pairs; sum over the 16 warps in order
  for (int t = threadIdx.x; t < n_out * 32; t += blockDim.x) {
    const int n = t >> 5, l = t & 31;
    float s = 0.f;
    for (int q = 0; q < 16; ++q) s += red[(q * n_out + n) * 33 + l];
    const int ii = blockIdx.x * 32 + l;
    if (ii < n_in) dW[(size_t)ii * n_out + n] = s;
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    if (lane < n_out) red[w * 33 + lane] = dbacc;
    __syncthreads();
    if (threadIdx.x < n_out) {
      float s = 0.f;
      for (int q = 0; q < 16; ++q) s += red[q * 33 + threadIdx.x];
      db[threadIdx.x] = s;
    }
  }
}

static size_t smem_fwd(int n_in, int n_out) { return (size_t)n_in * (n_out | 1) * sizeof(float); }
static size_t smem_bwd(int n_in, int n_out) {
  return (size_t)n_out * ((n_in + 3) & ~3) * sizeof(float);
}
constexpr size_t kSmallSmemMax = 160 * 1024;

bool dense_small_ok(int n_in, int n_out) {
  return n_out <= kSmallN && smem_fwd(n_in, n_out) <= kSmallSmemMax &&
         smem_bwd(n_in, n_out) <= kSmallSmemMax;
}

template <typename K>
static int opt_in_smem(K kernel, size_t smem, const char *name) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
  return TN_OK;
}

int dense_fwd_small(const SmallArgs &a, cudaStream_t st) {
  const size_t smem = smem_fwd(a.n_in, a.n_out);
  int rc = opt_in_smem(dense_fwd_small_kernel, smem, "tn_dense_fwd");
  if (rc) return rc;
  const int blocks = min(ceil_div(a.B, 8), 2 * kNumSM);
  dense_fwd_small_kernel<<<blocks, 256, smem, st>>>(a);
  TN_LAUNCH_CHECK("tn_dense_fwd(small)");
  return TN_OK;
}

int dense_bwd_data_small(const SmallArgs &a, cudaStream_t st) {
  const size_t smem = smem_bwd(a.n_in, a.n_out);
  int rc = opt_in_smem(dense_bwd_data_small_kernel, smem, "tn_dense_bwd_data");
  if (rc) return rc;
  const int blocks = min(ceil_div(a.B, 8), 2 * kNumSM);
  dense_bwd_data_small_kernel<<<blocks, 256, smem, st>>>(a);
  TN_LAUNCH_CHECK("tn_dense_bwd_data(small)");
  return TN_OK;
}

int dense_bwd_weights_small(const float *x, const float *g, float *dW, float *db, int B, int n_in,
                            int n_out, cudaStream_t st) {
  const size_t smem = (size_t)16 * n_out * 33 * sizeof(float);
  int rc = opt_in_smem(dense_bwd_weights_small_kernel, smem, "tn_dense_bwd_weights");
  if (rc) return rc;
  dense_bwd_weights_small_kernel<<<ceil_div(n_in, 32), 512, smem, st>>>(x, g, dW, db, B, n_in, n_out);
  TN_LAUNCH_CHECK("tn_dense_bwd_weights(small)");
  return TN_OK;
}

}  // namespace tn
