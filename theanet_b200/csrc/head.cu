// The classifier head of theanet's shipped networks: HiddenLayer -> SoftmaxLayer with a narrow
// output (params/mnist.prms:28-39: 500 -> 10).  (theanet/layer/outlayers.py:50-51,83-102,
// hidden.py:30-32; gradients via tt.grad, layer.py:83.)
//
// As GEMMs these are N = 10 problems, far below any MMA tile, and the whole head moves ~2 MB: it
// is launch- and latency-bound, so it is fused into two launches:
//
//   tn_softmax_head_fwd_bwd     one warp per sample row: scores = h.W + b, log-softmax, NLL term,
//                               g = (softmax - onehot)/B, and dL/dh = g.W^T turned directly into
//                               dL/dz of the hidden layer below (dropout mask * act')
//   tn_softmax_head_bwd_weights dW = h^T.g, db = column sums of g: (input tile, batch chunk) CTAs,
//                               the last CTA of every tile (ticket) adds the chunk partials in a
//                               fixed order -- deterministic, no second launch
//
// W is transposed into shared memory once per CTA so that lanes read consecutive float4; each
// lane owns quads of 4 consecutive inputs (one Philox block per quad, as everywhere else).
#include "common.cuh"

namespace tn {

struct HeadArgs {
  const float *h;         // (B, n_in): the hidden layer's stored (masked) output
  const float *W;         // (n_in, n_out)
  const float *bias;      // (n_out)
  const int32_t *y, *idx, *ctl;
  float *logprob, *g, *rowloss;  // (B, n_out), (B, n_out), (B)
  float *dh;              // (B, n_in) or null
  const float *mask_inj;  // (B, n_in) or null
  uint64_t seed;
  uint32_t thr;
  int B, n_in, n_out;
  int below;              // 1: multiply dh by mask * act'(h) of the layer below
  int mask_on;
  ActK ak;
  float inv_bg;
};

template <int NP, int NQ>
__global__ void __launch_bounds__(256) softmax_head_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LDK = NQ * 128;
  float *ws = sm;              // [n_out][LDK], zero padded
  float *bs = sm + NP * LDK;   // [NP]
  pdl_trigger();
  pdl_wait();
  for (int t = threadIdx.x; t < a.n_out * LDK; t += blockDim.x) {
    const int j = t / LDK, k = t % LDK;
    ws[t] = k < a.n_in ? a.W[(size_t)k * a.n_out + j] : 0.f;
  }
  if (threadIdx.x < NP) bs[threadIdx.x] = threadIdx.x < a.n_out ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const float4 *ws4 = reinterpret_cast<const float4 *>(ws);
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (a.n_in & 3) == 0;
  uint32_t step = 0, sample0 = 0;
  if (a.below && a.mask_on == 1) {
    step = (uint32_t)a.ctl[TN_CTL_STEP];
    sample0 = (uint32_t)a.ctl[TN_CTL_SAMPLE0];
  }
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.B; b += nwarps) {
    const float *hr = a.h + (size_t)b * a.n_in;
    float4 hq[NQ];
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int k0 = 4 * (lane + 32 * t);
      if (vec && k0 < a.n_in) {
        hq[t] = *reinterpret_cast<const float4 *>(hr + k0);
      } else {
        hq[t].x = k0 < a.n_in ? hr[k0] : 0.f;
        hq[t].y = k0 + 1 < a.n_in ? hr[k0 + 1] : 0.f;
        hq[t].z = k0 + 2 < a.n_in ? hr[k0 + 2] : 0.f;
        hq[t].w = k0 + 3 < a.n_in ? hr[k0 + 3] : 0.f;
      }
    }
    float z[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float s = 0.f;
      if (j < a.n_out) {
#pragma unroll
        for (int t = 0; t < NQ; ++t) {
          const float4 w = ws4[j * (LDK / 4) + lane + 32 * t];
          s = fmaf(hq[t].x, w.x, s);
          s = fmaf(hq[t].y, w.y, s);
          s = fmaf(hq[t].z, w.z, s);
          s = fmaf(hq[t].w, w.w, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += bs[j];
      }
      z[j] = s;
    }
    // log-softmax over the n_out scores (every lane holds all of them)
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (j < a.n_out) m = fmaxf(m, z[j]);
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (j < a.n_out) se += expf(z[j] - m);
    const float ls = logf(se);
    const int64_t row = a.idx ? (int64_t)a.idx[b] : (int64_t)a.ctl[TN_CTL_ROW0] + b;
    const int label = a.y[row];
    float gj[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      gj[j] = 0.f;
      if (j < a.n_out) {
        const float lp = (z[j] - m) - ls;
        float p = expf(lp);
        if (j == label) p -= 1.f;
        gj[j] = p * a.inv_bg;
        if (lane == j) {
          a.logprob[(size_t)b * a.n_out + j] = lp;
          a.g[(size_t)b * a.n_out + j] = gj[j];
          if (j == label) a.rowloss[b] = -lp;
        }
      }
    }
    if (!a.dh) continue;
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int k0 = 4 * (lane + 32 * t);
      if (k0 >= a.n_in) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (j < a.n_out) {
          const float4 w = ws4[j * (LDK / 4) + lane + 32 * t];
          acc.x = fmaf(gj[j], w.x, acc.x);
          acc.y = fmaf(gj[j], w.y, acc.y);
          acc.z = fmaf(gj[j], w.z, acc.z);
          acc.w = fmaf(gj[j], w.w, acc.w);
        }
      }
      float v[4] = {acc.x, acc.y, acc.z, acc.w};
      if (a.below) {
        const float hv[4] = {hq[t].x, hq[t].y, hq[t].z, hq[t].w};
        float mk[4] = {1.f, 1.f, 1.f, 1.f};
        if (a.mask_on == 1) {
          const Philox4 r = philox_block(a.seed, TN_RNG_DROPOUT, step, sample0 + (uint32_t)b,
                                         (uint32_t)(lane + 32 * t));
          mk[0] = r.x < a.thr ? 1.f : 0.f;
          mk[1] = r.y < a.thr ? 1.f : 0.f;
          mk[2] = r.z < a.thr ? 1.f : 0.f;
          mk[3] = r.w < a.thr ? 1.f : 0.f;
        } else if (a.mask_on == 2) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (k0 + e < a.n_in) mk[e] = a.mask_inj[(size_t)b * a.n_in + k0 + e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = act_bwd_k(a.ak, hv[e]);
          v[e] = (a.mask_on ? v[e] * mk[e] : v[e]) * d;
        }
      }
      float *o = a.dh + (size_t)b * a.n_in + k0;
      if (vec) {
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + e < a.n_in) o[e] = v[e];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int kHeadChunk = 128;  // batch rows per CTA

// workspace: partial dW [nchunks][n_in][n_out] | partial db [nchunks][n_out] | tickets [ntiles]
template <int NP>
__global__ void __launch_bounds__(256)
head_dw_kernel(const float *__restrict__ h, const float *__restrict__ g, float *__restrict__ dW,
               float *__restrict__ db, float *ws_f, int B, int n_in, int n_out,
               const float *__restrict__ rowloss, float *__restrict__ nll_sum) {
  __shared__ float red[8][NP][33];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tile = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  const int i = tile * 32 + lane;
  pdl_trigger();
  pdl_wait();
  float *part_w = ws_f;
  float *part_b = ws_f + (size_t)nchunks * n_in * n_out;
  int *tickets = reinterpret_cast<int *>(part_b + (size_t)nchunks * n_out);
  float acc[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) acc[j] = 0.f;
  float dbacc = 0.f;
  const int r1 = min(B, (chunk + 1) * kHeadChunk);
  for (int r = chunk * kHeadChunk + w; r < r1; r += 8) {
    const float hv = i < n_in ? h[(size_t)r * n_in + i] : 0.f;
    const float gv = lane < n_out ? g[(size_t)r * n_out + lane] : 0.f;
    dbacc += gv;
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (j < n_out) acc[j] = fmaf(hv, __shfl_sync(0xffffffffu, gv, j), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) red[w][j][lane] = acc[j];
  __syncthreads();
  for (int t = threadIdx.x; t < n_out * 32; t += 256) {
    const int l = t / n_out, j = t - l * n_out;   // consecutive threads -> consecutive j: coalesced
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][j][l];
    const int ii = tile * 32 + l;
    if (ii < n_in) part_w[((size_t)chunk * n_in + ii) * n_out + j] = s;
  }
  if (tile == 0) {
    __syncthreads();
    red[w][0][lane] = dbacc;
    __syncthreads();
    if (threadIdx.x < n_out) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][0][threadIdx.x];
      part_b[(size_t)chunk * n_out + threadIdx.x] = s;
    }
  }
  // ticket: the last CTA of this tile to finish adds the chunk partials in chunk order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(&tickets[tile], 1);
    s_last = t == nchunks - 1;
    if (s_last) tickets[tile] = 0;  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int t = threadIdx.x; t < n_out * 32; t += 256) {
    const int l = t / n_out, j = t - l * n_out;
    const int ii = tile * 32 + l;
    if (ii >= n_in) continue;
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += __ldcg(part_w + ((size_t)c * n_in + ii) * n_out + j);
    dW[(size_t)ii * n_out + j] = s;
  }
  if (tile == 0 && threadIdx.x < n_out) {
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += __ldcg(part_b + (size_t)c * n_out + threadIdx.x);
    db[threadIdx.x] = s;
  }
  if (tile == 0 && rowloss) {   // nll_sum = sum_b rowloss[b]: strided partials, fixed-order tree
    __syncthreads();
    float *r1 = &red[0][0][0];
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += 256) s += rowloss[b];
    r1[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) r1[threadIdx.x] += r1[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) nll_sum[0] = r1[0];
  }
}

template <typename K>
static int head_smem(K kernel, size_t smem, const char *who) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  }
  return TN_OK;
}

template <int NP, int NQ>
static int launch_head(const HeadArgs &a, cudaStream_t st) {
  const char *who = "tn_softmax_head_fwd_bwd";
  const size_t smem = ((size_t)NP * NQ * 128 + NP) * sizeof(float);
  int rc = head_smem(softmax_head_kernel<NP, NQ>, smem, who);
  if (rc) return rc;
  const int blocks = min(ceil_div(a.B, 8), kNumSM);
  launch_pdl(softmax_head_kernel<NP, NQ>, dim3(blocks), dim3(256), smem, st, a);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

template <int NP>
static int launch_head_q(const HeadArgs &a, cudaStream_t st) {
  const int nq = ceil_div(a.n_in, 128);
  if (nq <= 2) return launch_head<NP, 2>(a, st);
  if (nq <= 4) return launch_head<NP, 4>(a, st);
  return launch_head<NP, 8>(a, st);
}

}  // namespace tn

using namespace tn;

extern "C" int tn_softmax_head_supported(int n_in, int n_out) {
  return n_in >= 1 && n_in <= 1024 && n_out >= 1 && n_out <= 32;
}

extern "C" int tn_softmax_head_fwd_bwd(const float *h, const float *W, const float *bias,
                                       const int32_t *y, const int32_t *idx, const int32_t *ctl,
                                       int B, int n_in, int n_out, float inv_global_batch,
                                       float *logprob, float *g, float *rowloss, float *dh,
                                       int below, int act_below, int nn_below, double pkeep_below,
                                       uint64_t seed_below, const float *mask_inj_below,
                                       void *stream) {
  const char *who = "tn_softmax_head_fwd_bwd";
  TN_REQUIRE(h && W && bias && y && logprob && g && rowloss && (idx || ctl), TN_ERR_ARG,
             "%s: null argument", who);
  TN_REQUIRE(B > 0 && tn_softmax_head_supported(n_in, n_out), TN_ERR_UNSUPPORTED,
             "%s: needs n_in <= 1024 and n_out <= 32 (got %d, %d); use tn_dense_fwd + "
             "tn_softmax_nll_fwd_bwd", who, n_in, n_out);
  HeadArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.idx = idx; a.ctl = ctl;
  a.logprob = logprob; a.g = g; a.rowloss = rowloss; a.dh = dh; a.mask_inj = mask_inj_below;
  a.seed = seed_below; a.thr = bernoulli_threshold(pkeep_below);
  a.B = B; a.n_in = n_in; a.n_out = n_out; a.below = below && dh;
  a.mask_on = !a.below ? 0 : (mask_inj_below ? 2 : (pkeep_below < 1.0 ? 1 : 0));
  TN_REQUIRE(a.mask_on != 1 || ctl, TN_ERR_ARG, "%s: dropout needs ctl", who);
  a.ak = make_actk(act_below, nn_below); a.inv_bg = inv_global_batch;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_out <= 8) return launch_head_q<8>(a, st);
  if (n_out <= 12) return launch_head_q<12>(a, st);
  if (n_out <= 16) return launch_head_q<16>(a, st);
  return launch_head_q<32>(a, st);
}

extern "C" size_t tn_softmax_head_workspace_bytes(int B, int n_in, int n_out) {
  const size_t nchunks = (size_t)ceil_div(B, kHeadChunk);
  return (nchunks * n_in * n_out + nchunks * n_out) * sizeof(float) +
         (size_t)ceil_div(n_in, 32) * sizeof(int);
}

extern "C" int tn_softmax_head_bwd_weights(const float *h, const float *g, float *dW, float *db,
                                           void *workspace, int B, int n_in, int n_out,
                                           const float *rowloss, float *nll_sum, void *stream) {
  const char *who = "tn_softmax_head_bwd_weights";
  TN_REQUIRE(h && g && dW && db && workspace && (!rowloss || nll_sum), TN_ERR_ARG,
             "%s: null argument", who);
  TN_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_out <= 32, TN_ERR_UNSUPPORTED,
             "%s: needs n_out <= 32 (got %d)", who, n_out);
  dim3 grid(ceil_div(n_in, 32), ceil_div(B, kHeadChunk));
  cudaStream_t st = (cudaStream_t)stream;
  float *ws = (float *)workspace;
  if (n_out <= 8) launch_pdl(head_dw_kernel<8>, grid, dim3(256), 0, st, h, g, dW, db, ws, B, n_in, n_out, rowloss, nll_sum);
  else if (n_out <= 12) launch_pdl(head_dw_kernel<12>, grid, dim3(256), 0, st, h, g, dW, db, ws, B, n_in, n_out, rowloss, nll_sum);
  else if (n_out <= 16) launch_pdl(head_dw_kernel<16>, grid, dim3(256), 0, st, h, g, dW, db, ws, B, n_in, n_out, rowloss, nll_sum);
  else launch_pdl(head_dw_kernel<32>, grid, dim3(256), 0, st, h, g, dW, db, ws, B, n_in, n_out, rowloss, nll_sum);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}
