// Layer.get_updates / get_wtcost (theanet/layer/layer.py:70-117) for ALL parameter tensors in one
// pass over the flat theta / velocity / gradient buffers:
//     g'     = g*grad_scale + L1*sgn(theta) + 2*L2*theta
//     v'     = m*v + (1-m)*g'
//     theta' = theta - rate*lr*v          <- the OLD velocity (Theano updates are simultaneous)
//     maxnorm: 1-D clip; 2-D per-column (axis 0) norm; 4-D per-output-kernel norm
//     cost   = nll + sum_layers L1*sum|theta| + L2*sum theta^2   (pre-update theta)
// Pure HBM work: 3 reads + 2 writes of 4 B per parameter.
#include "common.cuh"

namespace tn {

constexpr int kMaxSegs = 64;
constexpr int kUpdThreads = 256;
constexpr int kUpdPerBlock = kUpdThreads * 4;

// Fused data-parallel step: instead of an NCCL all-reduce followed by the update,
// every rank reads the gradient buffers of ALL ranks directly over NVLink (CUDA-IPC mapped peer
// memory) and adds them in rank order -- the same order everywhere, so the replicas stay bit
// identical -- inside the optimiser kernel itself.  Handshake: each rank stores a step token into
// every peer's flag array when its own gradients are complete (it is the first thing this kernel
// does, after all backward kernels in stream order) and spins until all tokens have arrived.  The
// token is a per-rank EXECUTION counter kept on the device (word kMaxPeers of the rank's own flag
// array, advanced by the last CTA of every execution): ranks run the kernel in lockstep, so their
// counters agree, and re-running a step (graph warm-up followed by its replay, a resumed run that
// rewinds the step counter) can never find last time's token already in place.
// Gradient buffers are double-buffered by step parity, which makes the "peers are done reading"
// direction implicit: a buffer is rewritten two steps later, after its owner has passed the next
// step's handshake, which every reader only enters once it has finished this step's reads.
constexpr int kMaxPeers = 8;
struct PeerArgs {
  const float *grad[kMaxPeers];   // grad[r]: rank r's gradient buffer of this parity (r == rank: local)
  int *flags[kMaxPeers];          // flags[r]: rank r's flag array (int[2*kMaxPeers]); [rank] is local:
                                  // words 0..7 tokens by source rank, word 8 the execution counter
  int world, rank;
  int64_t peer_end;               // flat indices < peer_end are summed over the ranks here; the rest of
                                  // the buffer (and the NLL slot) was all-reduced before this kernel
};

__device__ __forceinline__ void st_release_sys(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct SegTable {
  tn_param_seg seg[kMaxSegs];
  int first_block[kMaxSegs + 1];
  int nseg;
};

__global__ void __launch_bounds__(kUpdThreads)
sgd_step_kernel(float *__restrict__ theta, float *__restrict__ vel, const float *__restrict__ grad,
                const __grid_constant__ SegTable tab, const int32_t *__restrict__ ctl,
                float grad_scale, float *__restrict__ wt_partial, int *ticket,
                const float *__restrict__ nll_sum, float nll_scale, float *__restrict__ cost_out,
                const __grid_constant__ PeerArgs peers, int64_t nll_slot) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kUpdThreads / 32];
  __shared__ float red2[kUpdThreads];
  __shared__ int s_last;
  const int W = peers.world;
  int s = 0;
  while (s + 1 < tab.nseg && (int)blockIdx.x >= tab.first_block[s + 1]) ++s;
  const tn_param_seg &sg = tab.seg[s];
  const int64_t base = sg.offset + (int64_t)(blockIdx.x - tab.first_block[s]) * kUpdPerBlock;
  const bool from_peers = W > 1 && base < peers.peer_end;     // segments never straddle peer_end
  int token = 0;
  if (W > 1) {
    token = __ldcg(peers.flags[peers.rank] + kMaxPeers) + 1;
    if (blockIdx.x == 0 && threadIdx.x < W) {
      __threadfence_system();
      st_release_sys(peers.flags[threadIdx.x] + peers.rank, token);   // "my gradients are complete"
    }
    // only the CTAs that read peer memory wait for the peers: with peer_end < total the bulk of
    // the buffer was reduced earlier and its CTAs go straight to work, hiding the handshake
    if (from_peers) {
      if (threadIdx.x < W) {
        const int *f = peers.flags[peers.rank] + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) - token < 0) {
          if (clock64() - t0 > 20000000000ll) __trap();   // ~10 s: a peer died; fail loudly
        }
      }
      __syncthreads();
    }
  }
  const int64_t end = sg.offset + sg.size;
  const float lr = ctl_lr(ctl);
  const float step = __fmul_rn(sg.rate, lr);
  const float m = sg.momentum, om = __fsub_rn(1.f, m);
  const float l2x2 = 2.f * sg.l2;
  const bool clip1d = sg.ndim == 1 && sg.maxnorm != 0.f;
  float wsum = 0.f;
  // One float4 per thread (every tensor starts 16-byte aligned and is padded to a multiple of 4;
  // pad words are read but never written).
  // data parallel: the quad comes from every rank (W independent 16-byte NVLink loads in flight,
  // L1 bypassed: peer data changes every step) and is added in rank order
  const int64_t i0 = base + 4 * (int64_t)threadIdx.x;
  if (i0 < end) {
    float4 g4;
    if (from_peers) {
      float4 gp[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < W) gp[r] = __ldcg(reinterpret_cast<const float4 *>(peers.grad[r] + i0));
      g4 = gp[0];
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r)
        if (r < W) {
          g4.x = __fadd_rn(g4.x, gp[r].x); g4.y = __fadd_rn(g4.y, gp[r].y);
          g4.z = __fadd_rn(g4.z, gp[r].z); g4.w = __fadd_rn(g4.w, gp[r].w);
        }
    } else {
      g4 = *reinterpret_cast<const float4 *>(grad + i0);
    }
    const float4 th4 = *reinterpret_cast<const float4 *>(theta + i0);
    const float thv[4] = {th4.x, th4.y, th4.z, th4.w};
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (i0 + e < end) {
        if (sg.l1 != 0.f) wsum = fmaf(sg.l1, fabsf(thv[e]), wsum);
        if (sg.l2 != 0.f) wsum = fmaf(sg.l2, thv[e] * thv[e], wsum);
      }
    if (sg.rate != 0.f) {
      const float4 v4 = *reinterpret_cast<const float4 *>(vel + i0);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      float vn[4], tn_[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float th = thv[e];
        float g = grad_scale == 1.f ? gv[e] : __fmul_rn(gv[e], grad_scale);
        if (sg.l1 != 0.f) {
          const float sgn = th > 0.f ? 1.f : (th < 0.f ? -1.f : 0.f);
          g = __fadd_rn(g, __fmul_rn(sg.l1, sgn));
        }
        if (sg.l2 != 0.f) g = __fadd_rn(g, __fmul_rn(l2x2, th));
        vn[e] = __fadd_rn(__fmul_rn(m, vv[e]), __fmul_rn(om, g));
        tn_[e] = __fsub_rn(th, __fmul_rn(step, vv[e]));
        if (clip1d) tn_[e] = fminf(fmaxf(tn_[e], -sg.maxnorm), sg.maxnorm);
      }
      if (i0 + 3 < end) {
        *reinterpret_cast<float4 *>(vel + i0) = make_float4(vn[0], vn[1], vn[2], vn[3]);
        *reinterpret_cast<float4 *>(theta + i0) = make_float4(tn_[0], tn_[1], tn_[2], tn_[3]);
      } else {                                   // last quad of a tensor: its pad words stay untouched
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (i0 + e < end) {
            vel[i0 + e] = vn[e];
            theta[i0 + e] = tn_[e];
          }
      }
    }
  }
  // block partial of the weight cost (fixed-order shuffle tree)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kUpdThreads / 32; ++w) t += red[w];
    wt_partial[blockIdx.x] = t;
  }
  if (!cost_out && W <= 1) return;
  // the last CTA to finish (ticket) adds the per-CTA weight-cost partials in a fixed order:
  // cost = nll * nll_scale + L1/L2 cost of the PRE-update theta  (no second launch)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    s_last = t == (int)gridDim.x - 1;
    if (s_last) *ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // every CTA has read the execution counter (they all passed the ticket): advance it
  if (W > 1 && threadIdx.x == 0) peers.flags[peers.rank][kMaxPeers] = token;
  if (!cost_out) return;
  float csum = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kUpdThreads) csum += __ldcg(wt_partial + i);
  red2[threadIdx.x] = csum;
  __syncthreads();
  for (int o = kUpdThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red2[threadIdx.x] += red2[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float nll = 0.f;
    if (W > 1 && peers.peer_end >= nll_slot) {
      for (int r = 0; r < W; ++r) nll = __fadd_rn(nll, __ldcg(peers.grad[r] + nll_slot));
    } else if (W > 1) {
      nll = __ldcg(peers.grad[peers.rank] + nll_slot);        // already summed over the ranks
    } else if (nll_sum) {
      nll = nll_sum[0];
    }
    cost_out[0] = nll * nll_scale + red2[0];
  }
}

__device__ __forceinline__ float maxnorm_scale(float sumsq, float maxnorm) {
  const float n = sqrtf(sumsq);
  const float d = fminf(fmaxf(n, 0.f), maxnorm);
  return __fdiv_rn(__fadd_rn(1e-7f, d), __fadd_rn(1e-7f, n));
}

// 2-D (rows x cols): per-column norm over axis 0; 8 columns (one 32-byte sector per row) x 128 row
// slices per CTA: cols/8 CTAs of 1024 threads (the first version ran 32 columns x 8 slices, i.e. 32
// CTAs of 256 threads for a 4096 x 1024 matrix, and took longer than the update itself)
constexpr int kMnCols = 8, kMnSlices = 128;
__global__ void __launch_bounds__(kMnCols * kMnSlices)
maxnorm_cols_kernel(float *__restrict__ w, int rows, int cols, float maxnorm) {
  __shared__ float red[kMnSlices / 4][kMnCols];
  __shared__ float scale[kMnCols];
  const int tx = threadIdx.x & (kMnCols - 1), ty = threadIdx.x / kMnCols;
  const int c = blockIdx.x * kMnCols + tx;
  float s = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int r = ty; r < rows; r += kMnSlices) {
      const float v = w[(size_t)r * cols + c];
      s = fmaf(v, v, s);
    }
  }
  // a warp holds 4 row slices x 8 columns: fold the slices (fixed order), then the 32 warps
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if ((threadIdx.x & 31) < kMnCols) red[threadIdx.x >> 5][tx] = s;
  __syncthreads();
  if (threadIdx.x < kMnCols) {
    float t = red[0][tx];
#pragma unroll
    for (int q = 1; q < kMnSlices / 4; ++q) t += red[q][tx];
    scale[tx] = maxnorm_scale(t, maxnorm);
  }
  __syncthreads();
  if (c < cols) {
    const float sc = scale[tx];
    if (sc != 1.f)
      for (int r = ty; r < rows; r += kMnSlices) w[(size_t)r * cols + c] *= sc;
  }
}

// 4-D (rows = output kernels, cols = C*f*f contiguous): one warp per row
__global__ void maxnorm_rows_kernel(float *__restrict__ w, int rows, int cols, float maxnorm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float *row = w + (size_t)r * cols;
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s = fmaf(row[j], row[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float sc = maxnorm_scale(s, maxnorm);
  if (sc != 1.f)
    for (int j = lane; j < cols; j += 32) row[j] *= sc;
}

__global__ void finish_cost_kernel(const float *__restrict__ wt_partial, int n,
                                   const float *__restrict__ nll_sum, float nll_scale,
                                   float *__restrict__ cost_out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += wt_partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) cost_out[0] = (nll_sum ? nll_sum[0] * nll_scale : 0.f) + red[0];
}

static int64_t update_blocks(const tn_param_seg *segs, int nseg, int *first_block) {
  int64_t nb = 0;
  for (int s = 0; s < nseg; ++s) {
    if (first_block) first_block[s] = (int)nb;
    nb += ceil_div64(segs[s].size, kUpdPerBlock);
  }
  if (first_block) first_block[nseg] = (int)nb;
  return nb;
}

}  // namespace tn

using namespace tn;

extern "C" size_t tn_update_workspace_bytes(int nseg, int64_t total) {
  // one float per CTA of the step kernel (bounded by total/1024 + nseg) + the completion ticket
  return (size_t)(total / kUpdPerBlock + nseg + 1) * sizeof(float) + 16;
}

static int update_impl(float *theta, float *vel, const float *grad, const tn_param_seg *segs,
                       int nseg, int64_t total, const int32_t *ctl, float grad_scale,
                       const float *nll_sum, float nll_scale, float *cost_out, void *workspace,
                       const PeerArgs &peers, void *stream) {
  TN_REQUIRE(theta && vel && grad && segs && ctl && workspace, TN_ERR_ARG,
             "tn_sgd_momentum_maxnorm_update: null argument");
  TN_REQUIRE(nseg > 0 && nseg <= kMaxSegs, TN_ERR_UNSUPPORTED,
             "tn_sgd_momentum_maxnorm_update: %d segments (max %d)", nseg, kMaxSegs);
  SegTable tab;
  tab.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    tab.seg[s] = segs[s];
    TN_REQUIRE(segs[s].size > 0 && segs[s].offset >= 0 && segs[s].offset + segs[s].size <= total,
               TN_ERR_SHAPE, "tn_sgd_momentum_maxnorm_update: segment %d out of range", s);
    TN_REQUIRE(segs[s].ndim == 1 || (int64_t)segs[s].rows * segs[s].cols == segs[s].size,
               TN_ERR_SHAPE, "tn_sgd_momentum_maxnorm_update: segment %d rows*cols != size", s);
  }
  const int64_t nb = update_blocks(segs, nseg, tab.first_block);
  cudaStream_t st = (cudaStream_t)stream;
  float *wt_partial = (float *)workspace;
  int *ticket = reinterpret_cast<int *>(wt_partial + (total / kUpdPerBlock + nseg + 1));
  launch_pdl(sgd_step_kernel, dim3((unsigned)nb), dim3(kUpdThreads), 0, st, theta, vel, grad, tab, ctl,
             grad_scale, wt_partial, ticket, nll_sum, nll_scale, cost_out, peers, total);
  TN_LAUNCH_CHECK("tn_sgd_momentum_maxnorm_update(step)");
  for (int s = 0; s < nseg; ++s) {
    const tn_param_seg &sg = segs[s];
    if (sg.maxnorm == 0.f || sg.rate == 0.f || sg.ndim == 1) continue;
    if (sg.ndim == 2) {
      maxnorm_cols_kernel<<<ceil_div(sg.cols, kMnCols), kMnCols * kMnSlices, 0, st>>>(
          theta + sg.offset, sg.rows, sg.cols, sg.maxnorm);
    } else {
      maxnorm_rows_kernel<<<ceil_div(sg.rows, 8), 256, 0, st>>>(theta + sg.offset, sg.rows,
                                                                sg.cols, sg.maxnorm);
    }
    TN_LAUNCH_CHECK("tn_sgd_momentum_maxnorm_update(maxnorm)");
  }
  return TN_OK;
}

extern "C" int tn_sgd_momentum_maxnorm_update(float *theta, float *vel, const float *grad,
                                              const tn_param_seg *segs, int nseg, int64_t total,
                                              const int32_t *ctl, float grad_scale,
                                              const float *nll_sum, float nll_scale,
                                              float *cost_out, void *workspace, void *stream) {
  PeerArgs peers{};
  peers.world = 1;
  return update_impl(theta, vel, grad, segs, nseg, total, ctl, grad_scale, nll_sum, nll_scale,
                     cost_out, workspace, peers, stream);
}

extern "C" int tn_allreduce_sgd_update(float *theta, float *vel, const float *const *peer_grads,
                                       int *const *peer_flags, int world, int rank,
                                       const tn_param_seg *segs, int nseg, int64_t total,
                                       int64_t peer_end, const int32_t *ctl, float grad_scale,
                                       float nll_scale, float *cost_out, void *workspace,
                                       void *stream) {
  TN_REQUIRE(peer_grads && peer_flags && world >= 1 && world <= kMaxPeers && rank >= 0 &&
                 rank < world,
             TN_ERR_ARG, "tn_allreduce_sgd_update: bad peer arguments (world %d, rank %d)", world, rank);
  PeerArgs peers{};
  peers.world = world;
  peers.rank = rank;
  peers.peer_end = (peer_end <= 0 || peer_end > total) ? total : peer_end;
  if (peers.peer_end < total) {
    bool on_boundary = false;
    for (int s = 0; s < nseg; ++s) on_boundary |= segs[s].offset == peers.peer_end;
    TN_REQUIRE(on_boundary, TN_ERR_ARG,
               "tn_allreduce_sgd_update: peer_end %lld is not the start of a segment", (long long)peer_end);
  }
  for (int r = 0; r < world; ++r) {
    TN_REQUIRE(peer_grads[r] && peer_flags[r], TN_ERR_ARG, "tn_allreduce_sgd_update: null peer %d", r);
    peers.grad[r] = peer_grads[r];
    peers.flags[r] = peer_flags[r];
  }
  return update_impl(theta, vel, peer_grads[rank], segs, nseg, total, ctl, grad_scale, nullptr,
                     nll_scale, cost_out, workspace, peers, stream);
}
