// Peer-memory plumbing for the fused gradient all-reduce + update (update.cu): buffers that other
// ranks (one process per GPU) map over NVLink with CUDA IPC.  Not in the reference (it has no
// multi-device path); SURVEY.md 8(f1).
#include "common.cuh"

using namespace tn;

// cudaMalloc'ed (never from a caching allocator: the IPC handle must describe exactly this block),
// zero-filled
extern "C" int tn_peer_alloc(size_t bytes, void **ptr) {
  TN_REQUIRE(ptr && bytes > 0, TN_ERR_ARG, "tn_peer_alloc: bad argument");
  cudaError_t e = cudaMalloc(ptr, bytes);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_alloc: %s", cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, bytes);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_alloc: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_peer_free(void *ptr) {
  cudaError_t e = cudaFree(ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_peer_free: %s", cudaGetErrorString(e));
  return TN_OK;
}

// handle_out: 64 bytes (cudaIpcMemHandle_t)
extern "C" int tn_ipc_get_handle(void *ptr, void *handle_out) {
  TN_REQUIRE(ptr && handle_out, TN_ERR_ARG, "tn_ipc_get_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_get_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_ipc_open_handle(const void *handle, void **ptr) {
  TN_REQUIRE(handle && ptr, TN_ERR_ARG, "tn_ipc_open_handle: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_open_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}

extern "C" int tn_ipc_close_handle(void *ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_ipc_close_handle: %s", cudaGetErrorString(e));
  return TN_OK;
}

// ------------------------------------------------------------------------------------------------
// Two-shot all-reduce (sum) of buf[offset, offset + count) over peer memory, in place, replacing the
// NCCL ring for the step's big gradient bucket (1.4 MB: a ring over 8 ranks is 14 latency-bound hops,
// 56 us measured; here every rank moves 2 x count x 4 bytes over NVSwitch in two phases).
//   phase 1 (reduce-scatter): rank r owns slice r.  When every rank has announced its gradients
//     (token t1), r adds slice r of ALL ranks' buffers in rank order -- the same order everywhere --
//     and writes the total into slice r of its own buffer.
//   phase 2 (all-gather): when rank q has announced its total (token t2 = t1 + 1, sent by the last
//     CTA of q to finish phase 1), everybody copies slice q from q's buffer into its own.
// Rank q sends t2 only after it has finished READING the other ranks' slice q, so overwriting my
// copy of slice q after seeing q's t2 is safe; my own slice is read by nobody else in phase 1.
// Every rank ends with bit-identical totals.  Flags (int[32] per rank, zero-filled once, word 16 + s
// = token from rank s, word 24 = this rank's execution counter, word 25 = CTA ticket) are separate
// from the update kernel's words 0..8.  The caller alternates between two buffers from step to
// step, like tn_allreduce_sgd_update: a buffer is rewritten two executions later, when every
// reader has long passed the next execution's first handshake.
// ------------------------------------------------------------------------------------------------
namespace tn {

constexpr int kArMaxPeers = 8;
constexpr int kArFlag0 = 16, kArEpoch = 24, kArTicket = 25;
constexpr int kArThreads = 512;

struct PeerArArgs {
  float *buf[kArMaxPeers];
  int *flags[kArMaxPeers];
  int world, rank;
  int64_t offset, count4;       // in floats / in float4
};

__device__ __forceinline__ void ar_st_release_sys(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ar_ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ar_wait_all(const PeerArArgs &a, int token) {
  if ((int)threadIdx.x < a.world) {
    const int *f = a.flags[a.rank] + kArFlag0 + threadIdx.x;
    const long long t0 = clock64();
    while (ar_ld_acquire_sys(f) - token < 0) {
      if (clock64() - t0 > 20000000000ll) __trap();   // ~10 s: a peer died; fail loudly
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kArThreads)
peer_allreduce_kernel(const __grid_constant__ PeerArArgs a) {
  __shared__ int s_last;
  const int W = a.world, me = a.rank;
  int *mine = a.flags[me];
  const int t1 = 2 * __ldcg(mine + kArEpoch) + 1, t2 = t1 + 1;
  if (blockIdx.x == 0 && (int)threadIdx.x < W) {
    __threadfence_system();
    ar_st_release_sys(a.flags[threadIdx.x] + kArFlag0 + me, t1);      // "my gradients are complete"
  }
  ar_wait_all(a, t1);
  // ---- phase 1: my slice = sum over ranks, in rank order
  const int64_t per = (a.count4 + W - 1) / W;
  const int64_t lo = me * per, hi = lo + per < a.count4 ? lo + per : a.count4;
  const int64_t stride = (int64_t)gridDim.x * kArThreads;
  for (int64_t i = lo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += stride) {
    float4 v[kArMaxPeers];
#pragma unroll
    for (int r = 0; r < kArMaxPeers; ++r)
      if (r < W) v[r] = __ldcg(reinterpret_cast<const float4 *>(a.buf[r] + a.offset) + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < kArMaxPeers; ++r)
      if (r < W) {
        s.x = __fadd_rn(s.x, v[r].x); s.y = __fadd_rn(s.y, v[r].y);
        s.z = __fadd_rn(s.z, v[r].z); s.w = __fadd_rn(s.w, v[r].w);
      }
    reinterpret_cast<float4 *>(a.buf[me] + a.offset)[i] = s;
  }
  // the last CTA to finish announces the total
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(mine + kArTicket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (s_last && (int)threadIdx.x < W) {
    __threadfence_system();
    ar_st_release_sys(a.flags[threadIdx.x] + kArFlag0 + me, t2);
  }
  ar_wait_all(a, t2);
  // ---- phase 2: gather the other ranks' totals
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const int64_t qlo = q * per, qhi = qlo + per < a.count4 ? qlo + per : a.count4;
    const float4 *src = reinterpret_cast<const float4 *>(a.buf[q] + a.offset);
    float4 *dst = reinterpret_cast<float4 *>(a.buf[me] + a.offset);
    for (int64_t i = qlo + (int64_t)blockIdx.x * kArThreads + threadIdx.x; i < qhi; i += stride)
      dst[i] = __ldcg(src + i);
  }
  // the last CTA to leave advances the execution counter and clears the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(mine + kArTicket, 1) == 2 * (int)gridDim.x - 1) {
      mine[kArTicket] = 0;
      mine[kArEpoch] = (t1 - 1) / 2 + 1;
      __threadfence();
    }
  }
}

}  // namespace tn

extern "C" int tn_peer_allreduce(float *const *peer_bufs, int *const *peer_flags, int world, int rank,
                                 int64_t offset, int64_t count, void *stream) {
  const char *who = "tn_peer_allreduce";
  TN_REQUIRE(peer_bufs && peer_flags && world >= 1 && world <= kArMaxPeers && rank >= 0 && rank < world,
             TN_ERR_ARG, "%s: bad peer arguments (world %d, rank %d)", who, world, rank);
  TN_REQUIRE(offset >= 0 && count > 0 && offset % 4 == 0 && count % 4 == 0, TN_ERR_ALIGN,
             "%s: offset and count must be multiples of 4 floats", who);
  PeerArArgs a{};
  a.world = world;
  a.rank = rank;
  a.offset = offset;
  a.count4 = count / 4;
  for (int r = 0; r < world; ++r) {
    TN_REQUIRE(peer_bufs[r] && peer_flags[r], TN_ERR_ARG, "%s: null peer %d", who, r);
    a.buf[r] = peer_bufs[r];
    a.flags[r] = peer_flags[r];
  }
  // enough CTAs to keep a few hundred KB in flight over NVLink, few enough to leave the SMs to the
  // backward kernels this overlaps with
  const int64_t per = (a.count4 + world - 1) / world;
  int grid = (int)((per + kArThreads - 1) / kArThreads);
  grid = grid < 1 ? 1 : (grid > 24 ? 24 : grid);
  peer_allreduce_kernel<<<grid, kArThreads, 0, (cudaStream_t)stream>>>(a);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}
