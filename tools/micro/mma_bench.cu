// Micro-benchmark 3: tcgen05.mma issue rate per SM (no TMA): one thread issues `n` MMAs of
// 128 x N x (32 bytes of K) on fixed shared-memory tiles, then commits; duration by clock64.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tn::tc;

template <int KIND>
__global__ void k(int n, int N, int a_mn, int b_mn, int commit_every, long long *out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tslot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tslot;
#ifdef USE_ELECT
  if (warp == 0) {
    const uint32_t idesc = make_idesc(KIND, a_mn, b_mn, 128, N);
    const uint32_t lay_mn = KIND == KIND_TF32 ? LAYOUT_SW128_32B : LAYOUT_SW128;
    const uint32_t sbo_mn = KIND == KIND_TF32 ? 512u : 1024u;
    const uint64_t ad = a_mn ? make_smem_desc(base, 4096, sbo_mn, lay_mn) : make_smem_desc(base, 16, 1024);
    const uint64_t bd = b_mn ? make_smem_desc(base + 16384, 4096, sbo_mn, lay_mn) : make_smem_desc(base + 16384, 16, 1024);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      if (elect_one()) umma<KIND>(tm, ad, bd, idesc, i ? 1u : 0u);
      __syncwarp();
      if (commit_every && (i % commit_every) == commit_every - 1) {
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), phase);
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
#else
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = make_idesc(KIND, a_mn, b_mn, 128, N);
    const uint32_t lay_mn = KIND == KIND_TF32 ? LAYOUT_SW128_32B : LAYOUT_SW128;
    const uint32_t sbo_mn = KIND == KIND_TF32 ? 512u : 1024u;
    const uint64_t ad = a_mn ? make_smem_desc(base, 4096, sbo_mn, lay_mn) : make_smem_desc(base, 16, 1024);
    const uint64_t bd = b_mn ? make_smem_desc(base + 16384, 4096, sbo_mn, lay_mn) : make_smem_desc(base + 16384, 16, 1024);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      umma<KIND>(tm, ad, bd, idesc, i ? 1u : 0u);
      if (commit_every && (i % commit_every) == commit_every - 1) {
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), phase);
    out[blockIdx.x] = clock64() - t0;
  }
#endif
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  long long *out; cudaMalloc(&out, 148 * 8);
  const int n = 2000;
  const size_t smem = 64 * 1024;
  cudaFuncSetAttribute(k<KIND_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<KIND_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {32, 64, 128, 256})
      for (int maj = 0; maj < 4; ++maj)
        for (int ce : {0, 8, 1}) {
          if (maj && ce) continue;
          const int a_mn = maj & 1, b_mn = maj >> 1;
          for (int grid : {1, 148}) {
            if (kind == 0) k<KIND_TF32><<<grid, 64, smem>>>(n, N, a_mn, b_mn, ce, out);
            else k<KIND_BF16><<<grid, 64, smem>>>(n, N, a_mn, b_mn, ce, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            const double cyc = (double)mx / n;
            const double macs = 128.0 * N * (kind == 0 ? 8 : 16);
            printf("%s N=%3d A:%s B:%s commit_every=%d grid %3d: %7.1f cycles/MMA  %7.0f MAC/clk/SM %s\n", kind ? "bf16" : "tf32", N,
                   a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", ce, grid, cyc, macs / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
        }
  return 0;
}
