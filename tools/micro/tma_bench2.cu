// Micro-benchmark 2: does TMA throughput per SM scale with the number of issuing warps, and with
// the bytes per mbarrier phase?  W warps, each with its own ring of S stages; per stage `nbox`
// boxes of (32 floats x 128 rows) = 16 KB each.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tn::tc;

__global__ void k(const __grid_constant__ CUtensorMap tm, int nkb, int nbox, int S, int kwrap) {
  extern __shared__ uint8_t raw[];
  const int warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const uint32_t base0 = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = nbox * 16384;
  const uint32_t base = base0 + warp * S * stage_bytes;
  const uint32_t bar0 = base0 + W * S * stage_bytes + warp * S * 8;
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) != 0) return;
  for (int kb = 0; kb < nkb + S; ++kb) {
    const int s = kb % S;
    if (kb >= S) mbar_wait(bar0 + 8 * s, ((kb / S) - 1) & 1);
    if (kb < nkb) {
      mbar_expect_tx(bar0 + 8 * s, stage_bytes);
      for (int j = 0; j < nbox; ++j)
        tma_load_2d(base + s * stage_bytes + j * 16384, &tm, bar0 + 8 * s, ((kb * nbox + j) % kwrap) * 32, (warp % 8) * 128);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)p;
  const int R = 1024, pitch = 1024;
  float *buf; cudaMalloc(&buf, (size_t)R * pitch * 4); cudaMemset(buf, 0, (size_t)R * pitch * 4);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)R};
  cuuint64_t str[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int nkb = 1000;
  struct C { int W, nbox, S; } cs[] = {{1,1,2},{1,1,4},{1,2,2},{1,4,2},{1,6,2},{2,1,2},{4,1,2},{4,1,3},{2,2,2},{2,4,1},{8,1,1}};
  for (auto c : cs) {
    const size_t smem = (size_t)c.W * c.S * c.nbox * 16384 + 2048;
    if (smem > 227 * 1024) { printf("skip W=%d nbox=%d S=%d\n", c.W, c.nbox, c.S); continue; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {1, 148}) {
      k<<<grid, 32 * c.W, smem>>>(tm, nkb, c.nbox, c.S, pitch / 32);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) k<<<grid, 32 * c.W, smem>>>(tm, nkb, c.nbox, c.S, pitch / 32);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      const double us = ms * 1e3 / 3;
      const double bytes = (double)nkb * c.nbox * 16384 * c.W;
      printf("warps %d  boxes/phase %d  stages %d  grid %3d: %7.1f ns/phase  %6.1f GB/s per SM  %7.1f GB/s total %s\n", c.W, c.nbox, c.S, grid,
             us * 1e3 / nkb, bytes / us * 1e-3, bytes * grid / us * 1e-3, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
  }
  return 0;
}
