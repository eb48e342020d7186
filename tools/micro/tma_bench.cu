// Micro-benchmark: TMA tile-load throughput per SM for the box shapes gemm_tc.cu uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I theanet_b200/csrc -o tma_bench tools/micro/tma_bench.cu -lcuda
// Each CTA (one thread active) streams `nkb` k-blocks through a ring of S stages; per k-block it
// issues `nbox` boxes of (32 floats x rows).  Prints GB/s per SM and aggregate.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace tn::tc;

__global__ void k_tma(const __grid_constant__ CUtensorMap tm, int nkb, int nbox, int rows, int S,
                      int mn_major, int row_tiles, int kwrap) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = nbox * rows * 128;
  const uint32_t bar0 = base + S * stage_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tile = blockIdx.x % row_tiles;
  for (int kb = 0; kb < nkb + S; ++kb) {
    const int s = kb % S;
    if (kb >= S) mbar_wait(bar0 + 8 * s, ((kb / S) - 1) & 1);
    if (kb < nkb) {
#if !defined(TN_TMA_FIRST)
      mbar_expect_tx(bar0 + 8 * s, stage_bytes);
#endif
      for (int j = 0; j < nbox; ++j) {
        if (!mn_major) tma_load_2d(base + s * stage_bytes + j * rows * 128, &tm, bar0 + 8 * s, (kb % kwrap) * 32, tile * rows * nbox + j * rows);
        else tma_load_2d(base + s * stage_bytes + j * rows * 128, &tm, bar0 + 8 * s, tile * 32 * nbox + 32 * j, (kb % kwrap) * rows);
      }
#if defined(TN_TMA_FIRST)
      mbar_expect_tx(bar0 + 8 * s, stage_bytes);
#endif
    }
  }
}

__global__ void k_tma_g(const CUtensorMap *tmg, int nkb, int nbox, int rows, int S, int mn_major,
                        int row_tiles, int kwrap) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = nbox * rows * 128;
  const uint32_t bar0 = base + S * stage_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
    prefetch_tmap(tmg);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tile = blockIdx.x % row_tiles;
  for (int kb = 0; kb < nkb + S; ++kb) {
    const int s = kb % S;
    if (kb >= S && !(nkb & 1)) mbar_wait(bar0 + 8 * s, ((kb / S) - 1) & 1);
    if (kb < nkb) {
      mbar_expect_tx(bar0 + 8 * s, stage_bytes);
      for (int j = 0; j < nbox; ++j) {
        if (!mn_major) tma_load_2d(base + s * stage_bytes + j * rows * 128, tmg, bar0 + 8 * s, (kb % kwrap) * 32, tile * rows * nbox + j * rows);
        else tma_load_2d(base + s * stage_bytes + j * rows * 128, tmg, bar0 + 8 * s, tile * 32 * nbox + 32 * j, (kb % kwrap) * rows);
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)p;
  const int R = 1024;
  float *buf; cudaMalloc(&buf, (size_t)R * 4096 * 4 + 4096);
  cudaMemset(buf, 0, (size_t)R * 4096 * 4);
  struct Case { const char *name; int pitch; int mn; int nbox; int rows; int atom32; int promo; };
  Case cases[] = {
    {"K-major 1x(32x128) pitch 720 SW128", 720, 0, 1, 128, 0, 2},
    {"K-major 1x(32x128) pitch 768 SW128", 768, 0, 1, 128, 0, 2},
    {"K-major 1x(32x128) pitch 1024 SW128", 1024, 0, 1, 128, 0, 2},
    {"K-major 1x(32x128) pitch 720 SW128 nopromo", 720, 0, 1, 128, 0, 0},
    {"K-major 1x(32x32) pitch 720 SW128", 720, 0, 1, 32, 0, 2},
    {"MN-major 1x(32x32) pitch 500 ATOM32", 500, 1, 1, 32, 1, 2},
    {"MN-major 4x(32x32) pitch 500 ATOM32", 500, 1, 4, 32, 1, 2},
    {"MN-major 4x(32x32) pitch 512 ATOM32", 512, 1, 4, 32, 1, 2},
    {"MN-major 4x(32x32) pitch 720 ATOM32", 720, 1, 4, 32, 1, 2},
    {"MN-major 4x(32x32) pitch 500 SW128", 500, 1, 4, 32, 0, 2},
  };
  for (auto &c : cases) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)c.pitch, (cuuint64_t)R};
    cuuint64_t str[1] = {(cuuint64_t)c.pitch * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)c.rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     (CUtensorMapL2promotion)c.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    for (int grid : {1, 32, 128}) {
      for (int S : {2, 6}) {
        const int kwrap = c.mn ? R / c.rows : c.pitch / 32;
        const int nkb = 2000;
        const int row_tiles = c.mn ? c.pitch / (32 * c.nbox) : R / (c.rows * c.nbox);
        const size_t smem = (size_t)S * c.nbox * c.rows * 128 + 2048;
        cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_tma<<<grid, 32, smem>>>(tm, nkb, c.nbox, c.rows, S, c.mn, row_tiles, kwrap);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int i = 0; i < reps; ++i) k_tma<<<grid, 32, smem>>>(tm, nkb, c.nbox, c.rows, S, c.mn, row_tiles, kwrap);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (0) {
          CUtensorMap *tmg; cudaMalloc(&tmg, 256); cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice);
          cudaFuncSetAttribute(k_tma_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          k_tma_g<<<grid, 32, smem>>>(tmg, nkb + 1, c.nbox, c.rows, S, c.mn, row_tiles, kwrap);
          cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
          cudaEventRecord(f0);
          for (int i = 0; i < reps; ++i) k_tma_g<<<grid, 32, smem>>>(tmg, nkb + 1, c.nbox, c.rows, S, c.mn, row_tiles, kwrap);
          cudaEventRecord(f1); cudaEventSynchronize(f1);
          float ms2; cudaEventElapsedTime(&ms2, f0, f1);
          printf("   [no waits, issue only: %6.1f ns/kblock] ", ms2 * 1e6 / reps / nkb);
          cudaFree(tmg);
        }
        cudaError_t err = cudaGetLastError();
        const double us = ms * 1e3 / reps;
        const double bytes = (double)nkb * c.nbox * c.rows * 128;
        printf("%-46s grid %3d S %d: %7.2f us/launch  %6.1f ns/kblock  %6.1f GB/s per SM  %7.1f GB/s total %s\n", c.name, grid, S, us,
               us * 1e3 / nkb, bytes / us * 1e-3, bytes * grid / us * 1e-3, err == cudaSuccess ? "" : cudaGetErrorString(err));
      }
    }
  }
  return 0;
}
