// ConvLayer on the 5th-generation tensor cores: im2col-free implicit GEMM in bf16 with fp32
// accumulation, for wide-channel layers (C % 64 == 0; the CIFAR-shaped config C4 of BASELINE.json).
// (theanet/layer/convpool.py:42-72 nnet.conv2d with filter_flip=True, stride 1; gradients via
// tt.grad, layer.py:83.)
//
// Activations are NHWC bf16.  The GEMM view of a convolution is
//     out[p, n] = sum_{tap=(u,v)} sum_c  x[p + tap, c] * Wp[n, tap, c]          p = output pixel
// and a k-block of the main loop is ONE filter tap and 64 channels: its A tile (128 pixels x 64
// channels) is a single 4-D TMA box of the activation tensor shifted by the tap -- out-of-bounds
// rows/columns are zero-filled by the TMA unit, which IS the zero padding of the convolution.  The
// B tile is a 2-D box of the packed filter matrix.  Both land in 128-byte-swizzled K-major tiles
// that tcgen05.mma (kind::f16, bf16 in, fp32 accumulate in TMEM) consumes directly.
//
//   tn_conv2d_tc_fprop   a = act(conv(x) + b) (bf16 NHWC) and, fused, pooled = maxpool2x2(a)
//   tn_conv2d_tc_dgrad   dx = conv(gz, unflipped W^T): the same kernel on the gradient tensor
//   tn_conv2d_tc_wgrad   dW[tap] = gz^T . x_shifted(tap): both operands MN-major (pixels are K),
//                        split-K over pixel tiles, fixed-order reduction + un-flip in a finish kernel
//
// Warp roles as in gemm_tc.cu: 2 TMA producer warps (stages owned round-robin), one MMA-issuing
// lane, 4 epilogue warps reading the accumulator with tcgen05.ld.  Lessons from the round-1
// micro-benchmarks are built in: N = 256 accumulator tiles where the layer has >= 256 maps
// (dependent tcgen05.mma issue at ~112-128 cycles each, so only N = 256 reaches the tensor peak),
// otherwise MMAs are dealt to independent accumulators; a compact epilogue loop instead of unrolled code.
#include <stdlib.h>

#include <mutex>

#include <cuda_bf16.h>

#include "common.cuh"
#include "dense_tc.cuh"
#include "tc_ptx.cuh"

namespace tn {
using namespace tc;

constexpr int CT_PROD = 2;
constexpr int CT_MMA_WARP = CT_PROD + 4;  // warps [CT_PROD, CT_PROD+4): epilogue; then MMA
constexpr int CT_THREADS = 32 * (CT_MMA_WARP + 1);
constexpr int CT_A_BYTES = 128 * 128;      // 128 pixels x 64 bf16

struct ConvTcArgs {
  int B, Ho, Wo;       // output geometry (input geometry is in the tensor map)
  int C;               // input channels of the GEMM (K = taps * C)
  int N;               // output channels of the GEMM
  int f, pad;          // filter size; input coordinate = output coordinate + tap - pad
  int TH, TB;          // tile = TB images x TH rows x Wo columns = 128 pixels
  const float *bias;   // [N] or null
  ActK ak;
  __nv_bfloat16 *out;     // (B, Ho, Wo, N)
  __nv_bfloat16 *pooled;  // (B, Ho/2, Wo/2, N) or null
};

template <int BN>
struct ConvTcCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = CT_A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int NACC = BN >= 256 ? 1 : (BN == 128 ? 2 : 4);
  static constexpr int TMEM_COLS = NACC * BN;  // 256
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}

template <int BN>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const ConvTcArgs g) {
  // Persistent: every CTA walks the (pixel tile, channel tile) work list with stride gridDim.x.
  // The shared-memory ring and its phases run on across tiles; the accumulator is double-buffered
  // in TMEM (2 x 256 columns) so that the epilogue of tile i overlaps the main loop of tile i+1.
  using Cfg = ConvTcCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr int NACC = Cfg::NACC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + S * Cfg::STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull = [&](int b) { return bar0 + 8u * (2 * S + b); };
  auto tempty = [&](int b) { return bar0 + 8u * (2 * S + 2 + b); };
  const uint32_t tslot = bar0 + 8u * (2 * S + 4);
  auto stA = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto stB = [&](int s) { return base + s * Cfg::STAGE_BYTES + CT_A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_y = g.Ho / g.TH;
  const int mtiles = ((g.B + g.TB - 1) / g.TB) * tiles_y;
  const int ntiles = (g.N + BN - 1) / BN;
  const int nwork = mtiles * ntiles;
  const int cchunks = g.C >> 6;
  const int nkb = g.f * g.f * cchunks;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == CT_MMA_WARP) tmem_alloc(tslot, 2 * Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (warp < CT_PROD) {
    if (lane == 0) {
      int it = 0;   // running k-block counter across tiles
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int mt = w / ntiles, n0 = (w - mt * ntiles) * BN;   // channel tiles of a pixel tile are neighbours
        const int b0 = (mt / tiles_y) * g.TB, y0 = (mt % tiles_y) * g.TH;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          if (s % CT_PROD != warp) continue;   // one owner per stage: phases stay in order
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(empty(s), ph ^ 1u);
          mbar_expect_tx(full(s), CT_A_BYTES + Cfg::B_BYTES);
          const int tap = kb / cchunks, cc = kb - tap * cchunks;
          const int r = tap / g.f, sx = tap - r * g.f;
          tma_load_4d(stA(s), &tmX, full(s), cc * 64, sx - g.pad, y0 + r - g.pad, b0);
          tma_load_2d(stB(s), &tmW, full(s), tap * g.C + cc * 64, n0);
        }
      }
    }
  } else if (warp == CT_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(KIND_BF16, 0, 0, 128, BN);
      int it = 0, ti = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++ti) {
        const int ab = ti & 1;
        mbar_wait(tempty(ab), ((uint32_t)(ti >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
        tcgen05_fence_after();
        const uint32_t td = tmem_base + (uint32_t)(ab * Cfg::TMEM_COLS);
        int idx = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(full(s), ph);
          tcgen05_fence_after();
#pragma unroll
          for (int j = 0; j < 4; ++j) {   // 64 channels = 4 x K16
            const uint64_t ad = make_smem_desc(stA(s) + j * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(stB(s) + j * 32, 16, 1024);
            umma<KIND_BF16>(td + (uint32_t)((idx % NACC) * BN), ad, bd, idesc, idx >= NACC ? 1u : 0u);
            ++idx;
          }
          umma_commit(empty(s));
        }
        umma_commit(tfull(ab));
      }
    }
  } else {
    // ===== epilogue: one output pixel per thread =====
    const int q = warp & 3;
    const int row = q * 32 + lane;               // pixel inside the tile
    const int per_img = g.TH * g.Wo;
    const int bl = row / per_img, rem = row - bl * per_img;
    const int yl = rem / g.Wo, xo = rem - yl * g.Wo;
    const bool pool = g.pooled != nullptr;
    const float sl = g.ak.act == TN_ACT_LINEAR ? 1.f : g.ak.s_neg;
    int ti = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++ti) {
      const int mt = w / ntiles, n0 = (w - mt * ntiles) * BN;
      const int b = (mt / tiles_y) * g.TB + bl, y = (mt % tiles_y) * g.TH + yl;
      const bool ok = b < g.B;
      const int ab = ti & 1;
      mbar_wait(tfull(ab), (uint32_t)(ti >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * Cfg::TMEM_COLS);
      __nv_bfloat16 *orow = g.out + (((size_t)b * g.Ho + y) * g.Wo + xo) * g.N + n0;
      const bool writer = pool && ((xo | y) & 1) == 0;
      __nv_bfloat16 *prow =
          pool ? g.pooled + (((size_t)b * (g.Ho >> 1) + (y >> 1)) * (g.Wo >> 1) + (xo >> 1)) * g.N + n0
               : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= g.N) break;
        float v[16];
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          uint32_t u[16];
          tmem_ld16(tlane + (uint32_t)(a * BN + c0), u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = a ? v[i] + __uint_as_float(u[i]) : __uint_as_float(u[i]);
        }
        if (g.bias) {   // bias + ReLU-family activation; the negative slope is a plain multiply
                        // (the exact (z*NN)/100 of the fp32 path is below bf16 resolution)
          const float4 *b4 = reinterpret_cast<const float4 *>(g.bias + n0 + c0);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 bb = __ldg(b4 + i4);
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = v[4 * i4 + e] + bv[e];
              v[4 * i4 + e] = z > 0.f ? z : z * sl;
            }
          }
        }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (ok) {
          uint4 *o = reinterpret_cast<uint4 *>(orow + c0);
          o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (pool) {   // 2x2 max over the bf16-rounded values: partners are lanes ^1 and ^Wo
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 m = *reinterpret_cast<__nv_bfloat162 *>(&pk[i]);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, pk[i], 1);
            m = __hmax2(m, *reinterpret_cast<__nv_bfloat162 *>(&o1));
            uint32_t mm = *reinterpret_cast<uint32_t *>(&m);
            uint32_t o2 = __shfl_xor_sync(0xffffffffu, mm, g.Wo);
            m = __hmax2(m, *reinterpret_cast<__nv_bfloat162 *>(&o2));
            pk[i] = *reinterpret_cast<uint32_t *>(&m);
          }
          if (writer && ok) {
            uint4 *o = reinterpret_cast<uint4 *>(prow + c0);
            o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tempty(ab));   // 128 arrivals hand the accumulator buffer back to the MMA lane
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == CT_MMA_WARP) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 2 * Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: D[m, c] (per tap) = sum over pixels gz[p, m] * x[p + tap, c]; A and B are MN-major tiles
// (pixels = K run along the shared-memory rows)
// ---------------------------------------------------------------------------------------------
struct WgradTcArgs {
  int B, Ho, Wo, M, C, f, pad, TH, TB;
  int ptiles;       // pixel tiles in total
  int nsplit;       // split-K factor (grid.z)
  float *partial;   // [nsplit][taps][M][C]
};

constexpr int WG_BN = 64;      // input channels per CTA (one 128-byte atom)
constexpr int WG_STAGE = 2 * CT_A_BYTES + CT_A_BYTES;  // gz: 2 x (128 px x 64 maps) + x: 128 px x 64 ch
constexpr int WG_STAGES = 4;
constexpr int WG_NACC = 4;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 256;

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                     const WgradTcArgs g) {
  constexpr int S = WG_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + S * WG_STAGE;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t accum = bar0 + 8u * (2 * S);
  const uint32_t tslot = accum + 8u;
  auto stA = [&](int s) { return base + s * WG_STAGE; };
  auto stB = [&](int s) { return base + s * WG_STAGE + 2 * CT_A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ctiles = g.C / WG_BN;
  const int tap = blockIdx.x / ctiles, c0 = (blockIdx.x - tap * ctiles) * WG_BN;
  const int m0 = blockIdx.y * 128;
  const int r = tap / g.f, sx = tap - r * g.f;
  const int tiles_y = g.Ho / g.TH;
  // this CTA's share of the pixel tiles
  const int per = (g.ptiles + g.nsplit - 1) / g.nsplit;
  const int t_lo = blockIdx.z * per, t_hi = min(g.ptiles, t_lo + per);
  const int nkb = max(0, t_hi - t_lo);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmG);
    prefetch_tmap(&tmX);
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == CT_MMA_WARP) tmem_alloc(tslot, WG_NACC * WG_BN);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (warp < CT_PROD) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        if (s % CT_PROD != warp) continue;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), WG_STAGE);
        const int t = t_lo + kb;
        const int b0 = (t / tiles_y) * g.TB, y0 = (t % tiles_y) * g.TH;
        tma_load_4d(stA(s), &tmG, full(s), m0, 0, y0, b0);
        tma_load_4d(stA(s) + CT_A_BYTES, &tmG, full(s), m0 + 64, 0, y0, b0);
        tma_load_4d(stB(s), &tmX, full(s), c0, sx - g.pad, y0 + r - g.pad, b0);
      }
    }
  } else if (warp == CT_MMA_WARP) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc(KIND_BF16, 1, 1, 128, WG_BN);
      int idx = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(full(s), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // 128 pixels = 8 x K16; a K16 step is 16 rows = 2 KB
          const uint64_t ad = make_smem_desc(stA(s) + j * 2048, CT_A_BYTES, 1024);
          const uint64_t bd = make_smem_desc(stB(s) + j * 2048, CT_A_BYTES, 1024);
          umma<KIND_BF16>(tmem_base + (uint32_t)((idx % WG_NACC) * WG_BN), ad, bd, idesc,
                          idx >= WG_NACC ? 1u : 0u);
          ++idx;
        }
        umma_commit(empty(s));
      }
      umma_commit(accum);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    float *prow = g.partial + (((size_t)blockIdx.z * g.f * g.f + tap) * g.M + m) * g.C + c0;
    if (nkb > 0) {
      mbar_wait(accum, 0);
      tcgen05_fence_after();
    }
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < WG_BN; cc += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      if (nkb > 0) {
#pragma unroll
        for (int a = 0; a < WG_NACC; ++a) {
          uint32_t u[16];
          tmem_ld16(tlane + (uint32_t)(a * WG_BN + cc), u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(u[i]);
        }
      }
      if (m < g.M) {
        float4 *o = reinterpret_cast<float4 *>(prow + cc);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == CT_MMA_WARP) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, WG_NACC * WG_BN);
  }
}

// dW[m][c][f-1-u][f-1-v] = sum_split partial[split][(u,v)][m][c]  (fixed order)
__global__ void conv_tc_wgrad_finish_kernel(const float *__restrict__ partial, int nsplit, int M,
                                            int C, int f, float *__restrict__ dW) {
  const int total = f * f * M * C;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(size_t)k * total + t];
    const int c = t % C;
    int rr = t / C;
    const int m = rr % M;
    const int tap = rr / M;
    const int u = tap / f, v = tap - u * f;
    dW[((m * C + c) * f + (f - 1 - u)) * f + (f - 1 - v)] = s;
  }
}

// db[m] = sum_p gz[p][m] (bf16 NHWC, fp32 accumulate), two fixed-order stages:
// stage 1: CTA (channel block of 64, pixel slice) -> partial[slice][m]; stage 2 adds the slices.
constexpr int kColSlices = 296;     // two CTAs per SM
// a thread owns 8 consecutive channels (one 16-byte load per pixel), M/8 threads span a pixel row
// and the CTA walks 1024/(M/8) pixels per pass with four loads in flight per thread: HBM-bound
// (the first version read 2 bytes per thread per dependent iteration and ran at a third of that)
__global__ void __launch_bounds__(1024)
colsum_bf16_kernel(const __nv_bfloat16 *__restrict__ gz, int64_t P, int M,
                   float *__restrict__ partial) {
  __shared__ float red[1024 * 8];
  const int vpr = M >> 3;                         // threads per pixel
  const int rpp = 1024 / vpr;                     // pixels per pass
  const int tx = threadIdx.x % vpr, ty = threadIdx.x / vpr;
  const int64_t per = (P + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per, p1 = p0 + per < P ? p0 + per : P;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ty < rpp) {
    const uint4 *src = reinterpret_cast<const uint4 *>(gz) + tx;
    for (int64_t p = p0 + ty; p < p1; p += 4 * rpp) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t q = p + (int64_t)u * rpp;
        v[u] = q < p1 ? __ldg(src + q * vpr) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {             // bf16 -> fp32 is a 16-bit shift
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty * M + 8 * tx + j] = acc[j];
  }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += 1024) {
    float t = 0.f;
    for (int k = 0; k < rpp; ++k) t += red[k * M + m];
    partial[(size_t)blockIdx.x * M + m] = t;
  }
}
// one warp per channel: the lanes walk the slices, then a fixed-order shuffle tree (a thread per
// channel adding 296 slices one after the other took 20 us)
__global__ void colsum_finish_kernel(const float *__restrict__ partial, int nslices, int M,
                                     float *__restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float s = 0.f;
  for (int k = lane; k < nslices; k += 32) s += partial[(size_t)k * M + m];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) db[m] = s;
}

// ---------------------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------------------
// fprop: Wp[m][(u,v)][c] = W[m][c][f-1-u][f-1-v];  dgrad: Wp[c][(u,v)][m] = W[m][c][u][v]
__global__ void pack_weights_kernel(const float *__restrict__ W, __nv_bfloat16 *__restrict__ Wp,
                                    int M, int C, int f, int dgrad) {
  const int total = M * C * f * f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    if (!dgrad) {
      const int c = t % C;
      int rr = t / C;
      const int tap = rr % (f * f);
      const int m = rr / (f * f);
      const int u = tap / f, v = tap - u * f;
      Wp[t] = __float2bfloat16_rn(W[((m * C + c) * f + (f - 1 - u)) * f + (f - 1 - v)]);
    } else {
      const int m = t % M;
      int rr = t / M;
      const int tap = rr % (f * f);
      const int c = rr / (f * f);
      const int u = tap / f, v = tap - u * f;
      Wp[t] = __float2bfloat16_rn(W[((m * C + c) * f + u) * f + v]);
    }
  }
}

__global__ void nchw_to_nhwc_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y,
                                         int64_t total, int C, int HW) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const int64_t rr = t / C;
    const int p = (int)(rr % HW);
    const int64_t b = rr / HW;
    y[t] = __float2bfloat16_rn(x[(b * C + c) * HW + p]);
  }
}

__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y,
                                         int64_t total, int C, int HW) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % HW);
    const int64_t rr = t / HW;
    const int c = (int)(rr % C);
    const int64_t b = rr / C;
    y[t] = __bfloat162float(x[(b * HW + p) * C + c]);
  }
}

// ---- first layer (few input channels): im2col to 64-wide bf16 rows, then a 1x1 tensor-core conv --
// xcol[b,y,x,k] = xpad[b, c, y+u-pad, x+v-pad] for k = (c*f+u)*f+v < C*f*f, 0 for the padding up to 64
template <int F>
__global__ void im2col_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ xcol,
                                   uint32_t npix, int C, int S, int f_rt, int pad, FastDiv32 divS) {
  // one thread = one pixel: C*f*f coalesced loads (threads run along x), one 128-byte row out
  (void)f_rt;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < npix; t += gridDim.x * blockDim.x) {
    const uint32_t r2 = divS.div(t);
    const int xx = (int)(t - r2 * S);
    const uint32_t b = divS.div(r2);
    const int yy = (int)(r2 - b * S);
    float v[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) v[k] = 0.f;
    const float *img = x + (size_t)b * C * S * S;
#pragma unroll
    for (int c = 0; c < 7; ++c) {      // compile-time k keeps v[] in registers
      if (c < C) {
#pragma unroll
        for (int u = 0; u < F; ++u) {
          const int y = yy + u - pad;
#pragma unroll
          for (int vv = 0; vv < F; ++vv) {
            const int x2 = xx + vv - pad;
            if (y >= 0 && y < S && x2 >= 0 && x2 < S)
              v[(c * F + u) * F + vv] = img[((size_t)c * S + y) * S + x2];
          }
        }
      }
    }
    uint4 *o = reinterpret_cast<uint4 *>(xcol) + (size_t)t * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      o[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                        pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
  }
}
// Wcol[m][k] = W[m][c][f-1-u][f-1-v] (k as above), zero padded to 64
__global__ void pack_weights_im2col_kernel(const float *__restrict__ W, __nv_bfloat16 *__restrict__ Wc,
                                           int M, int C, int f) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * 64) return;
  const int m = t >> 6, k = t & 63;
  float v = 0.f;
  if (k < C * f * f) {
    const int c = k / (f * f), uv = k - c * f * f;
    const int u = uv / f, vv = uv - u * f;
    v = W[((m * C + c) * f + (f - 1 - u)) * f + (f - 1 - vv)];
  }
  Wc[t] = __float2bfloat16_rn(v);
}
// dW[m][c][f-1-u][f-1-v] = dWcol[m][k]
__global__ void unpack_wgrad_im2col_kernel(const float *__restrict__ dWc, float *__restrict__ dW,
                                           int M, int C, int f) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = C * f * f;
  if (t >= M * K) return;
  const int m = t / K, k = t - m * K;
  const int c = k / (f * f), uv = k - c * f * f;
  const int u = uv / f, vv = uv - u * f;
  dW[((m * C + c) * f + (f - 1 - u)) * f + (f - 1 - vv)] = dWc[m * 64 + k];
}
// 2x2 max-pool of an NHWC bf16 tensor (for outputs too wide for the fused epilogue pool)
__global__ void maxpool2_nhwc_kernel(const __nv_bfloat16 *__restrict__ a, __nv_bfloat16 *__restrict__ p,
                                     uint32_t total8, int S, int M, FastDiv32 divM8, FastDiv32 divP) {
  const int P = S >> 1, M8 = M >> 3;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total8; t += gridDim.x * blockDim.x) {
    uint32_t rr = divM8.div(t);
    const int m8 = (int)(t - rr * M8);
    uint32_t r2 = divP.div(rr);
    const int ox = (int)(rr - r2 * P);
    const uint32_t b = divP.div(r2);
    const int oy = (int)(r2 - b * P);
    const uint4 *src = reinterpret_cast<const uint4 *>(a) + (((size_t)b * S + 2 * oy) * S + 2 * ox) * M8 + m8;
    uint4 q[4] = {src[0], src[M8], src[(size_t)S * M8], src[(size_t)S * M8 + M8]};
    uint4 r = q[0];
    __nv_bfloat162 *rv = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const __nv_bfloat162 *qv = reinterpret_cast<const __nv_bfloat162 *>(&q[w]);
#pragma unroll
      for (int e = 0; e < 4; ++e) rv[e] = __hmax2(rv[e], qv[e]);
    }
    reinterpret_cast<uint4 *>(p)[t] = r;
  }
}

// gz[b,y,x,m] = [a == pooled(window)] * dtop(window) * act'(a)   (2x2 windows, NHWC bf16 a / pooled;
// every tied maximum receives the gradient, as Theano's MaxPoolGrad).  dtop is either float32 NCHW
// (B, M, P, P) -- the gradient arriving from a dense layer -- or bf16 NHWC (B, P, P, M) -- the dx of
// the tensor-core conv above.  Without a pool layer (pooled == null) dtop has the shape of a.
__global__ void poolbwd_nhwc_kernel(const __nv_bfloat16 *__restrict__ a,
                                    const __nv_bfloat16 *__restrict__ pooled,
                                    const void *__restrict__ dtop, int dtop_nchw_f32,
                                    __nv_bfloat16 *__restrict__ gz, uint32_t total8, int S, int M,
                                    FastDiv32 divM8, FastDiv32 divS, ActK ak) {
  // one thread = 8 consecutive channels of one pixel (16-byte loads / stores)
  const int P = S >> 1, M8 = M >> 3;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total8; t += gridDim.x * blockDim.x) {
    uint32_t rr = divM8.div(t);
    const int m0 = (int)(t - rr * M8) * 8;
    uint32_t r2 = divS.div(rr);
    const int x = (int)(rr - r2 * S);
    const uint32_t b = divS.div(r2);
    const int y = (int)(r2 - b * S);
    const uint4 av4 = reinterpret_cast<const uint4 *>(a)[t];
    const __nv_bfloat16 *av = reinterpret_cast<const __nv_bfloat16 *>(&av4);
    float g[8];
    bool hit[8];
    if (pooled) {
      const size_t o = (((size_t)b * P + (y >> 1)) * P + (x >> 1)) * M + m0;
      const uint4 pv4 = *reinterpret_cast<const uint4 *>(pooled + o);
      const __nv_bfloat16 *pv = reinterpret_cast<const __nv_bfloat16 *>(&pv4);
#pragma unroll
      for (int e = 0; e < 8; ++e) hit[e] = av[e] == pv[e];
      if (dtop_nchw_f32) {
        const float *d = reinterpret_cast<const float *>(dtop);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          g[e] = hit[e] ? d[(((size_t)b * M + m0 + e) * P + (y >> 1)) * P + (x >> 1)] : 0.f;
      } else {
        const uint4 dv4 = *reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(dtop) + o);
        const __nv_bfloat16 *dv = reinterpret_cast<const __nv_bfloat16 *>(&dv4);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = hit[e] ? __bfloat162float(dv[e]) : 0.f;
      }
    } else if (dtop_nchw_f32) {
      const float *d = reinterpret_cast<const float *>(dtop);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = d[(((size_t)b * M + m0 + e) * S + y) * S + x];
    } else {
      const uint4 dv4 = reinterpret_cast<const uint4 *>(dtop)[t];
      const __nv_bfloat16 *dv = reinterpret_cast<const __nv_bfloat16 *>(&dv4);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = __bfloat162float(dv[e]);
    }
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pk[e] = pack_bf16x2(g[2 * e] * act_bwd_t<false>(ak, __bfloat162float(av[2 * e])),
                          g[2 * e + 1] * act_bwd_t<false>(ak, __bfloat162float(av[2 * e + 1])));
    reinterpret_cast<uint4 *>(gz)[t] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn ct_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// NHWC bf16 activation tensor (B, H, W, C) with a (64 ch, bw, bh, bb) box
static int map_nhwc(CUtensorMap *map, const void *ptr, int B, int H, int W, int C, int bw, int bh,
                    int bb, const char *who) {
  EncodeTiledFn fn = ct_encode_fn();
  TN_REQUIRE(fn, TN_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available", who);
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TN_REQUIRE(r == CUDA_SUCCESS, TN_ERR_CUDA, "%s: cuTensorMapEncodeTiled(NHWC) failed (%d)", who, (int)r);
  return TN_OK;
}

// packed filter matrix (rows, K) bf16 with a (64, box_rows) box
static int map_w(CUtensorMap *map, const void *ptr, int rows, int K, int box_rows, const char *who) {
  EncodeTiledFn fn = ct_encode_fn();
  TN_REQUIRE(fn, TN_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available", who);
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TN_REQUIRE(r == CUDA_SUCCESS, TN_ERR_CUDA, "%s: cuTensorMapEncodeTiled(W) failed (%d)", who, (int)r);
  return TN_OK;
}

// tile geometry: whole rows; 128 pixels = TB images x TH rows x Wo columns
static int tile_geom(int Ho, int Wo, int *TH, int *TB, const char *who) {
  TN_REQUIRE(Wo >= 4 && Wo <= 128 && (Wo & (Wo - 1)) == 0, TN_ERR_UNSUPPORTED,
             "%s: output width %d must be a power of two in [4, 128]", who, Wo);
  int th = 128 / Wo;
  if (th > Ho) th = Ho;
  TN_REQUIRE(Ho % th == 0 && 128 % (Wo * th) == 0, TN_ERR_UNSUPPORTED,
             "%s: output %dx%d does not tile into 128-pixel row blocks", who, Ho, Wo);
  *TH = th;
  *TB = 128 / (Wo * th);
  TN_REQUIRE(*TB == 1 || th == Ho, TN_ERR_UNSUPPORTED, "%s: bad tile geometry", who);
  return TN_OK;
}

template <int BN>
static int launch_conv_tc(const CUtensorMap &tmX, const CUtensorMap &tmW, const ConvTcArgs &g,
                          const char *who, cudaStream_t st) {
  using Cfg = ConvTcCfg<BN>;
  auto k = conv_tc_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  const int nwork = ceil_div(g.B, g.TB) * (g.Ho / g.TH) * ceil_div(g.N, BN);
  k<<<nwork < kNumSM ? nwork : kNumSM, CT_THREADS, Cfg::SMEM, st>>>(tmX, tmW, g);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

static int conv_tc_run(const void *x, const void *Wp, ConvTcArgs g, int H, int W, const char *who,
                       cudaStream_t st) {
  TN_REQUIRE(g.C % 64 == 0 && g.N % 16 == 0, TN_ERR_UNSUPPORTED,
             "%s: needs input channels %% 64 == 0 and output channels %% 16 == 0 (got %d, %d)", who,
             g.C, g.N);
  int rc = tile_geom(g.Ho, g.Wo, &g.TH, &g.TB, who);
  if (rc) return rc;
  int BN = g.N >= 256 ? 256 : (g.N >= 128 ? 128 : 64);
  if (const char *e = getenv("TN_CONV_BN")) {
    const int v = atoi(e);
    if (v == 64 || v == 128 || v == 256) BN = v;
  }
  CUtensorMap tmX, tmW;
  rc = map_nhwc(&tmX, x, g.B, H, W, g.C, g.Wo, g.TH, g.TB, who);
  if (rc) return rc;
  rc = map_w(&tmW, Wp, g.N, g.f * g.f * g.C, BN, who);
  if (rc) return rc;
  switch (BN) {
    case 256: return launch_conv_tc<256>(tmX, tmW, g, who, st);
    case 128: return launch_conv_tc<128>(tmX, tmW, g, who, st);
    default: return launch_conv_tc<64>(tmX, tmW, g, who, st);
  }
}

}  // namespace tn

using namespace tn;

static int blocks_for(int64_t n) { return (int)min64(ceil_div64(n, 256), (int64_t)kNumSM * 16); }

extern "C" int tn_nchw_f32_to_nhwc_bf16(const float *x, void *y, int B, int C, int H, int W,
                                        void *stream) {
  TN_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, TN_ERR_ARG, "tn_nchw_f32_to_nhwc_bf16: bad argument");
  const int64_t total = (int64_t)B * C * H * W;
  nchw_to_nhwc_bf16_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16 *)y, total, C, H * W);
  TN_LAUNCH_CHECK("tn_nchw_f32_to_nhwc_bf16");
  return TN_OK;
}

extern "C" int tn_nhwc_bf16_to_nchw_f32(const void *x, float *y, int B, int C, int H, int W,
                                        void *stream) {
  TN_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, TN_ERR_ARG, "tn_nhwc_bf16_to_nchw_f32: bad argument");
  const int64_t total = (int64_t)B * C * H * W;
  nhwc_bf16_to_nchw_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)x, y, total, C, H * W);
  TN_LAUNCH_CHECK("tn_nhwc_bf16_to_nchw_f32");
  return TN_OK;
}

extern "C" int tn_im2col_bf16(const float *x, void *xcol, int B, int C, int S, int f, int pad_lo,
                              void *stream) {
  TN_REQUIRE(x && xcol && B > 0 && C > 0 && S > 0 && f > 0, TN_ERR_ARG, "tn_im2col_bf16: bad argument");
  TN_REQUIRE(C * f * f <= 64, TN_ERR_UNSUPPORTED, "tn_im2col_bf16: C*f*f = %d exceeds 64", C * f * f);
  const int64_t npix = (int64_t)B * S * S;
  TN_REQUIRE(npix * 64 < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_im2col_bf16: tensor too large");
  TN_REQUIRE(f == 3 && C <= 7, TN_ERR_UNSUPPORTED,
             "tn_im2col_bf16: built for 3x3 filters on <= 7 channels (got f=%d C=%d)", f, C);
  im2col_bf16_kernel<3><<<blocks_for(npix), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16 *)xcol, (uint32_t)npix, C, S, f, pad_lo, FastDiv32(S));
  TN_LAUNCH_CHECK("tn_im2col_bf16");
  return TN_OK;
}

extern "C" int tn_conv2d_tc_pack_weights_im2col(const float *W, void *Wcol, int M, int C, int f,
                                                void *stream) {
  TN_REQUIRE(W && Wcol && M > 0 && C * f * f <= 64, TN_ERR_ARG, "tn_conv2d_tc_pack_weights_im2col: bad argument");
  pack_weights_im2col_kernel<<<ceil_div(M * 64, 256), 256, 0, (cudaStream_t)stream>>>(
      W, (__nv_bfloat16 *)Wcol, M, C, f);
  TN_LAUNCH_CHECK("tn_conv2d_tc_pack_weights_im2col");
  return TN_OK;
}

extern "C" int tn_conv2d_tc_unpack_wgrad_im2col(const float *dWcol, float *dW, int M, int C, int f,
                                                void *stream) {
  TN_REQUIRE(dWcol && dW && M > 0 && C * f * f <= 64, TN_ERR_ARG, "tn_conv2d_tc_unpack_wgrad_im2col: bad argument");
  unpack_wgrad_im2col_kernel<<<ceil_div(M * C * f * f, 256), 256, 0, (cudaStream_t)stream>>>(dWcol, dW, M, C, f);
  TN_LAUNCH_CHECK("tn_conv2d_tc_unpack_wgrad_im2col");
  return TN_OK;
}

extern "C" int tn_maxpool2_nhwc_bf16(const void *a, void *pooled, int B, int S, int M, void *stream) {
  TN_REQUIRE(a && pooled && B > 0 && S > 0 && S % 2 == 0 && M > 0 && M % 8 == 0, TN_ERR_ARG,
             "tn_maxpool2_nhwc_bf16: bad argument");
  const int64_t total8 = (int64_t)B * (S / 2) * (S / 2) * (M / 8);
  TN_REQUIRE(total8 < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_maxpool2_nhwc_bf16: tensor too large");
  maxpool2_nhwc_kernel<<<blocks_for(total8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)a, (__nv_bfloat16 *)pooled, (uint32_t)total8, S, M, FastDiv32(M / 8),
      FastDiv32(S / 2));
  TN_LAUNCH_CHECK("tn_maxpool2_nhwc_bf16");
  return TN_OK;
}

extern "C" int tn_poolbwd_nhwc_bf16(const void *a, const void *pooled, const void *dtop,
                                    int dtop_nchw_f32, void *gz, int B, int S, int M, int act,
                                    int act_nn, void *stream) {
  TN_REQUIRE(a && dtop && gz && B > 0 && S > 0 && M > 0, TN_ERR_ARG, "tn_poolbwd_nhwc_bf16: bad argument");
  TN_REQUIRE(act_is_fast(act), TN_ERR_UNSUPPORTED, "tn_poolbwd_nhwc_bf16: activation %d unsupported", act);
  TN_REQUIRE(!pooled || S % 2 == 0, TN_ERR_SHAPE, "tn_poolbwd_nhwc_bf16: odd size %d with a 2x2 pool", S);
  TN_REQUIRE(M % 8 == 0, TN_ERR_UNSUPPORTED, "tn_poolbwd_nhwc_bf16: needs M %% 8 == 0 (got %d)", M);
  const int64_t total8 = (int64_t)B * S * S * (M / 8);
  TN_REQUIRE(total8 < (1ll << 32), TN_ERR_UNSUPPORTED, "tn_poolbwd_nhwc_bf16: tensor too large");
  poolbwd_nhwc_kernel<<<blocks_for(total8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16 *)a, (const __nv_bfloat16 *)pooled, dtop, dtop_nchw_f32,
      (__nv_bfloat16 *)gz, (uint32_t)total8, S, M, FastDiv32(M / 8), FastDiv32(S),
      make_actk(act, act_nn));
  TN_LAUNCH_CHECK("tn_poolbwd_nhwc_bf16");
  return TN_OK;
}

extern "C" int tn_conv2d_tc_pack_weights(const float *W, void *Wp, int M, int C, int f, int dgrad,
                                         void *stream) {
  TN_REQUIRE(W && Wp && M > 0 && C > 0 && f > 0, TN_ERR_ARG, "tn_conv2d_tc_pack_weights: bad argument");
  pack_weights_kernel<<<blocks_for((int64_t)M * C * f * f), 256, 0, (cudaStream_t)stream>>>(
      W, (__nv_bfloat16 *)Wp, M, C, f, dgrad);
  TN_LAUNCH_CHECK("tn_conv2d_tc_pack_weights");
  return TN_OK;
}

extern "C" int tn_conv2d_tc_supported(int C, int S, int M, int f, int out_sz) {
  int th, tb;
  return C % 64 == 0 && M % 64 == 0 && f >= 1 && f <= 7 && out_sz == S /* 'same' */ &&
         tile_geom(out_sz, out_sz, &th, &tb, "tn_conv2d_tc_supported") == TN_OK &&
         ct_encode_fn() != nullptr;
}

extern "C" int tn_conv2d_tc_fprop(const void *x, const void *Wp, const float *bias, void *a,
                                  void *pooled, int B, int C, int S, int M, int f, int pad_lo,
                                  int out_sz, int act, int act_nn, void *stream) {
  const char *who = "tn_conv2d_tc_fprop";
  TN_REQUIRE(x && Wp && bias && a, TN_ERR_ARG, "%s: null argument", who);
  TN_REQUIRE(act_is_fast(act), TN_ERR_UNSUPPORTED, "%s: activation %d needs the CUDA-core path", who, act);
  TN_REQUIRE(!pooled || (out_sz % 2 == 0 && out_sz <= 16), TN_ERR_UNSUPPORTED,
             "%s: the fused 2x2 pool needs an even output size <= 16 (got %d)", who, out_sz);
  ConvTcArgs g{};
  g.B = B; g.Ho = out_sz; g.Wo = out_sz; g.C = C; g.N = M; g.f = f; g.pad = pad_lo;
  g.bias = bias; g.ak = make_actk(act, act_nn);
  g.out = (__nv_bfloat16 *)a; g.pooled = (__nv_bfloat16 *)pooled;
  return conv_tc_run(x, Wp, g, S, S, who, (cudaStream_t)stream);
}

extern "C" int tn_conv2d_tc_dgrad(const void *gz, const void *Wp_dgrad, void *dx, int B, int C,
                                  int S, int M, int f, int pad_lo, int out_sz, void *stream) {
  const char *who = "tn_conv2d_tc_dgrad";
  TN_REQUIRE(gz && Wp_dgrad && dx, TN_ERR_ARG, "%s: null argument", who);
  ConvTcArgs g{};
  // output of this GEMM is the layer INPUT (S x S, C channels); its input is gz (out_sz, M maps)
  g.B = B; g.Ho = S; g.Wo = S; g.C = M; g.N = C; g.f = f; g.pad = f - 1 - pad_lo;
  g.bias = nullptr; g.ak = make_actk(TN_ACT_LINEAR, 0);
  g.out = (__nv_bfloat16 *)dx; g.pooled = nullptr;
  return conv_tc_run(gz, Wp_dgrad, g, out_sz, out_sz, who, (cudaStream_t)stream);
}

static int wgrad_split(int tiles_out, int ptiles) {
  int ns = ceil_div(2 * kNumSM, tiles_out);
  if (ns > ptiles) ns = ptiles;
  if (ns > 64) ns = 64;
  return ns < 1 ? 1 : ns;
}

extern "C" size_t tn_conv2d_tc_wgrad_workspace_bytes(int B, int C, int M, int f, int out_sz) {
  int th = 1, tb = 1;
  if (tile_geom(out_sz, out_sz, &th, &tb, "tn_conv2d_tc_wgrad_workspace_bytes")) return 0;
  const int ptiles = ceil_div(B, tb) * (out_sz / th);
  const int tiles_out = f * f * (C / WG_BN) * ceil_div(M, 128);
  return ((size_t)wgrad_split(tiles_out, ptiles) * f * f * M * C + (size_t)kColSlices * M) * sizeof(float);
}

extern "C" int tn_conv2d_tc_wgrad(const void *x, const void *gz, float *dW, float *db,
                                  void *workspace, int B, int C, int S, int M, int f, int pad_lo,
                                  int out_sz, void *stream) {
  const char *who = "tn_conv2d_tc_wgrad";
  TN_REQUIRE(x && gz && dW && db && workspace, TN_ERR_ARG, "%s: null argument", who);
  TN_REQUIRE(C % 64 == 0 && M % 64 == 0, TN_ERR_UNSUPPORTED, "%s: needs C %% 64 == 0 and M %% 64 == 0", who);
  WgradTcArgs g{};
  g.B = B; g.Ho = out_sz; g.Wo = out_sz; g.M = M; g.C = C; g.f = f; g.pad = pad_lo;
  int rc = tile_geom(out_sz, out_sz, &g.TH, &g.TB, who);
  if (rc) return rc;
  g.ptiles = ceil_div(B, g.TB) * (out_sz / g.TH);
  const int tiles_out = f * f * (C / WG_BN) * ceil_div(M, 128);
  g.nsplit = wgrad_split(tiles_out, g.ptiles);
  g.partial = (float *)workspace;
  CUtensorMap tmG, tmX;
  rc = map_nhwc(&tmG, gz, B, out_sz, out_sz, M, out_sz, g.TH, g.TB, who);
  if (rc) return rc;
  rc = map_nhwc(&tmX, x, B, S, S, C, out_sz, g.TH, g.TB, who);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  dim3 grid(f * f * (C / WG_BN), ceil_div(M, 128), g.nsplit);
  conv_tc_wgrad_kernel<<<grid, CT_THREADS, WG_SMEM, st>>>(tmG, tmX, g);
  TN_LAUNCH_CHECK(who);
  conv_tc_wgrad_finish_kernel<<<blocks_for((int64_t)f * f * M * C), 256, 0, st>>>(
      (const float *)workspace, g.nsplit, M, C, f, dW);
  TN_LAUNCH_CHECK("tn_conv2d_tc_wgrad(finish)");
  float *cpart = (float *)workspace + (size_t)g.nsplit * f * f * M * C;
  colsum_bf16_kernel<<<kColSlices, 1024, 0, st>>>((const __nv_bfloat16 *)gz,
                                                   (int64_t)B * out_sz * out_sz, M, cpart);
  TN_LAUNCH_CHECK("tn_conv2d_tc_wgrad(db partial)");
  colsum_finish_kernel<<<ceil_div(M, 8), 256, 0, st>>>(cpart, kColSlices, M, db);
  TN_LAUNCH_CHECK("tn_conv2d_tc_wgrad(db)");
  return TN_OK;
}
