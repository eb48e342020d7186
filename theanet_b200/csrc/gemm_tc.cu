// HiddenLayer matrix products on the 5th-generation tensor cores (theanet/layer/hidden.py:30-32
// tt.dot and its two gradients via tt.grad, layer.py:83).
//
// One warp-specialised kernel serves out = x.W, dx = g.W^T and dW = x^T.g:
//   warps 0..3  : TMA producers -- cp.async.bulk.tensor tiles (128-byte swizzle) into a ring of
//                 shared-memory stages, completion on mbarriers; k-blocks round-robin;
//   warp 8      : allocates TMEM, one lane issues tcgen05.mma (kind::tf32, fp32 accumulate in
//                 TMEM) and hands stages back with tcgen05.commit;
//   warps 4..7  : epilogue -- tcgen05.ld the 128 x BN accumulator, fuse bias / activation /
//                 Philox dropout mask (forward) or mask * act' (backward-data), store fp32.
// Operands are consumed in the layout theanet's .pkl exposes (row-major x (B, n_in), W (n_in,
// n_out)): whichever of M/N/K is contiguous in memory, the tile is described to the tensor core as
// K-major or MN-major, so no transposed copy of anything is ever made.
//
// float32 parity (SURVEY.md 7 "hard parts" #2): kind::tf32 reads only the top 19 bits of each
// operand.  In the default 3xTF32 mode the epilogue warps double as a transform stage: they split
// every landed tile into hi = top 19 bits and lo = x - hi (exact), and the MMA lane accumulates
// A_lo.B_hi + A_hi.B_lo + A_hi.B_hi -- float32-grade products at tensor-core rate.  The tensor
// core's accumulator truncates when it aligns addends, a bias that grows linearly with K (measured:
// 5e-6 relative at K=720, 2e-5 at K=4500), so in this mode TMEM only ever holds the partial sum of
// ONE 32-deep k-block (two buffers, ping-pong): the same warps pull each finished block into
// float32 registers with round-to-nearest adds ("promotion"), which brings the result to within
// CUDA-core SGEMM accuracy (~3e-7).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "dense_tc.cuh"
#include "tc_ptx.cuh"

namespace tn {
using namespace tc;

constexpr int TC_BM = 128;       // accumulator rows = TMEM lanes
constexpr int TC_KA = 2;         // 128-byte swizzle rows (32 fp32) per operand row and stage
constexpr int TC_BK = 32 * TC_KA;  // k-block depth: the fixed cost of a pipeline stage (mbarrier
                                   // round trip, ~0.35 us measured) is amortised over 64-deep blocks
constexpr int TC_PROD = 4;       // TMA producer warps (warps 0..3), k-blocks round-robin: a warp
                                 // gets one mbarrier phase per ~0.36 us (tools/micro/tma_bench2.cu),
                                 // throughput scales with the number of issuing warps
constexpr int TC_MMA_WARP = 2 * TC_PROD;           // warps 4..7: epilogue, warp 8: MMA + TMEM
constexpr int TC_THREADS = 32 * (TC_MMA_WARP + 1);
constexpr int TC_A_BYTES = TC_BM * 128 * TC_KA;

struct TcArgs {
  float *C;
  int ldc, M, N, K;
  int a_mn, b_mn;  // operand majors (1 = M/N contiguous in memory)
  int epi;         // 0 forward, 1 backward-data, 2 plain store
  const float *bias, *aux, *mask_inj;
  const int32_t *ctl;
  uint64_t seed;
  uint32_t thr;
  int mask_on;
  ActK ak;
  float scale;
};

// kind::tf32 reads the top 19 bits of a 32-bit operand and ignores the 13 below: the raw tile the
// TMA landed IS the hi operand (x & 0xffffe000 as far as the tensor core can tell), so the split
// only has to write lo = x - hi.  Set to 1 to store the masked hi tile back as well (A/B check;
// tools/gemm_tc_check.py shows the same 4e-7 error either way, a rounding unit would show ~2e-4).
#ifndef TN_TC_STORE_HI
#define TN_TC_STORE_HI 0
#endif

template <int BN, int SPLIT>
struct TcCfg {
  static constexpr int B_BYTES = BN * 128 * TC_KA;
  static constexpr int STAGE_BYTES = (TC_A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : (STAGES_RAW < 1 ? 1 : STAGES_RAW);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // Back-to-back tcgen05.mma on ONE accumulator are serialised by the accumulate dependency
  // (~112 cycles each whatever N <= 128 is; tools/micro/mma_bench.cu).  The MMAs of a k-block are
  // therefore dealt round-robin to NACC independent accumulators that the epilogue warps add up.
  static constexpr int NACC = BN <= 64 ? 4 : 2;
  static constexpr int TMEM_COLS = (SPLIT ? 2 : 1) * NACC * BN;  // power of two, <= 512
};

// Epilogue of 4 consecutive columns n..n+3 of row m: dropout mask, bias + activation (forward) or
// mask * act' (backward-data), store.  Out of line on purpose: unrolled into every 16-column
// chunk it made the kernel several times larger than the 32 KB instruction cache.
static __device__ __noinline__ void tc_epi_quad(const TcArgs &g, int m, int n, float v0, float v1,
                                                float v2, float v3, uint32_t step,
                                                uint32_t sample0) {
  float v[4] = {v0, v1, v2, v3};
  float mk[4] = {1.f, 1.f, 1.f, 1.f};
  if (g.epi != 2) {
    if (g.mask_on == 1) {
      const Philox4 pr = philox_block(g.seed, TN_RNG_DROPOUT, step, sample0 + (uint32_t)m,
                                      (uint32_t)(n >> 2));
      mk[0] = pr.x < g.thr ? 1.f : 0.f;
      mk[1] = pr.y < g.thr ? 1.f : 0.f;
      mk[2] = pr.z < g.thr ? 1.f : 0.f;
      mk[3] = pr.w < g.thr ? 1.f : 0.f;
    } else if (g.mask_on == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) mk[j] = g.mask_inj[(size_t)m * g.N + n + j];
    }
  }
  if (g.epi == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < g.N) {
        const float a = act_fwd_k(g.ak, v[j] + g.bias[n + j]);
        v[j] = g.mask_on ? a * mk[j] : a;
        if (g.scale != 1.f) v[j] *= g.scale;
      }
    }
  } else if (g.epi == 1 && g.aux) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < g.N) {
        const float d = act_bwd_k(g.ak, g.aux[(size_t)m * g.N + n + j]);
        v[j] = (g.mask_on ? v[j] * mk[j] : v[j]) * d;
      }
    }
  }
  float *c = g.C + (size_t)m * g.ldc + n;
  if (n + 3 < g.N) {
    *reinterpret_cast<float4 *>(c) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < g.N) c[j] = v[j];
  }
}

template <int BN, int SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcArgs g) {
  using Cfg = TcCfg<BN, SPLIT>;
  constexpr int S = Cfg::STAGES;
  constexpr int NACC = Cfg::NACC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + S * Cfg::STAGE_BYTES;  // full[S], xf[S], empty[S], accum, tmem slot
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto xf = [&](int s) { return bar0 + 8u * (S + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * S + s); };
  const uint32_t accum = bar0 + 8u * (3 * S);
  auto tfull = [&](int b) { return accum + 8u * (1 + b); };    // SPLIT: k-block partial ready
  auto tempty = [&](int b) { return accum + 8u * (3 + b); };   // SPLIT: TMEM buffer drained
  const uint32_t tslot = accum + 8u * 5;
  auto stA = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto stB = [&](int s) { return base + s * Cfg::STAGE_BYTES + TC_A_BYTES; };
  auto stAlo = [&](int s) { return base + s * Cfg::STAGE_BYTES + TC_A_BYTES + Cfg::B_BYTES; };
  auto stBlo = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * TC_A_BYTES + Cfg::B_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int nkb = (g.K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), 1);
      mbar_init(xf(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tslot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (warp < TC_PROD) {
    // ===== TMA producers =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        // a stage belongs to ONE producer warp, so its phases are waited on in order (a warp
        // running two phases ahead on a shared stage would pass the parity test spuriously)
        if (s % TC_PROD != warp) continue;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), TC_A_BYTES + Cfg::B_BYTES);
        const int k0 = kb * TC_BK;
        if (!g.a_mn) {   // K-major: one (32 x 128 rows) box per 128-byte k-atom
#pragma unroll
          for (int a = 0; a < TC_KA; ++a)
            tma_load_2d(stA(s) + a * (TC_BM * 128), &tmA, full(s), k0 + 32 * a, m0);
        } else {         // M-major: one (32 wide x BK deep) box per 32-row column block
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j)
            tma_load_2d(stA(s) + j * (TC_BK * 128), &tmA, full(s), m0 + 32 * j, k0);
        }
        if (!g.b_mn) {
#pragma unroll
          for (int a = 0; a < TC_KA; ++a)
            tma_load_2d(stB(s) + a * (BN * 128), &tmB, full(s), k0 + 32 * a, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(stB(s) + j * (TC_BK * 128), &tmB, full(s), n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(KIND_TF32, g.a_mn, g.b_mn, TC_BM, BN);
      // per 8-deep k-step: K-major advances 32 B inside the swizzle row, MN-major one 1 KB group
      const uint32_t a_lbo = g.a_mn ? TC_BK * 128u : 16u, b_lbo = g.b_mn ? TC_BK * 128u : 16u;
      // byte offset of 8-deep k-step j inside a stage tile
      auto a_off = [&](int j) -> uint32_t {
        return g.a_mn ? 1024u * j : (uint32_t)((j >> 2) * (TC_BM * 128) + (j & 3) * 32);
      };
      auto b_off = [&](int j) -> uint32_t {
        return g.b_mn ? 1024u * j : (uint32_t)((j >> 2) * (BN * 128) + (j & 3) * 32);
      };
      // MN-major tf32 tiles use the 32-byte-atom swizzle (4-row atoms), K-major the plain one
      const uint32_t a_sbo = g.a_mn ? 512u : 1024u, b_sbo = g.b_mn ? 512u : 1024u;
      const uint32_t a_lay = g.a_mn ? LAYOUT_SW128_32B : LAYOUT_SW128;
      const uint32_t b_lay = g.b_mn ? LAYOUT_SW128_32B : LAYOUT_SW128;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(SPLIT ? xf(s) : full(s), ph);
        const int tb = kb & 1;                       // SPLIT: TMEM ping-pong buffer
        if (SPLIT) mbar_wait(tempty(tb), ((uint32_t)(kb >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t td = tmem_base + (SPLIT ? (uint32_t)(tb * NACC * BN) : 0u);
        int idx = 0;   // MMA number inside the k-block -> accumulator idx % NACC
        auto issue = [&](uint64_t a, uint64_t b) {
          const uint32_t acc_on = SPLIT ? (idx >= NACC) : (kb > 0 || idx >= NACC);
          umma<KIND_TF32>(td + (uint32_t)((idx % NACC) * BN), a, b, idesc, acc_on ? 1u : 0u);
          ++idx;
        };
#pragma unroll
        for (int j = 0; j < TC_BK / 8; ++j) {
          const uint64_t ad = make_smem_desc(stA(s) + a_off(j), a_lbo, a_sbo, a_lay);
          const uint64_t bd = make_smem_desc(stB(s) + b_off(j), b_lbo, b_sbo, b_lay);
          if (SPLIT) {
            const uint64_t al = make_smem_desc(stAlo(s) + a_off(j), a_lbo, a_sbo, a_lay);
            const uint64_t bl = make_smem_desc(stBlo(s) + b_off(j), b_lbo, b_sbo, b_lay);
            issue(al, bd);   // every k-block starts from zero (promotion, see header)
            issue(ad, bl);
            issue(ad, bd);
          } else {
            issue(ad, bd);
          }
        }
        umma_commit(empty(s));  // stage reusable once these MMAs have read it
        if (SPLIT) umma_commit(tfull(tb));
      }
      if (!SPLIT) umma_commit(accum);
    }
  } else {
    // ===== transform + promotion (3xTF32) and epilogue: warps 4..7 =====
    const int et = threadIdx.x - 32 * TC_PROD;  // 0..127
    const int q = warp & 3;           // TMEM lane quarter this warp may read
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float acc[SPLIT ? BN : 1];
    if (SPLIT) {
#pragma unroll
      for (int i = 0; i < BN; ++i) acc[i] = 0.f;
      // acc += partial sum of k-block j (TMEM buffer j & 1), then hand the buffer back
      auto drain = [&](int j) {
        const int tb = j & 1;
        mbar_wait(tfull(tb), (uint32_t)(j >> 1) & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float part[16];
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            uint32_t r[16];
            tmem_ld16(tlane + (uint32_t)((tb * NACC + a) * BN + c0), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              part[i] = a ? part[i] + __uint_as_float(r[i]) : __uint_as_float(r[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c0 + i] += part[i];
        }
        tcgen05_fence_before();
        mbar_arrive(tempty(tb));
      };
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(full(s), ph);
        // the A and B tiles are adjacent: one pass over (A_BYTES + B_BYTES)/16 vectors; the lo
        // twin of every vector sits at the same offset in the lo half of the stage
        constexpr int NV = (TC_A_BYTES + Cfg::B_BYTES) / 16;
        const uint32_t hi0 = stA(s), lo0 = stAlo(s);
#pragma unroll 4
        for (int v = et; v < NV; v += 128) {
          uint32_t x0, x1, x2, x3;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                       : "r"(hi0 + 16u * v));
          const uint32_t h0 = x0 & 0xffffe000u, h1 = x1 & 0xffffe000u, h2 = x2 & 0xffffe000u,
                         h3 = x3 & 0xffffe000u;
          const float l0 = __uint_as_float(x0) - __uint_as_float(h0);
          const float l1 = __uint_as_float(x1) - __uint_as_float(h1);
          const float l2 = __uint_as_float(x2) - __uint_as_float(h2);
          const float l3 = __uint_as_float(x3) - __uint_as_float(h3);
#if TN_TC_STORE_HI
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hi0 + 16u * v), "r"(h0),
                       "r"(h1), "r"(h2), "r"(h3)
                       : "memory");
#endif
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo0 + 16u * v), "f"(l0),
                       "f"(l1), "f"(l2), "f"(l3)
                       : "memory");
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
        mbar_arrive(xf(s));
        if (kb > 0) drain(kb - 1);  // overlaps the MMAs of block kb
      }
      drain(nkb - 1);
    } else {
      mbar_wait(accum, 0);
      tcgen05_fence_after();
    }
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < g.M;
    uint32_t step = 0, sample0 = 0;
    if (g.epi != 2 && g.mask_on == 1) {
      step = (uint32_t)g.ctl[TN_CTL_STEP];
      sample0 = (uint32_t)g.ctl[TN_CTL_SAMPLE0];
    }
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n0 + c0 >= g.N) break;  // warp-uniform
      float r[16];
      if (SPLIT) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = acc[c0 + i];
      } else {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          uint32_t u[16];
          tmem_ld16(tlane + (uint32_t)(a * BN + c0), u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = a ? r[i] + __uint_as_float(u[i]) : __uint_as_float(u[i]);
        }
      }
      if (!row_ok) continue;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        const int n = n0 + c0 + 4 * v4;
        if (n >= g.N) break;
        tc_epi_quad(g, m, n, r[4 * v4], r[4 * v4 + 1], r[4 * v4 + 2], r[4 * v4 + 3], step, sample0);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// 2-D fp32 tensor map over a row-major (rows, cols) matrix with leading dimension ld (elements),
// box = (box_cols, box_rows), 128-byte swizzle (atom32: 32-byte swizzle atoms, for MN-major tf32
// tiles), out-of-bounds elements read as zero
int tc_make_map_2d(CUtensorMap *map, const float *ptr, int rows, int cols, int ld, int box_cols,
                   int box_rows, int atom32, const char *who) {
  EncodeTiledFn fn = encode_fn();
  TN_REQUIRE(fn, TN_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available from the driver", who);
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TN_REQUIRE(r == CUDA_SUCCESS, TN_ERR_CUDA,
             "%s: cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%dx%d", who, (int)r,
             rows, cols, ld, box_cols, box_rows);
  return TN_OK;
}

template <int BN, int SPLIT>
static int launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcArgs &g,
                     const char *who, cudaStream_t st) {
  using Cfg = TcCfg<BN, SPLIT>;
  auto k = gemm_tc_kernel<BN, SPLIT>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, TC_BM));
  k<<<grid, TC_THREADS, Cfg::SMEM, st>>>(tmA, tmB, g);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

// D (M x N, ldc) = A . B^T with A (M x K) and B (N x K) described by (ptr, ld, mn-major flag):
//   !mn : element (r, k) at ptr[r*ld + k]     mn : element (r, k) at ptr[k*ld + r]
static int gemm_tc(const float *A, int lda, int a_mn, const float *B, int ldb, int b_mn, TcArgs g,
                   int split, const char *who, cudaStream_t st) {
  // tile width: the widest BN that still gives ~one CTA per SM-pair's worth of parallelism
  const int mt = ceil_div(g.M, TC_BM);
  int BN = 128;
  while (BN > 32 && mt * ceil_div(g.N, BN) < 96) BN >>= 1;
  if (const char *e = getenv("TN_TC_BN")) {  // tuning knob: force the tile width (32, 64, 128)
    const int v = atoi(e);
    if (v == 32 || v == 64 || v == 128) BN = v;
  }
  if (split && BN > 64) BN = 64;   // hi + lo copies of a 128-wide stage would leave one stage
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = tc_make_map_2d(&tmA, A, g.M, g.K, lda, 32, TC_BM, 0, who);
  else rc = tc_make_map_2d(&tmA, A, g.K, g.M, lda, 32, TC_BK, 1, who);
  if (rc) return rc;
  if (!b_mn) rc = tc_make_map_2d(&tmB, B, g.N, g.K, ldb, 32, BN, 0, who);
  else rc = tc_make_map_2d(&tmB, B, g.K, g.N, ldb, 32, TC_BK, 1, who);
  if (rc) return rc;
  g.a_mn = a_mn;
  g.b_mn = b_mn;
  if (split) {
    switch (BN) {
      case 64: return launch_tc<64, 1>(tmA, tmB, g, who, st);
      default: return launch_tc<32, 1>(tmA, tmB, g, who, st);
    }
  }
  switch (BN) {
    case 128: return launch_tc<128, 0>(tmA, tmB, g, who, st);
    case 64: return launch_tc<64, 0>(tmA, tmB, g, who, st);
    default: return launch_tc<32, 0>(tmA, tmB, g, who, st);
  }
}

static bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

bool dense_tc_ok(int n_in, int n_out, const void *p0, const void *p1, const void *p2) {
  // TMA needs 16-byte aligned bases and row strides (n_in, n_out multiples of 4 floats)
  return n_in % 4 == 0 && n_out % 4 == 0 && n_out >= 32 && n_in >= 32 && al16(p0) && al16(p1) &&
         al16(p2) && encode_fn() != nullptr;
}

int dense_tc_fwd(const float *x, const float *W, const float *bias, float *out, int B, int n_in,
                 int n_out, int act, float act_nn, int mask_on, uint32_t thr, uint64_t seed,
                 const int32_t *ctl, const float *mask_inj, float scale, int split,
                 cudaStream_t st) {
  TcArgs g{};
  g.C = out; g.ldc = n_out; g.M = B; g.N = n_out; g.K = n_in;
  g.epi = 0; g.bias = bias; g.mask_inj = mask_inj; g.ctl = ctl; g.seed = seed; g.thr = thr;
  g.mask_on = mask_on; g.ak = make_actk(act, (int)act_nn); g.scale = scale;
  // A = x (B x n_in, K contiguous); B[n][k] = W[k][n] (N contiguous)
  return gemm_tc(x, n_in, 0, W, n_out, 1, g, split, "tn_dense_fwd(tc)", st);
}

int dense_tc_bwd_data(const float *gr, const float *W, float *dx, int B, int n_in, int n_out,
                      const float *prev_out, int act, float act_nn, int mask_on, uint32_t thr,
                      uint64_t seed, const int32_t *ctl, const float *mask_inj, int split,
                      cudaStream_t st) {
  TcArgs g{};
  g.C = dx; g.ldc = n_in; g.M = B; g.N = n_in; g.K = n_out;
  g.epi = 1; g.aux = prev_out; g.mask_inj = mask_inj; g.ctl = ctl; g.seed = seed; g.thr = thr;
  g.mask_on = mask_on; g.ak = make_actk(act, (int)act_nn); g.scale = 1.f;
  // A = g (B x n_out, K contiguous); B[n][k] = W[n][k] (K contiguous)
  return gemm_tc(gr, n_out, 0, W, n_out, 0, g, split, "tn_dense_bwd_data(tc)", st);
}

int dense_tc_bwd_weights(const float *x, const float *gr, float *dW, int B, int n_in, int n_out,
                         int split, cudaStream_t st) {
  TcArgs g{};
  g.C = dW; g.ldc = n_out; g.M = n_in; g.N = n_out; g.K = B;
  g.epi = 2; g.scale = 1.f; g.ak = make_actk(TN_ACT_LINEAR, 0);
  // A[m][k] = x[k][m] (M contiguous); B[n][k] = g[k][n] (N contiguous)
  return gemm_tc(x, n_in, 1, gr, n_out, 1, g, split, "tn_dense_bwd_weights(tc)", st);
}

}  // namespace tn
