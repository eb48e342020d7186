// HiddenLayer matrix products on the 5th-generation tensor cores (theanet/layer/hidden.py:30-32
// tt.dot and its two gradients via tt.grad, layer.py:83).
//
// Two kernels live here.  gemm_tc_sk_kernel (further down: cluster split-K, partial tiles exchanged
// over distributed shared memory) is the default float32 path since round 2; gemm_tc_kernel, the
// first generation described next, serves the single-pass TF32 mode and stays as the A/B reference
// (tn_set_dense_mode).  Both share the tile formats, the epilogue and the host-side tensor maps.
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

#include <algorithm>
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
  long long *dbg;  // optional phase timestamps of gemm_tc_sk_kernel (tn_dense_debug_timestamps)
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
// mask * act' (backward-data), store.  `b4` = bias[n..n+3] (forward only; zeros where n+j >= N).
static __device__ __forceinline__ void tc_epi_quad_body(const TcArgs &g, int m, int n, float v0,
                                                        float v1, float v2, float v3, float4 b4,
                                                        uint32_t step, uint32_t sample0) {
  float v[4] = {v0, v1, v2, v3};
  const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
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
  // ReLU family (the shipped networks): selects on warp-uniform conditions instead of the
  // per-element switch of the general path (its BSSY / BRXU chains cost ~700 cycles per quad)
  const bool fam = g.ak.act == TN_ACT_LEAKY || g.ak.act == TN_ACT_LINEAR || g.ak.act == TN_ACT_RELU;
  if (g.epi == 0) {
    if (fam) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float zz = v[j] + bq[j];
        const float neg = g.ak.act == TN_ACT_LEAKY ? div100_rn(__fmul_rn(zz, g.ak.nn))
                                                   : (g.ak.act == TN_ACT_LINEAR ? zz : 0.f);
        const float a = zz > 0.f ? zz : neg;
        v[j] = (a * mk[j]) * g.scale;      // mk = 1 without dropout, scale = 1 when training: exact
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < g.N) {
          const float a = act_fwd_k(g.ak, v[j] + bq[j]);
          v[j] = g.mask_on ? a * mk[j] : a;
          if (g.scale != 1.f) v[j] *= g.scale;
        }
      }
    }
  } else if (g.epi == 1 && g.aux && fam) {
    float ax[4] = {0.f, 0.f, 0.f, 0.f};
    const float *ap = g.aux + (size_t)m * g.N + n;
    if (n + 3 < g.N) {
      const float4 a4 = *reinterpret_cast<const float4 *>(ap);
      ax[0] = a4.x; ax[1] = a4.y; ax[2] = a4.z; ax[3] = a4.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) ax[j] = ap[j];
    }
    const float s_pos = 1.f;
    const float s_neg = g.ak.act == TN_ACT_LEAKY ? g.ak.s_neg : (g.ak.act == TN_ACT_LINEAR ? 1.f : 0.f);
    const float s_zero = g.ak.act == TN_ACT_LEAKY ? g.ak.s_zero : (g.ak.act == TN_ACT_LINEAR ? 1.f : 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = ax[j] > 0.f ? s_pos : (ax[j] < 0.f ? s_neg : s_zero);
      v[j] = (v[j] * mk[j]) * d;
    }
  } else if (g.epi == 1 && g.aux) {
    float ax[4] = {0.f, 0.f, 0.f, 0.f};
    const float *ap = g.aux + (size_t)m * g.N + n;
    if (n + 3 < g.N) {
      const float4 a4 = *reinterpret_cast<const float4 *>(ap);
      ax[0] = a4.x; ax[1] = a4.y; ax[2] = a4.z; ax[3] = a4.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) ax[j] = ap[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < g.N) {
        const float d = act_bwd_k(g.ak, ax[j]);
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

// bias[n..n+3] of the forward epilogue (zeros otherwise / past the last column)
static __device__ __forceinline__ float4 tc_bias_quad(const TcArgs &g, int n) {
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.epi == 0) {
    if (n < g.N) b.x = g.bias[n];
    if (n + 1 < g.N) b.y = g.bias[n + 1];
    if (n + 2 < g.N) b.z = g.bias[n + 2];
    if (n + 3 < g.N) b.w = g.bias[n + 3];
  }
  return b;
}

// Out of line on purpose: unrolled into every 16-column chunk of the first-generation kernel's
// epilogue it made the kernel several times larger than the 32 KB instruction cache.
static __device__ __noinline__ void tc_epi_quad(const TcArgs &g, int m, int n, float v0, float v1,
                                                float v2, float v3, uint32_t step,
                                                uint32_t sample0) {
  tc_epi_quad_body(g, m, n, v0, v1, v2, v3, tc_bias_quad(g, n), step, sample0);
}
// the same with the bias quad already in registers (cluster kernel: one quad column per thread)
static __device__ __noinline__ void tc_epi_quad_b(const TcArgs &g, int m, int n, float4 v, float4 b4,
                                                  uint32_t step, uint32_t sample0) {
  tc_epi_quad_body(g, m, n, v.x, v.y, v.z, v.w, b4, step, sample0);
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
// Cluster split-K 3xTF32 variant (the default float32 path).  The per-k-block promotion above buys
// CUDA-core-grade sums at the price of a TMEM drain every 64 deep; it also forced narrow tiles (the
// same A tile re-read by up to 16 CTAs: ~120 MB of L2->SMEM traffic for a 1024x720x500 product).
// Here the K range is split over a thread-block CLUSTER along blockIdx.z: every CTA accumulates a
// 128 x BN tile over ~K/nsplit (a couple of hundred deep, where the tensor core's truncating
// accumulator costs ~1e-6) entirely in TMEM and parks the raw partial tile in its own shared
// memory.  After one cluster barrier each CTA owns 128/nsplit rows of the tile: it reads those rows
// from every peer's shared memory over DSMEM (ld.shared::cluster), adds them in split order with
// round-to-nearest float adds -- deterministic -- and runs the fused epilogue with coalesced
// float4 stores.  No workspace in HBM, no tickets, no serial last-CTA tail; wide tiles and one wave
// of CTAs whatever the shape.  (Pushing rows to their owners with st.shared::cluster instead was
// slower: a thread's TMEM row scatters 16-byte stores over the SM-to-SM network.)
//   warps 0..3 : TMA producers (k-blocks round-robin)        warp 8 : MMA issuer + TMEM owner
//   warps 4..7 : lo = x - hi transform of every landed stage, then the TMEM drain
//   all warps  : the cross-CTA reduction + epilogue of the CTA's row slice
// Everything a single thread must issue (TMA, tcgen05.mma, tcgen05.commit) sits in
// `if (elect_one())` inside warp-uniform control flow: a threadIdx-based `if (lane == 0)` makes
// nvcc wrap each UTCHMMA in an ELECT / BRA.U.ANY loop with R2UR moves (cuobjdump -sass).
// ---------------------------------------------------------------------------------------------
constexpr int SK_BK = 32;                       // one 128-byte swizzle row per operand row
constexpr int SK_A_BYTES = TC_BM * 128;
constexpr int SK_MAX_SPLIT = 8;                 // portable cluster size

template <int BN>
struct SkCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * (SK_A_BYTES + B_BYTES);   // hi + lo
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int NACC = 4;                 // independent accumulators (see TcCfg::NACC)
  static constexpr int TMEM_COLS = NACC * BN;    // 256 or 512
  static constexpr int QPR = BN / 4;             // float4 quads per tile row
  static_assert(TC_BM * BN * 4 <= STAGES * STAGE_BYTES, "partial tile must fit the stage ring");
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// optional phase timestamps (tools/gemm_phase_times.py): 16 x clock64 per CTA
#define TN_GSTAMP(i)                                                                              \
  do {                                                                                            \
    if (g.dbg) g.dbg[(size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (i)] = clock64(); \
  } while (0)

// Cross-CTA reduction + epilogue of one row slice (see the kernel): NSMAX >= ns peers, IB items
// per pass with all their DSMEM loads issued before the first add.
template <int BN, int NSMAX, int IB>
static __device__ __forceinline__ void sk_reduce(const TcArgs &g, uint32_t base, int ns, int z,
                                                 int r0, int items, int m0, int eq, int en,
                                                 float4 ebias, uint32_t step, uint32_t sample0) {
  constexpr int QPR = BN / 4;
  uint32_t peer[NSMAX];
#pragma unroll
  for (int sp = 0; sp < NSMAX; ++sp) peer[sp] = cluster_map(base, (uint32_t)min(sp, ns - 1));
#pragma unroll 1
  for (int it0 = threadIdx.x; it0 < items; it0 += IB * TC_THREADS) {
    float4 v[IB][NSMAX];
#pragma unroll
    for (int b = 0; b < IB; ++b) {
      const int it = it0 + b * TC_THREADS;
      const int r = r0 + it / QPR;
      const uint32_t off = (uint32_t)r * (BN * 4) + 16u * ((uint32_t)eq ^ ((uint32_t)r & (QPR - 1)));
#pragma unroll
      for (int sp = 0; sp < NSMAX; ++sp)
        if (sp < ns && it < items) {
          if (sp == z) {     // this CTA's own partial: plain shared-memory load, not the SM-to-SM path
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[b][sp].x), "=f"(v[b][sp].y), "=f"(v[b][sp].z), "=f"(v[b][sp].w)
                         : "r"(base + off)
                         : "memory");
          } else {
            v[b][sp] = ld_cluster_f4(peer[sp] + off);
          }
        }
    }
#pragma unroll
    for (int b = 0; b < IB; ++b) {
      const int it = it0 + b * TC_THREADS;
      if (it >= items) break;
      float4 sum = v[b][0];
#pragma unroll
      for (int sp = 1; sp < NSMAX; ++sp)
        if (sp < ns) { sum.x += v[b][sp].x; sum.y += v[b][sp].y; sum.z += v[b][sp].z; sum.w += v[b][sp].w; }
      if (en < g.N) tc_epi_quad_b(g, m0 + r0 + it / QPR, en, sum, ebias, step, sample0);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ TcArgs g) {
  using Cfg = SkCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr int NACC = Cfg::NACC;
  constexpr int QPR = Cfg::QPR;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + S * Cfg::STAGE_BYTES;  // full[S], xf[S], empty[S], accum, tmem slot
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto xf = [&](int s) { return bar0 + 8u * (S + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * S + s); };
  const uint32_t accum = bar0 + 8u * (3 * S);
  const uint32_t tslot = accum + 8u;
  auto stA = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto stB = [&](int s) { return base + s * Cfg::STAGE_BYTES + SK_A_BYTES; };
  auto stAlo = [&](int s) { return base + s * Cfg::STAGE_BYTES + SK_A_BYTES + Cfg::B_BYTES; };
  auto stBlo = [&](int s) { return base + s * Cfg::STAGE_BYTES + 2 * SK_A_BYTES + Cfg::B_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int nkb_all = (g.K + SK_BK - 1) / SK_BK;
  const int z = blockIdx.z, ns = gridDim.z;          // cluster = (1, 1, ns): z is the cluster rank
  const int kb0 = (int)(((long long)z * nkb_all) / ns);
  const int nkb = (int)(((long long)(z + 1) * nkb_all) / ns) - kb0;   // >= 1 (host: ns <= nkb_all)

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(full(s), 1);
      mbar_init(xf(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tslot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));
  pdl_wait();       // barriers, TMEM and descriptors are set up; global memory from here on
  if (threadIdx.x == 0) {
    TN_GSTAMP(0);
    TN_GSTAMP(1);
  }
  // epilogue operands of this thread's quad column, fetched while the pipeline fills
  static_assert(TC_THREADS % QPR == 0, "a thread keeps one quad column in the reduction");
  const int eq = threadIdx.x % QPR, en = n0 + 4 * eq;
  const float4 ebias = tc_bias_quad(g, en);
  uint32_t step = 0, sample0 = 0;
  if (g.epi != 2 && g.mask_on == 1) {
    step = (uint32_t)g.ctl[TN_CTL_STEP];
    sample0 = (uint32_t)g.ctl[TN_CTL_SAMPLE0];
  }

  if (warp < TC_PROD) {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      if (s % TC_PROD != warp) continue;       // a stage belongs to one producer warp
      if (elect_one()) {
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), SK_A_BYTES + Cfg::B_BYTES);
        const int k0 = (kb0 + kb) * SK_BK;
        if (!g.a_mn) {
          tma_load_2d(stA(s), &tmA, full(s), k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < TC_BM / 32; ++j)
            tma_load_2d(stA(s) + j * (SK_BK * 128), &tmA, full(s), m0 + 32 * j, k0);
        }
        if (!g.b_mn) {
          tma_load_2d(stB(s), &tmB, full(s), k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(stB(s) + j * (SK_BK * 128), &tmB, full(s), n0 + 32 * j, k0);
        }
      }
      __syncwarp();
    }
  } else if (warp == TC_MMA_WARP) {
    const uint32_t idesc = make_idesc(KIND_TF32, g.a_mn, g.b_mn, TC_BM, BN);
    const uint32_t a_lbo = g.a_mn ? SK_BK * 128u : 16u, b_lbo = g.b_mn ? SK_BK * 128u : 16u;
    const uint32_t a_step = g.a_mn ? 1024u : 32u, b_step = g.b_mn ? 1024u : 32u;
    const uint32_t a_sbo = g.a_mn ? 512u : 1024u, b_sbo = g.b_mn ? 512u : 1024u;
    const uint32_t a_lay = g.a_mn ? LAYOUT_SW128_32B : LAYOUT_SW128;
    const uint32_t b_lay = g.b_mn ? LAYOUT_SW128_32B : LAYOUT_SW128;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      mbar_wait_warp(xf(s), ph);
      tcgen05_fence_after();
      if (elect_one()) {
        // 12 MMAs per k-block dealt round-robin to the NACC accumulators; the first NACC of the
        // CTA's K range overwrite
        const uint32_t first = kb == 0 ? 0u : 1u;
#pragma unroll
        for (int j = 0; j < SK_BK / 8; ++j) {
          const uint64_t ad = make_smem_desc(stA(s) + a_step * j, a_lbo, a_sbo, a_lay);
          const uint64_t bd = make_smem_desc(stB(s) + b_step * j, b_lbo, b_sbo, b_lay);
          const uint64_t al = make_smem_desc(stAlo(s) + a_step * j, a_lbo, a_sbo, a_lay);
          const uint64_t bl = make_smem_desc(stBlo(s) + b_step * j, b_lbo, b_sbo, b_lay);
          // idx = 3j, 3j+1, 3j+2 -> accumulator idx % 4; accumulate unless one of the first four
          umma<KIND_TF32>(tmem_base + (uint32_t)(((3 * j) % NACC) * BN), al, bd, idesc,
                          (3 * j) >= NACC ? 1u : first);
          umma<KIND_TF32>(tmem_base + (uint32_t)(((3 * j + 1) % NACC) * BN), ad, bl, idesc,
                          (3 * j + 1) >= NACC ? 1u : first);
          umma<KIND_TF32>(tmem_base + (uint32_t)(((3 * j + 2) % NACC) * BN), ad, bd, idesc,
                          (3 * j + 2) >= NACC ? 1u : first);
        }
        umma_commit(empty(s));
        if (kb == nkb - 1) umma_commit(accum);
      }
      __syncwarp();
    }
  } else {
    // ===== lo = x - hi of every landed stage (the raw tile is the hi operand), then the drain =====
    const int et = threadIdx.x - 32 * TC_PROD;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      mbar_wait(full(s), ph);
      if (kb == 0 && et == 0) TN_GSTAMP(2);
      constexpr int NV = (SK_A_BYTES + Cfg::B_BYTES) / 16;
      const uint32_t hi0 = stA(s), lo0 = stAlo(s);
#pragma unroll 4
      for (int v = et; v < NV; v += 128) {
        uint32_t x0, x1, x2, x3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                     : "r"(hi0 + 16u * v));
        const float l0 = __uint_as_float(x0) - __uint_as_float(x0 & 0xffffe000u);
        const float l1 = __uint_as_float(x1) - __uint_as_float(x1 & 0xffffe000u);
        const float l2 = __uint_as_float(x2) - __uint_as_float(x2 & 0xffffe000u);
        const float l3 = __uint_as_float(x3) - __uint_as_float(x3 & 0xffffe000u);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo0 + 16u * v), "f"(l0),
                     "f"(l1), "f"(l2), "f"(l3)
                     : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(xf(s));
    }
    if (et == 0) TN_GSTAMP(3);
    mbar_wait(accum, 0);
    tcgen05_fence_after();
    if (et == 0) TN_GSTAMP(4);
    // row (q*32 + lane) of the partial tile -> this CTA's shared memory (the stage ring is idle:
    // every load has landed and every MMA has completed), quads XOR-swizzled by the row so that a
    // warp's 32 rows hit different banks
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int row = q * 32 + lane;
    const uint32_t prow = base + (uint32_t)row * (BN * 4);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t u[NACC][16];
#pragma unroll
      for (int a = 0; a < NACC; ++a) tmem_ld16(tlane + (uint32_t)(a * BN + c0), u[a]);
      tmem_ld_wait();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          r[i] = (__uint_as_float(u[0][4 * v4 + i]) + __uint_as_float(u[1][4 * v4 + i])) +
                 (__uint_as_float(u[2][4 * v4 + i]) + __uint_as_float(u[3][4 * v4 + i]));
        const uint32_t qs = (uint32_t)((c0 >> 2) + v4) ^ ((uint32_t)row & (QPR - 1));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(prow + 16u * qs), "f"(r[0]),
                     "f"(r[1]), "f"(r[2]), "f"(r[3])
                     : "memory");
      }
    }
    if (et == 0) TN_GSTAMP(5);
  }
  tcgen05_fence_before();
  cluster_sync_all();                      // every CTA's partial tile is parked and visible
  if (threadIdx.x == 0) TN_GSTAMP(6);
  if (warp == TC_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }

  // ---- this CTA's row slice: add the ns partials in split order (fixed: deterministic) + fused
  // epilogue.  item = (row, quad), quad fastest: TC_THREADS is a multiple of QPR, so a thread keeps
  // ONE quad column (its bias quad is already in registers).  A DSMEM load takes ~1000 cycles
  // under load: all the loads of up to 4 items (<= 16 float4) are in flight before the first add.
  const int rps = (TC_BM + ns - 1) / ns;           // rows per slice
  const int r0 = z * rps;
  const int r1 = min(min(r0 + rps, TC_BM), g.M - m0);
  const int items = max(r1 - r0, 0) * QPR;
  if (ns <= 4)
    sk_reduce<BN, 4, 4>(g, base, ns, z, r0, items, m0, eq, en, ebias, step, sample0);
  else
    sk_reduce<BN, 8, 2>(g, base, ns, z, r0, items, m0, eq, en, ebias, step, sample0);
  if (threadIdx.x == 0) TN_GSTAMP(7);
  cluster_sync_all();                      // nobody leaves while a peer may still read its tile
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

// ---- cluster split-K variant: work split and launch ------------------------------------------------
struct SkPlan {
  int BN, nsplit, tiles;
};

static int sk_env(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int BN>
static int sk_max_clusters(int ns) {
  static int cache[SK_MAX_SPLIT + 1] = {0};
  if (cache[ns]) return cache[ns];
  using Cfg = SkCfg<BN>;
  auto k = gemm_tc_sk_kernel<BN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, 1, ns);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = ns;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = kNumSM / ns;     // no device to ask (or an old driver): assume every SM can join a cluster
  }
  cache[ns] = n;
  return n;
}

// one wave of CTAs: tiles x splits ~ number of SMs, at least two k-blocks per split, the cluster
// small enough that every tile's cluster is co-resident
static SkPlan sk_plan(int M, int N, int K, int max_sms) {
  SkPlan p;
  p.BN = N > 64 ? 128 : 64;
  const int fb = sk_env("TN_SK_BN", 0);
  if (fb == 64 || fb == 128) p.BN = fb;
  p.tiles = ceil_div(M, TC_BM) * ceil_div(N, p.BN);
  const int nkb = ceil_div(K, SK_BK);
  const int sms = max_sms > 0 ? std::min(max_sms, kNumSM) : kNumSM;
  int ns = std::max(1, sms / p.tiles);
  ns = std::min(ns, std::max(1, nkb / 2));
  ns = std::min(ns, SK_MAX_SPLIT);
  const int fs = sk_env("TN_SK_SPLIT", 0);
  if (fs > 0) ns = std::min(std::min(fs, nkb), SK_MAX_SPLIT);
  else
    while (ns > 1 && (p.BN == 128 ? sk_max_clusters<128>(ns) : sk_max_clusters<64>(ns)) < p.tiles) --ns;
  p.nsplit = ns;
  return p;
}

static long long *g_gemm_dbg = nullptr;

template <int BN>
static int launch_sk(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcArgs &g, int mt, int nt,
                     int ns, const char *who, cudaStream_t st) {
  using Cfg = SkCfg<BN>;
  auto k = gemm_tc_sk_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nt, mt, ns);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = ns;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, k, tmA, tmB, g);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

// same contract as gemm_tc()
static int gemm_tc_sk(const float *A, int lda, int a_mn, const float *B, int ldb, int b_mn, TcArgs g,
                      int max_sms, const char *who, cudaStream_t st) {
  const SkPlan p = sk_plan(g.M, g.N, g.K, max_sms);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = tc_make_map_2d(&tmA, A, g.M, g.K, lda, 32, TC_BM, 0, who);
  else rc = tc_make_map_2d(&tmA, A, g.K, g.M, lda, 32, SK_BK, 1, who);
  if (rc) return rc;
  if (!b_mn) rc = tc_make_map_2d(&tmB, B, g.N, g.K, ldb, 32, p.BN, 0, who);
  else rc = tc_make_map_2d(&tmB, B, g.K, g.N, ldb, 32, SK_BK, 1, who);
  if (rc) return rc;
  g.a_mn = a_mn;
  g.b_mn = b_mn;
  g.dbg = g_gemm_dbg;
  const int mt = ceil_div(g.M, TC_BM), nt = ceil_div(g.N, p.BN);
  if (p.BN == 128) return launch_sk<128>(tmA, tmB, g, mt, nt, p.nsplit, who, st);
  return launch_sk<64>(tmA, tmB, g, mt, nt, p.nsplit, who, st);
}

void dense_tc_set_debug(long long *buf) { g_gemm_dbg = buf; }

static bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

bool dense_tc_ok(int n_in, int n_out, const void *p0, const void *p1, const void *p2) {
  // TMA needs 16-byte aligned bases and row strides (n_in, n_out multiples of 4 floats)
  return n_in % 4 == 0 && n_out % 4 == 0 && n_out >= 32 && n_in >= 32 && al16(p0) && al16(p1) &&
         al16(p2) && encode_fn() != nullptr;
}

int dense_tc_fwd(const float *x, const float *W, const float *bias, float *out, int B, int n_in,
                 int n_out, int act, float act_nn, int mask_on, uint32_t thr, uint64_t seed,
                 const int32_t *ctl, const float *mask_inj, float scale, int split, int max_sms,
                 cudaStream_t st) {
  TcArgs g{};
  g.C = out; g.ldc = n_out; g.M = B; g.N = n_out; g.K = n_in;
  g.epi = 0; g.bias = bias; g.mask_inj = mask_inj; g.ctl = ctl; g.seed = seed; g.thr = thr;
  g.mask_on = mask_on; g.ak = make_actk(act, (int)act_nn); g.scale = scale;
  // A = x (B x n_in, K contiguous); B[n][k] = W[k][n] (N contiguous)
  if (split == 1) return gemm_tc_sk(x, n_in, 0, W, n_out, 1, g, max_sms, "tn_dense_fwd(tc,split-K)", st);
  return gemm_tc(x, n_in, 0, W, n_out, 1, g, split, "tn_dense_fwd(tc)", st);
}

int dense_tc_bwd_data(const float *gr, const float *W, float *dx, int B, int n_in, int n_out,
                      const float *prev_out, int act, float act_nn, int mask_on, uint32_t thr,
                      uint64_t seed, const int32_t *ctl, const float *mask_inj, int split,
                      int max_sms, cudaStream_t st) {
  TcArgs g{};
  g.C = dx; g.ldc = n_in; g.M = B; g.N = n_in; g.K = n_out;
  g.epi = 1; g.aux = prev_out; g.mask_inj = mask_inj; g.ctl = ctl; g.seed = seed; g.thr = thr;
  g.mask_on = mask_on; g.ak = make_actk(act, (int)act_nn); g.scale = 1.f;
  // A = g (B x n_out, K contiguous); B[n][k] = W[n][k] (K contiguous)
  if (split == 1)
    return gemm_tc_sk(gr, n_out, 0, W, n_out, 0, g, max_sms, "tn_dense_bwd_data(tc,split-K)", st);
  return gemm_tc(gr, n_out, 0, W, n_out, 0, g, split, "tn_dense_bwd_data(tc)", st);
}

int dense_tc_bwd_weights(const float *x, const float *gr, float *dW, int B, int n_in, int n_out,
                         int split, int max_sms, cudaStream_t st) {
  TcArgs g{};
  g.C = dW; g.ldc = n_out; g.M = n_in; g.N = n_out; g.K = B;
  g.epi = 2; g.scale = 1.f; g.ak = make_actk(TN_ACT_LINEAR, 0);
  // A[m][k] = x[k][m] (M contiguous); B[n][k] = g[k][n] (N contiguous)
  if (split == 1)
    return gemm_tc_sk(x, n_in, 1, gr, n_out, 1, g, max_sms, "tn_dense_bwd_weights(tc,split-K)", st);
  return gemm_tc(x, n_in, 1, gr, n_out, 1, g, split, "tn_dense_bwd_weights(tc)", st);
}

}  // namespace tn
