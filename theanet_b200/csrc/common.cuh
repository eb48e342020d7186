// Shared device helpers: error plumbing, activations (theanet/layer/layer.py:27-39) and the
// Philox4x32-10 streams (CPU twin: oracle/philox.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/theanet_b200.h"

namespace tn {

void set_error(const char *fmt, ...);
void count_launch();   // bumps the counter behind tn_launch_count()

#define TN_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      tn::set_error(__VA_ARGS__);        \
      return (code);                     \
    }                                    \
  } while (0)

#define TN_LAUNCH_CHECK(name)                                                     \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      tn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return TN_ERR_CUDA;                                                         \
    }                                                                             \
    tn::count_launch();                                                           \
  } while (0)

constexpr int kNumSM = 148;

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// A training step is a chain of a dozen short kernels: the ~2 us between the end of one and the
// first instruction of the next (launch latency inside a CUDA graph, then barrier / TMEM / index
// setup) is a tenth of the step.  Kernels launched through launch_pdl() are scheduled while their
// predecessor in the stream drains; they do whatever needs no global memory and pdl_wait() before
// the first global access -- the wait returns once the predecessor has completed and flushed, so
// the memory semantics are exactly those of a plain launch.  TN_PDL=0 turns the attribute off
// (A/B runs: 0.176 -> 0.166 ms per C2 step).  An explicit early trigger
// (griddepcontrol.launch_dependents at kernel entry, -DTN_PDL_TRIGGER=1) lets the successor's
// CTAs become resident even earlier, where they only compete with the side-branch kernels for
// SMs: measured 0.183 ms, so the implicit trigger at kernel exit is what ships.
#ifndef TN_PDL_TRIGGER
#define TN_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if TN_PDL_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// Exact unsigned division of any 32-bit n by a run-time constant d (Granlund-Montgomery round-up
// method): one multiply-high, one add, two shifts instead of the ~30-instruction emulated divide.
struct FastDiv32 {
  uint32_t d, m, s;
  FastDiv32() : d(1), m(0), s(0) {}
  explicit FastDiv32(uint32_t dd) : d(dd), m(0), s(0) {
    if (dd > 1) {
      while ((1ull << s) < dd) ++s;
      m = (uint32_t)(((1ull << 32) * ((1ull << s) - dd)) / dd + 1);
    }
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    if (d == 1) return n;
    const uint32_t t = __umulhi(n, m);
    return (t + ((n - t) >> 1)) >> (s - 1);
  }
};

// ------------------------------------------------------------------------------------------------
// Activations.  Forward follows layer.py:36 literally (max(0,x) + (min(0,x)*NN)/100, each op
// rounded in float32) so reluNN is bit-exact against the numpy restatement.
// ------------------------------------------------------------------------------------------------
// t / 100 correctly rounded without the division unit: q0 = t*RN(1/100), one FMA residual, one FMA
// correction (Markstein).  Checked exhaustively over whole binades against IEEE division
// (tools notes in DESIGN.md); exponent extremes fall back to __fdiv_rn.
__device__ __forceinline__ float div100_rn(float t) {
  const float q0 = __fmul_rn(t, 0.01f);
  const float e = __fmaf_rn(-100.f, q0, t);
  const float q = __fmaf_rn(e, 0.01f, q0);
  const float at = fabsf(t);
  if (at > 1e-28f && at < 1e30f) return q;
  return t == 0.f ? t : __fdiv_rn(t, 100.f);
}

// The transcendental activations live out of line: inlined into every unrolled epilogue they made
// the kernels 100-370 KB of SASS, far beyond the 32 KB instruction cache, and the epilogues ran at
// instruction-fetch speed (DESIGN.md "measured").  ReLU-family and linear stay inline.
static __device__ __noinline__ float act_fwd_slow(float z, int act) {
  switch (act) {
    case TN_ACT_TANH: return tanhf(z);
    case TN_ACT_SCALED_TANH: return 1.7f * tanhf(__fdiv_rn(2.f * z, 3.f));
    case TN_ACT_SIGMOID: return __fdiv_rn(1.f, 1.f + expf(-z));
    case TN_ACT_SOFTPLUS: return z > 0.f ? z + log1pf(expf(-z)) : log1pf(expf(z));
  }
  return z;
}
static __device__ __noinline__ float act_bwd_slow(float a, int act) {
  switch (act) {
    case TN_ACT_TANH: return 1.f - a * a;
    case TN_ACT_SCALED_TANH: {
      const float t = __fdiv_rn(a, 1.7f);
      return (1.7f * 2.f / 3.f) * (1.f - t * t);
    }
    case TN_ACT_SIGMOID: return a * (1.f - a);
    case TN_ACT_SOFTPLUS: return 1.f - expf(-a);
  }
  return 1.f;
}

__device__ __forceinline__ float act_fwd(float z, int act, float nn) {
  if (act == TN_ACT_LEAKY) return z > 0.f ? z : div100_rn(__fmul_rn(fminf(0.f, z), nn));
  if (act == TN_ACT_LINEAR) return z;
  if (act == TN_ACT_RELU) return fmaxf(0.f, z);
  return act_fwd_slow(z, act);
}

// d act / dz expressed through the stored output a = act(z).  For reluNN (NN >= 1) sign(a) ==
// sign(z), and at exactly 0 Theano's maximum/minimum gradients both fire (slope 1 + NN/100).
// For relu / relu00 a == 0 stands for every z <= 0 and gets slope 0 (z == 0 exactly is the only
// deviation from Theano, a measure-zero event).
__device__ __forceinline__ float act_bwd_from_out(float a, int act, float nn) {
  if (act == TN_ACT_LEAKY) {
    if (nn == 0.f) return a > 0.f ? 1.f : 0.f;
    const float s = div100_rn(nn);
    return a > 0.f ? 1.f : (a < 0.f ? s : 1.f + s);
  }
  if (act == TN_ACT_LINEAR) return 1.f;
  if (act == TN_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  return act_bwd_slow(a, act);
}

// Activation bundle with the leaky-ReLU slopes hoisted to the host: the common case costs a
// uniform compare instead of the 7-way switch (which nvcc turns into an indirect branch per element).
struct ActK {
  int act;
  float nn, s_neg, s_zero;
};
inline ActK make_actk(int act, int act_nn) {
  ActK k;
  k.act = act;
  k.nn = (float)act_nn;
  const float sl = (float)act_nn / 100.f;
  k.s_neg = act_nn == 0 ? 0.f : sl;
  k.s_zero = act_nn == 0 ? 0.f : 1.f + sl;
  return k;
}
inline bool act_is_fast(int act) {
  return act == TN_ACT_LEAKY || act == TN_ACT_LINEAR || act == TN_ACT_RELU;
}
// GEN = false: the kernel is only launched with ReLU-family / linear activations and contains no
// call at all (a possible call in a hot kernel costs registers even when it is never taken)
template <bool GEN>
__device__ __forceinline__ float act_fwd_t(const ActK &k, float z) {
  if (k.act == TN_ACT_LEAKY) return z > 0.f ? z : div100_rn(__fmul_rn(z, k.nn));
  if (k.act == TN_ACT_LINEAR) return z;
  if (!GEN || k.act == TN_ACT_RELU) return fmaxf(0.f, z);
  return act_fwd_slow(z, k.act);
}
template <bool GEN>
__device__ __forceinline__ float act_bwd_t(const ActK &k, float a) {
  if (k.act == TN_ACT_LEAKY) return a > 0.f ? 1.f : (a < 0.f ? k.s_neg : k.s_zero);
  if (k.act == TN_ACT_LINEAR) return 1.f;
  if (!GEN || k.act == TN_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  return act_bwd_slow(a, k.act);
}
__device__ __forceinline__ float act_fwd_k(const ActK &k, float z) {
  if (k.act == TN_ACT_LEAKY) return z > 0.f ? z : div100_rn(__fmul_rn(z, k.nn));
  if (k.act == TN_ACT_LINEAR) return z;
  if (k.act == TN_ACT_RELU) return fmaxf(0.f, z);
  return act_fwd_slow(z, k.act);
}
__device__ __forceinline__ float act_bwd_k(const ActK &k, float a) {
  if (k.act == TN_ACT_LEAKY) return a > 0.f ? 1.f : (a < 0.f ? k.s_neg : k.s_zero);
  if (k.act == TN_ACT_LINEAR) return 1.f;
  if (k.act == TN_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  return act_bwd_slow(a, k.act);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10.  key = (seed lo, seed hi); ctr = (block, sample, step, purpose); element j of a
// sample is word j%4 of block j/4.
// ------------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ Philox4 philox_block(uint64_t seed, int purpose, uint32_t step,
                                                         uint32_t sample, uint32_t block) {
  return philox4x32_10(block, sample, step, (uint32_t)purpose, (uint32_t)(seed & 0xffffffffu),
                       (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint32_t philox_word(const Philox4 &p, int k) {
  return k == 0 ? p.x : (k == 1 ? p.y : (k == 2 ? p.z : p.w));
}

// uint32 threshold t with P(word < t) ~ p (oracle/philox.py bernoulli_threshold); p >= 1 is the
// caller's business.
inline uint32_t bernoulli_threshold(double p) {
  double t = floor(p * 4294967296.0);
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

__device__ __forceinline__ float ctl_lr(const int32_t *ctl) {
  return __int_as_float(ctl[TN_CTL_LR_BITS]);
}

}  // namespace tn
