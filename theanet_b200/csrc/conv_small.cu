// ConvLayer + 2x2 PoolLayer for small channel counts, second generation
// (theanet/layer/convpool.py:54-72 nnet.conv2d + bias + activation, :106-107 pool_2d, and their
// gradients through tt.grad, layer.py:83).
//
// The shipped networks (params/mnist.prms:13-27) run 3x3 'valid' convolutions over 1..20 maps of
// 13..28 pixel images followed by a 2x2 max-pool.  As GEMMs they are N=4,K=9 and N=20,K=36 --
// below any tensor-core tile and, with the 3xTF32 split float32 parity needs, no cheaper there than
// on the FMA pipe -- so they stay direct convolutions.  The first-generation kernels
// (conv_fused.cu) kept ONE image per CTA in shared memory: rows of 11..13 pixels left half of
// every warp idle, staging went element by element through index arithmetic and FFMA was 4% of the
// issued instructions.  Here
//
//   * a CTA stages a GROUP of images with one contiguous, vectorised copy (NCHW images of a
//     minibatch are adjacent in memory) and flattens its work items over the whole group, so
//     lanes stay busy whatever the image size is;
//   * forward: one thread = one 2x2 pool window x 4 output maps (16 accumulators, a 4x4 input
//     patch per channel in registers, filter taps as broadcast float4 loads): 144 FMAs per 25
//     shared-memory loads, and the pool is a register max -- no second pass;
//   * backward: ONE kernel per layer.  dL/dz = [a == pooled] * dL/dpooled * act'(pooled) (Theano's
//     tie-duplicating MaxPoolGrad) is rebuilt once per group, channel-last with a zero border, and
//     feeds both the weight gradient (thread = 4 maps x 1 channel x all 9 taps = 36 accumulators
//     living in registers for the whole kernel, two output rows per sliding-window pass) and the
//     input gradient (thread = 4 pixels x 4 channels, float4 loads over the map dimension);
//   * the cross-CTA sum of the weight-gradient partials is a two-level ticket (team of CTAs, then
//     teams) inside the same launch, in a fixed order: deterministic, no finishing kernel.
//
// Supported: filter 3x3, mode 'valid', pool 2 (ceil or ignore_border), ReLU-family / linear
// activations.  Everything else keeps using conv_fused.cu / conv_direct.cu.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "conv_small.cuh"

namespace tn {

constexpr int kST = 256;   // threads per CTA
constexpr int kSF = 3;     // filter size of this path
constexpr int kQB = 3;     // backward staging: quads of pooled cells in flight per thread

struct SmallArgs {
  const float *x, *W, *bias, *dtop, *below;
  float *a, *pooled, *dx, *dW, *db;
  uint8_t *tie;         // per pooled cell: bit (2*dy+dx) set where a[2pi+dy, 2pj+dx] == pooled
  float *partial, *teampart;
  unsigned *tickets;
  unsigned *gbar;       // grid barrier {arrivals, generation} of the cooperative final reduction
  int coop;             // 1: every CTA of the grid is resident at once (host checked the occupancy)
  long long *dbg;       // optional phase timestamps (tools/phase_times.py), 64 per CTA
  int B, C, S, M, O, P, Pc, NB;
  int Sp;               // forward: even row pitch of the staged images (8-byte patch loads)
  int G, CG;            // ceil(M/4), ceil(C/4)
  int Hp, ps;           // bordered dL/dz side (O + 2(f-1)), floats per pixel (maps, padded)
  int gx, gg;           // zeroed slack (floats) behind the staged images / the bordered dL/dz:
                        // windows of the last row overhang by up to a row / a few pixels
  int T, nsl;           // weight gradient: (map group, channel) combos, row slices
  int L, nseg, upi;     // ... segment length, segments per row pair, units per image
  int strips;           // input gradient: 4-pixel strips per row
  int npl;              // bias gradient: pixel lanes
  int nout, nout4, team;
  ActK ak, akb;
  FastDiv32 dPcPc, dPc, dNB, dPP, dP, dM, dUPI, dNSEG, dStrips, dS, dT, dHH, dH, dCFF, dNW;

};

// dst (16-byte aligned shared memory) <- n consecutive floats at src
__device__ __forceinline__ void stage_contig(float *dst, const float *__restrict__ src, int n) {
  const int tid = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = n >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int t = tid; t < n4; t += blockDim.x) d4[t] = __ldg(s4 + t);
    for (int t = 4 * n4 + tid; t < n; t += blockDim.x) dst[t] = __ldg(src + t);
  } else {
    for (int t = tid; t < n; t += blockDim.x) dst[t] = __ldg(src + t);
  }
}

// The same copy with cp.async (no registers, nothing waits until stage_wait()): used where the
// accumulators leave no room to keep many loads in flight
__device__ __forceinline__ void stage_contig_async(float *dst, const float *__restrict__ src, int n) {
  const int tid = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = n >> 2;
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    for (int t = tid; t < n4; t += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * t), "l"(src + 4 * t) : "memory");
    for (int t = 4 * n4 + tid; t < n; t += blockDim.x) dst[t] = __ldg(src + t);
  } else {
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    for (int t = tid; t < n; t += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * t), "l"(src + t) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// dst[row * pitch + col] <- src[row * S + col] (pitch > S: odd image sides get an even pitch)
__device__ __forceinline__ void stage_pitched(float *dst, const float *__restrict__ src, int n, int S,
                                              int pitch, const FastDiv32 &dS) {
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int row = (int)dS.div((uint32_t)t);
    dst[row * pitch + (t - row * S)] = __ldg(src + t);
  }
  // the pad columns are read by the windows of the last image column; what they feed is either
  // dropped or multiplied by a zero gradient -- and 0 * NaN is NaN, so they must hold zeros, not
  // whatever the previous kernel left in shared memory
  const int rows = n / S;
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
    for (int c = S; c < pitch; ++c) dst[r * pitch + c] = 0.f;
}

// ReLU family on this path, literally layer.py:36: max(0,z) + (min(0,z)*NN)/100 with BOTH roundings
// (product, then division), so that the forward activations are bit-identical to the NumPy
// restatement.  A single-rounding z*(NN/100) is <= 1 ulp away, but a slope below 1 maps
// neighbouring floats onto the same output, WHICH neighbours collide depends on that last ulp, and
// the tie-duplicating MaxPoolGrad routes the gradient by exactly those collisions.  relu (slope 0)
// and linear (slope 1) are exact in one fused multiply-add.
__device__ __forceinline__ float act_small(const ActK &k, float z) {
  if (k.act == TN_ACT_LEAKY && k.nn != 0.f) return z > 0.f ? z : div100_rn(__fmul_rn(z, k.nn));
  return fmaf(fminf(z, 0.f), k.s_neg, fmaxf(z, 0.f));
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// smem: ws[(c*F+u)*F+v][4G] (taps flipped: true convolution) | bs[4G] | xs[NB][C][S][Sp] + guard
template <int F>
__global__ void __launch_bounds__(kST) small_fprop_kernel(const SmallArgs k) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int C = k.C, S = k.S, M = k.M, O = k.O, P = k.P, Pc = k.Pc, G = k.G, NB = k.NB, Sp = k.Sp;
  const int coP = 4 * G, SSp = S * Sp, CSSp = C * SSp, CSS = C * S * S, OO = O * O, PP = P * P;
  const int PcPc = Pc * Pc;
  float *ws = sm;
  float *bs = ws + C * F * F * coP;
  float *xs = bs + coP;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < C * F * F * coP; t += nt) {
    const int co = t % coP;
    int r = t / coP;
    const int v = r % F; r /= F;
    const int u = r % F;
    const int c = r / F;
    ws[t] = co < M ? k.W[((co * C + c) * F + (F - 1 - u)) * F + (F - 1 - v)] : 0.f;
  }
  for (int t = tid; t < coP; t += nt) bs[t] = t < M ? k.bias[t] : 0.f;
  for (int t = tid; t < k.gx; t += nt) xs[NB * CSSp + t] = 0.f;
  const int items = G * NB * PcPc;   // (g, b, pi, pj), pj fastest: a warp shares its taps
  // relu / linear as members of the leaky family: slope 0 / 1 on the negative side
  ActK ak = k.ak;
  if (ak.act == TN_ACT_RELU) ak.s_neg = 0.f;
  if (ak.act == TN_ACT_LINEAR) ak.s_neg = 1.f;

  for (int grp = blockIdx.x; grp * NB < k.B; grp += gridDim.x) {
    const int b0 = grp * NB, nb = min(NB, k.B - b0);
    if (Sp == S) stage_contig(xs, k.x + (size_t)b0 * CSS, nb * CSS);
    else stage_pitched(xs, k.x + (size_t)b0 * CSS, nb * CSS, S, Sp, k.dS);
    __syncthreads();
    for (int it = tid; it < items; it += nt) {
      const int q1 = (int)k.dPcPc.div(it);
      const int cell = it - q1 * PcPc;
      const int pi = (int)k.dPc.div(cell), pj = cell - pi * Pc;
      const int g = (int)k.dNB.div(q1), b = q1 - g * NB;
      if (b >= nb) continue;
      // conv outputs (2pi+dy, 2pj+dx), dy, dx in {0,1}, of maps 4g..4g+3
      float acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
      // the pitch is even and so is 2*pj: every patch row is two aligned 8-byte loads
      const float2 *xb = reinterpret_cast<const float2 *>(xs + b * CSSp + (2 * pi) * Sp + 2 * pj);
      const float4 *w4 = reinterpret_cast<const float4 *>(ws) + g;
      const int hp = Sp >> 1, hss = SSp >> 1;
      for (int c = 0; c < C; ++c) {
        float p[F + 1][F + 1];   // rows/cols past the image only feed outputs that are dropped
#pragma unroll
        for (int r = 0; r <= F; ++r)
#pragma unroll
          for (int e = 0; e <= F; e += 2) {
            const float2 t2 = xb[r * hp + (e >> 1)];
            p[r][e] = t2.x;
            if (e + 1 <= F) p[r][e + 1] = t2.y;
          }
#pragma unroll
        for (int u = 0; u < F; ++u)
#pragma unroll
          for (int v = 0; v < F; ++v) {
            const float4 w = w4[((c * F + u) * F + v) * G];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) {
                const float xv = p[dy + u][dx + v];
                acc[dy * 2 + dx][0] = fmaf(xv, w.x, acc[dy * 2 + dx][0]);
                acc[dy * 2 + dx][1] = fmaf(xv, w.y, acc[dy * 2 + dx][1]);
                acc[dy * 2 + dx][2] = fmaf(xv, w.z, acc[dy * 2 + dx][2]);
                acc[dy * 2 + dx][3] = fmaf(xv, w.w, acc[dy * 2 + dx][3]);
              }
          }
        xb += hss;
      }
      const int i0 = 2 * pi, j0 = 2 * pj;
      const bool r1 = i0 + 1 < O, c1 = j0 + 1 < O;
      const bool pool_ok = pi < P && pj < P;
      const float4 b4 = reinterpret_cast<const float4 *>(bs)[g];
      const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
      const int img = (b0 + b) * M + 4 * g;                 // (image, map) plane of q = 0
      float *ar = k.a ? k.a + (size_t)img * OO + i0 * O + j0 : nullptr;
      const int po = img * PP + pi * P + pj;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (4 * g + q >= M) break;
        const float v00 = act_small(ak, acc[0][q] + bq[q]);
        const float v01 = act_small(ak, acc[1][q] + bq[q]);
        const float v10 = act_small(ak, acc[2][q] + bq[q]);
        const float v11 = act_small(ak, acc[3][q] + bq[q]);
        float mx = v00;
        if (c1) mx = fmaxf(mx, v01);
        if (r1) mx = fmaxf(mx, v10);
        if (r1 && c1) mx = fmaxf(mx, v11);
        if (ar) {
          float *aq = ar + q * OO;
          aq[0] = v00;
          if (c1) aq[1] = v01;
          if (r1) aq[O] = v10;
          if (r1 && c1) aq[O + 1] = v11;
        }
        if (pool_ok) {
          k.pooled[po + q * PP] = mx;
          if (k.tie)
            k.tie[po + q * PP] = (uint8_t)((v00 == mx ? 1 : 0) | ((c1 && v01 == mx) ? 2 : 0) |
                                           ((r1 && v10 == mx) ? 4 : 0) | ((r1 && c1 && v11 == mx) ? 8 : 0));
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward: dW, db (and dx when the layer below needs it) in one launch
// ---------------------------------------------------------------------------------------------
// smem: gz[NB][Hp][Hp][ps] + guard | xs[NB][C][S][S] + guard | wd[(m*F+u)*F+v][4CG]
#define TN_PHASE(i)                                                                       \
  do {                                                                                    \
    if (k.dbg && threadIdx.x == 0) k.dbg[(size_t)blockIdx.x * 64 + (i)] = clock64();      \
  } while (0)

// Shared tail of the backward kernels.  On entry red[slice][o] (o = e*T + combo, nW words per
// slice) holds every slice's weight-gradient accumulators and dbs[lane][4G] the bias-gradient
// lanes, all visible (barrier passed).  Sums them in fixed-order radix-4 trees, writes the CTA's
// partial and runs the two-level ticket; the last CTA scatters dW (taps flipped back) and db.
template <int F>
__device__ __noinline__ void small_finish(const SmallArgs &k, float *red, float *dbs, int *flag) {
  constexpr int FF = F * F;
  const int tid = threadIdx.x;
  const int C = k.C, M = k.M, T = k.T, mP = 4 * k.G;
  const int nW = T * 4 * FF, nWp = nW | 1;    // odd slice stride: conflict-free slice writes
  int &s_flag = *flag;
  TN_PHASE(48);
  for (int step = 1; step < k.nsl; step *= 4) {
    const int nown = (k.nsl + 4 * step - 1) / (4 * step);
#pragma unroll 4
    for (int w = tid; w < nown * nW; w += kST) {
      const int ow = (int)k.dNW.div((uint32_t)w);
      const int o = w - ow * nW, sl = ow * 4 * step;
      float *v = red + sl * nWp + o;
      float s = v[0];
      if (sl + step < k.nsl) s += v[step * nWp];
      if (sl + 2 * step < k.nsl) s += v[2 * step * nWp];
      if (sl + 3 * step < k.nsl) s += v[3 * step * nWp];
      v[0] = s;
    }
    __syncthreads();
  }
  TN_PHASE(49);
  float *pout = k.partial + (size_t)blockIdx.x * k.nout4;
  for (int o = tid; o < nW; o += kST) {
    const int e = (int)k.dT.div((uint32_t)o), cb = o - e * T;
    pout[cb * 4 * FF + e] = red[o];
  }
  TN_PHASE(50);
  // bias gradient: the same tree over the pixel lanes, dbs[lane * G + group][4] -> [lane][4G]
  for (int step = 1; step < k.npl; step *= 4) {
    const int nown = (k.npl + 4 * step - 1) / (4 * step);
#pragma unroll 2
    for (int w = tid; w < nown * mP; w += kST) {
      const int ow = w / mP, m = w - ow * mP, p0 = ow * 4 * step;
      float *v = dbs + p0 * mP + m;
      float s = v[0];
      if (p0 + step < k.npl) s += v[step * mP];
      if (p0 + 2 * step < k.npl) s += v[2 * step * mP];
      if (p0 + 3 * step < k.npl) s += v[3 * step * mP];
      v[0] = s;
    }
    __syncthreads();
  }
  if (tid < mP) pout[nW + tid] = dbs[tid];
  for (int t = nW + mP + tid; t < k.nout4; t += kST) pout[t] = 0.f;

  TN_PHASE(56);
  if (k.coop) {
    // ---- all CTAs are resident: one grid barrier, then EVERY CTA sums a few outputs over all the
    // partials (thread t takes partials t, t+256, ... in order, then a fixed-shape tree:
    // deterministic) -- instead of two ticket levels that end in one CTA summing everything.
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      unsigned gen;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(k.gbar + 1) : "memory");
      if (atomicAdd(k.gbar, 1u) == gridDim.x - 1) {
        k.gbar[0] = 0u;                                   // ready for the next launch
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(k.gbar + 1), "r"(gen + 1u) : "memory");
      } else {
        unsigned now;
        const long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(k.gbar + 1) : "memory");
          if (clock64() - t0 > 4000000000ll) __trap();    // not co-resident after all: fail loudly
        } while (now == gen);
      }
    }
    __syncthreads();
    TN_PHASE(57);
    const int nout = nW + M;
    const int per = (nout + (int)gridDim.x - 1) / (int)gridDim.x;
    const int o0 = blockIdx.x * per, o1 = min(o0 + per, nout);
    float *wred = dbs;                                    // 8 warps x outputs, reused
    for (int o = o0; o < o1; ++o) {
      const int src = o < nW ? o : nW + (o - nW);         // [nW, nW + M): bias gradient
      float sv = 0.f;
      for (int pc = tid; pc < (int)gridDim.x; pc += kST) sv += __ldcg(k.partial + (size_t)pc * k.nout4 + src);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, sh);
      if ((tid & 31) == 0) wred[tid >> 5] = sv;
      __syncthreads();
      if (tid == 0) {
        float tot = wred[0];
#pragma unroll
        for (int w = 1; w < kST / 32; ++w) tot += wred[w];
        if (o < nW) {   // o = ((mg*C + c)*4 + q)*FF + u*F + v   (correlation taps: flip back)
          const int e = o % FF;
          int r = o / FF;
          const int q = r & 3; r >>= 2;
          const int cc = r % C, g = r / C;
          const int m = 4 * g + q, u = e / F, v = e - u * F;
          if (m < M) k.dW[((m * C + cc) * F + (F - 1 - u)) * F + (F - 1 - v)] = tot;
        } else {
          k.db[o - nW] = tot;
        }
      }
      __syncthreads();
    }
    TN_PHASE(58);
    return;
  }
  // ---- two-level ticket: last CTA of a team sums the team, last team sums the teams --------
  const int team = blockIdx.x / k.team;
  const int nteam = (gridDim.x + k.team - 1) / k.team;
  const int tsize = min(k.team, (int)gridDim.x - team * k.team);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = atomicAdd(&k.tickets[team], 1u) == (unsigned)(tsize - 1);
  __syncthreads();
  TN_PHASE(57);
  if (!s_flag) return;
  __threadfence();
  const int n4 = k.nout4 >> 2;
  {
    const float4 *src = reinterpret_cast<const float4 *>(k.partial) + (size_t)team * k.team * n4;
    float4 *dst = reinterpret_cast<float4 *>(k.teampart) + (size_t)team * n4;
    for (int t = tid; t < n4; t += kST) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
      for (int j = 0; j < tsize; ++j) {
        const float4 v = __ldcg(src + (size_t)j * n4 + t);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      dst[t] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = atomicAdd(&k.tickets[nteam], 1u) == (unsigned)(nteam - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  {
    const float4 *src = reinterpret_cast<const float4 *>(k.teampart);
    for (int t = tid; t < n4; t += kST) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
      for (int j = 0; j < nteam; ++j) {
        const float4 v = __ldcg(src + (size_t)j * n4 + t);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const int o = 4 * t + e4;
        if (o < nW) {   // o = ((mg*C + c)*4 + q)*FF + u*F + v   (correlation taps: flip back)
          const int e = o % FF;
          int r = o / FF;
          const int q = r & 3; r >>= 2;
          const int cc = r % C, g = r / C;
          const int m = 4 * g + q, u = e / F, v = e - u * F;
          if (m < M) k.dW[((m * C + cc) * F + (F - 1 - u)) * F + (F - 1 - v)] = sv[e4];
        } else if (o < nW + M) {
          k.db[o - nW] = sv[e4];
        }
      }
    }
  }
  for (int t = tid; t <= nteam; t += kST) k.tickets[t] = 0u;   // ready for the next launch
  TN_PHASE(58);
}

template <int F>
__global__ void __launch_bounds__(kST, 2) small_bwd_kernel(const __grid_constant__ SmallArgs k) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_flag;
  TN_PHASE(0);
  constexpr int FF = F * F;
  constexpr int pd = F - 1;
  const int C = k.C, S = k.S, M = k.M, O = k.O, P = k.P, G = k.G, CG = k.CG, NB = k.NB;
  const int Hp = k.Hp, ps = k.ps, cP = 4 * CG, mP = 4 * G;
  const int SS = S * S, CSS = C * SS, OO = O * O, PP = P * P;
  const int gzimg = Hp * Hp * ps;
  float *dbs = sm;                       // [kST][4] bias-gradient partial sums, outside the aliased area
  float *gz = dbs + 4 * kST;
  float *xs = gz + NB * gzimg + k.gg;
  float *wd = xs + ((NB * CSS + k.gx + 3) & ~3);
  const int tid = threadIdx.x;
  const bool need_dx = k.dx != nullptr;
  reinterpret_cast<float4 *>(dbs)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);

  // Zero what the staging below never writes -- the border frame, map pixels past the pooled
  // windows, the guards -- and nothing else: no barrier is needed between this and the staging,
  // and the filter taps / first image group are already in flight meanwhile.
  {
    const int lim = k.tie ? 2 * P : min(2 * P, O);     // the tie path also writes window overhang
    const int HH = Hp * Hp, ps4 = ps >> 2;
    float4 *gz4 = reinterpret_cast<float4 *>(gz);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = tid; p < NB * HH; p += kST) {
      const int pin = p - (int)k.dHH.div((uint32_t)p) * HH;
      const int y = (int)k.dH.div((uint32_t)pin);
      const int yy = y - pd, xx = pin - y * Hp - pd;
      if (yy < 0 || xx < 0 || yy >= lim || xx >= lim) {
        for (int i = 0; i < ps4; ++i) gz4[p * ps4 + i] = z4;
      } else if (M < ps) {
        // interior pixel: the staging writes maps 0..M-1 only; the padded map slots are read by the
        // input-gradient loop against zero weights -- 0 * NaN is NaN, so they hold zeros
        for (int m = M; m < ps; ++m) gz[p * ps + m] = 0.f;
      }
    }
    for (int t = tid; t < k.gg; t += kST) gz[NB * gzimg + t] = 0.f;
    for (int t = tid; t < k.gx; t += kST) xs[NB * CSS + t] = 0.f;
  }
  if (need_dx) {   // wd[(m*F+u)*F+v][c] = W[m][c][u][v]: coalesced reads, scattered stores
    const int CFF = C * FF, nwt = M * CFF;
    for (int t0 = tid; t0 < nwt; t0 += 4 * kST) {
      float wv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (t0 + r * kST < nwt) wv[r] = __ldg(k.W + t0 + r * kST);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int t = t0 + r * kST;
        if (t < nwt) {
          const int m = (int)k.dCFF.div((uint32_t)t);
          const int rr = t - m * CFF;
          const int co = rr / FF, e = rr - co * FF;
          wd[(m * FF + e) * cP + co] = wv[r];
        }
      }
    }
    if (cP != C || mP != M) {   // padded channels / maps read as zero (disjoint from the above)
      for (int t = tid; t < mP * FF * cP; t += kST) {
        const int co = t % cP;
        if (co >= C || t / (cP * FF) >= M) wd[t] = 0.f;
      }
    }
  }
  // weight gradient: thread = (row slice, map group, channel)
  const int T = k.T;
  const bool wact = tid < T * k.nsl;
  const int slice = tid / T, combo = tid - slice * T;
  const int mg = combo / C, c = combo - mg * C;
  float acc[4][FF];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < FF; ++e) acc[q][e] = 0.f;
  // bias gradient: thread = (pixel lane, map group)
  const bool dbact = tid < G * k.npl;
  const int pl = tid / G, mgd = tid - pl * G;
  const int grow = Hp * ps;
  TN_PHASE(1);

  int stage_no = 0;
  for (int grp = blockIdx.x; grp * NB < k.B; grp += gridDim.x, ++stage_no) {
    const int b0 = grp * NB, nb = min(NB, k.B - b0);
    const int ph0 = min(stage_no, 5) * 8;
    stage_contig_async(xs, k.x + (size_t)b0 * CSS, nb * CSS);
    if (nb < NB) {
      // ragged last group: the windows of its last image overhang into the NEXT image slot, which
      // holds whatever an earlier group (or an earlier kernel) left there.  Those products are
      // multiplied by the zero border of dL/dz -- but 0 * NaN is NaN: give them zeros to read.
      for (int t = tid; t < k.gx; t += kST) xs[nb * CSS + t] = 0.f;
      for (int t = tid; t < k.gg; t += kST) gz[nb * gzimg + t] = 0.f;
    }
    {  // dL/dz of the conv layer, one thread per pooled cell: g = dL/dpooled * act'(pooled) goes to
       // every element of the window that equals the maximum (Theano's MaxPoolGrad)
      const int cell0 = b0 * M * PP;
      const int ncell = nb * M * PP;
      if (k.tie) {
        // tie pattern recorded by the forward kernel.  Window elements past the map edge have their
        // bit clear and land in the zero border.  Four consecutive cells per thread and all loads
        // of a batch issued before the first use: the phase costs ~one memory round trip.
        const bool vec = ((cell0 | ncell) & 3) == 0;
        const int nq = vec ? ncell >> 2 : 0;
        const float4 *pp4 = reinterpret_cast<const float4 *>(k.pooled + cell0);
        const float4 *dp4 = reinterpret_cast<const float4 *>(k.dtop + cell0);
        const uchar4 *tp4 = reinterpret_cast<const uchar4 *>(k.tie + cell0);
        for (int q0 = tid; q0 < nq; q0 += kQB * kST) {
          float4 po4[kQB], d4[kQB];
          uchar4 mk4[kQB];
#pragma unroll
          for (int r = 0; r < kQB; ++r) {
            const int q = q0 + r * kST;
            if (q < nq) {
              po4[r] = __ldg(pp4 + q);
              d4[r] = __ldg(dp4 + q);
              mk4[r] = __ldg(tp4 + q);
            }
          }
#pragma unroll
          for (int r = 0; r < kQB; ++r) {
            const int q = q0 + r * kST;
            if (q >= nq) break;
            const int t = 4 * q;
            const int bm = (int)k.dPP.div(t);
            const int p = t - bm * PP;
            int pi = (int)k.dP.div(p), pj = p - pi * P;
            int b = (int)k.dM.div(bm), m = bm - b * M;
            const float pov[4] = {po4[r].x, po4[r].y, po4[r].z, po4[r].w};
            const float dv[4] = {d4[r].x, d4[r].y, d4[r].z, d4[r].w};
            const unsigned mkv[4] = {mk4[r].x, mk4[r].y, mk4[r].z, mk4[r].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float gg = dv[i] * act_bwd_t<false>(k.ak, pov[i]);
              float *gr = gz + b * gzimg + ((2 * pi + pd) * Hp + 2 * pj + pd) * ps + m;
              gr[0] = (mkv[i] & 1u) ? gg : 0.f;
              gr[ps] = (mkv[i] & 2u) ? gg : 0.f;
              gr[grow] = (mkv[i] & 4u) ? gg : 0.f;
              gr[grow + ps] = (mkv[i] & 8u) ? gg : 0.f;
              if (++pj == P) {          // next cell of the (image, map, row, column) order
                pj = 0;
                if (++pi == P) {
                  pi = 0;
                  if (++m == M) { m = 0; ++b; }
                }
              }
            }
          }
        }
        for (int t = 4 * nq + tid; t < ncell; t += kST) {
          const float po = __ldg(k.pooled + cell0 + t), d = __ldg(k.dtop + cell0 + t);
          const unsigned mk = __ldg(k.tie + cell0 + t);
          const int bm = (int)k.dPP.div(t);
          const int p = t - bm * PP;
          const int pi = (int)k.dP.div(p), pj = p - pi * P;
          const int b = (int)k.dM.div(bm), m = bm - b * M;
          const float gg = d * act_bwd_t<false>(k.ak, po);
          float *gr = gz + b * gzimg + ((2 * pi + pd) * Hp + 2 * pj + pd) * ps + m;
          gr[0] = (mk & 1u) ? gg : 0.f;
          gr[ps] = (mk & 2u) ? gg : 0.f;
          gr[grow] = (mk & 4u) ? gg : 0.f;
          gr[grow + ps] = (mk & 8u) ? gg : 0.f;
        }
      } else {
        const float *ap = k.a + (size_t)b0 * M * OO;
        for (int t = tid; t < ncell; t += kST) {
          const float po = __ldg(k.pooled + cell0 + t), d = __ldg(k.dtop + cell0 + t);
          const int bm = (int)k.dPP.div(t);
          const int p = t - bm * PP;
          const int pi = (int)k.dP.div(p), pj = p - pi * P;
          const int b = (int)k.dM.div(bm), m = bm - b * M;
          const float gg = d * act_bwd_t<false>(k.ak, po);
          const float *ar = ap + bm * OO + (2 * pi) * O + 2 * pj;
          float *gr = gz + b * gzimg + ((2 * pi + pd) * Hp + 2 * pj + pd) * ps + m;
          const bool r1 = 2 * pi + 1 < O, c1 = 2 * pj + 1 < O;
          gr[0] = __ldg(ar) == po ? gg : 0.f;
          if (c1) gr[ps] = __ldg(ar + 1) == po ? gg : 0.f;
          if (r1) {
            gr[grow] = __ldg(ar + O) == po ? gg : 0.f;
            if (c1) gr[grow + ps] = __ldg(ar + O + 1) == po ? gg : 0.f;
          }
        }
      }
    }
    stage_wait();
    __syncthreads();
    TN_PHASE(ph0 + 2);

    if (wact) {
      const int nun = nb * k.upi;
      for (int un = slice; un < nun; un += k.nsl) {
        const int b = (int)k.dUPI.div(un);
        const int r = un - b * k.upi;
        const int rp = (int)k.dNSEG.div(r), seg = r - rp * k.nseg;
        const int i0 = 2 * rp, j0 = seg * k.L;
        const int len = min(k.L, O - j0);
        // output rows i0, i0+1 (the second may be the zero border) against input rows i0..i0+F
        const float *xr = xs + (b * C + c) * SS + i0 * S + j0;
        const float *g0 = gz + b * gzimg + ((i0 + pd) * Hp + pd + j0) * ps + 4 * mg;
        // sliding window over the columns, kept in a register ring: column j0+col lives in slot
        // col % F, the j loop is unrolled F-fold so every slot index is a compile-time constant
        float xw[F + 1][F];
#pragma unroll
        for (int r2 = 0; r2 <= F; ++r2)
#pragma unroll
          for (int v = 0; v < F - 1; ++v) xw[r2][v] = xr[r2 * S + v];
        for (int j = 0; j < len; j += F) {
#pragma unroll
          for (int jj = 0; jj < F; ++jj) {
            if (j + jj < len) {
#pragma unroll
              for (int r2 = 0; r2 <= F; ++r2) xw[r2][(jj + F - 1) % F] = xr[r2 * S + j + jj + F - 1];
              const float4 ga = *reinterpret_cast<const float4 *>(g0 + (j + jj) * ps);
              const float4 gb = *reinterpret_cast<const float4 *>(g0 + grow + (j + jj) * ps);
#pragma unroll
              for (int u = 0; u < F; ++u)
#pragma unroll
                for (int v = 0; v < F; ++v) {
                  const float x0 = xw[u][(jj + v) % F], x1 = xw[u + 1][(jj + v) % F];
                  acc[0][u * F + v] = fmaf(gb.x, x1, fmaf(ga.x, x0, acc[0][u * F + v]));
                  acc[1][u * F + v] = fmaf(gb.y, x1, fmaf(ga.y, x0, acc[1][u * F + v]));
                  acc[2][u * F + v] = fmaf(gb.z, x1, fmaf(ga.z, x0, acc[2][u * F + v]));
                  acc[3][u * F + v] = fmaf(gb.w, x1, fmaf(ga.w, x0, acc[3][u * F + v]));
                }
            }
          }
        }
      }
    }
    TN_PHASE(ph0 + 3);
    if (dbact) {   // the border pixels are zero: one linear pass over the whole bordered maps
      // the running sums live in shared memory between stages: registers are scarce in dgrad
      float4 dba = reinterpret_cast<float4 *>(dbs)[tid];
      const int npx = nb * Hp * Hp;
#pragma unroll 4
      for (int px = pl; px < npx; px += k.npl) {
        const float4 g = *reinterpret_cast<const float4 *>(gz + px * ps + 4 * mgd);
        dba.x += g.x; dba.y += g.y; dba.z += g.z; dba.w += g.w;
      }
      reinterpret_cast<float4 *>(dbs)[tid] = dba;
    }

    TN_PHASE(ph0 + 4);
    if (need_dx) {
      // dx[c,y,x] = sum_{m,u,v} gzb[m,y+u,x+v] W[m,c,u,v]; item = (channel group, image, row, strip)
      // rows fastest: neighbouring lanes read float4s one bordered row (Hp*ps floats) apart, which
      // spreads over the banks; neighbouring strips (4*ps floats apart) would collide
      const int nit = CG * NB * S * k.strips;
      for (int it = tid; it < nit; it += kST) {
        const int r = (int)k.dS.div(it);
        const int y = it - r * S;
        const int r2 = (int)k.dStrips.div(r);
        const int s = r - r2 * k.strips;
        const int cg = (int)k.dNB.div(r2), b = r2 - cg * NB;
        if (b >= nb) continue;
        const int x0 = 4 * s;
        float ac[4][4];   // [pixel][channel]
#pragma unroll
        for (int l = 0; l < 4; ++l)
#pragma unroll
          for (int q = 0; q < 4; ++q) ac[l][q] = 0.f;
        const float *gb0 = gz + b * gzimg + (y * Hp + x0) * ps;
        const float4 *wq = reinterpret_cast<const float4 *>(wd) + cg;
        for (int m4 = 0; m4 < G; ++m4) {
#pragma unroll
          for (int u = 0; u < F; ++u) {
            float4 gv[4 + F - 1];
#pragma unroll
            for (int e = 0; e < 4 + F - 1; ++e)
              gv[e] = *reinterpret_cast<const float4 *>(gb0 + (u * Hp + e) * ps + 4 * m4);
#pragma unroll
            for (int v = 0; v < F; ++v) {
              const float4 *wp = wq + (((4 * m4) * F + u) * F + v) * CG;
              const float4 w0 = wp[0], w1 = wp[FF * CG], w2 = wp[2 * FF * CG], w3 = wp[3 * FF * CG];
#pragma unroll
              for (int l = 0; l < 4; ++l) {
                const float4 g = gv[l + v];
                ac[l][0] = fmaf(g.w, w3.x, fmaf(g.z, w2.x, fmaf(g.y, w1.x, fmaf(g.x, w0.x, ac[l][0]))));
                ac[l][1] = fmaf(g.w, w3.y, fmaf(g.z, w2.y, fmaf(g.y, w1.y, fmaf(g.x, w0.y, ac[l][1]))));
                ac[l][2] = fmaf(g.w, w3.z, fmaf(g.z, w2.z, fmaf(g.y, w1.z, fmaf(g.x, w0.z, ac[l][2]))));
                ac[l][3] = fmaf(g.w, w3.w, fmaf(g.z, w2.w, fmaf(g.y, w1.w, fmaf(g.x, w0.w, ac[l][3]))));
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = 4 * cg + q;
          if (ch >= C) break;
          const int o = ((b0 + b) * C + ch) * SS + y * S + x0;
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            if (x0 + l < S) {
              float v = ac[l][q];
              if (k.below) v *= act_bwd_t<false>(k.akb, __ldg(k.below + o + l));
              k.dx[o + l] = v;
            }
          }
        }
      }
    }
    TN_PHASE(ph0 + 5);
    __syncthreads();
    TN_PHASE(ph0 + 6);
  }

  // ---- per-CTA partial: row slices combined in a fixed order -------------------------------
  // V[slice][o], o = e * T + combo (e = accumulator index q*FF+tap): lanes write consecutive words.
  // Radix-4 tree over the slices, in place: at each level the slice at a multiple of 4*step adds
  // its three neighbours step apart, ((a + b) + c) + d -- short dependent chains, a fixed order.
  float *red = gz;   // aliases the staging buffers (the loop ended on a barrier); dbs stays intact
  const int nW = T * 4 * FF;
  if (wact) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < FF; ++e) red[slice * (nW | 1) + (q * FF + e) * T + combo] = acc[q][e];
  }
  __syncthreads();
  small_finish<F>(k, red, dbs, &s_flag);
}

// ---------------------------------------------------------------------------------------------
// backward without dx (first weighted layer): dW, db straight from the pooled cells
// ---------------------------------------------------------------------------------------------
// No dL/dz image is built at all: a thread owns one (map group, channel) pair for the whole kernel
// (36 accumulators) and walks pooled cells; per cell it forms the window's four dL/dz values of its
// four maps in registers (pooled, dL/dpooled, tie bits) and multiplies them with the 4x4 input
// patch under the window: 144 FMAs per cell, the mirror image of the forward kernel.
// smem: dbs[kST][4] | xs[NB][C][S][Sp] + guard (aliased by the slice sums at the end)
template <int F>
__global__ void __launch_bounds__(kST, 2) small_wgrad_kernel(const __grid_constant__ SmallArgs k) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_flag;
  TN_PHASE(0);
  constexpr int FF = F * F;
  const int C = k.C, S = k.S, M = k.M, O = k.O, P = k.P, G = k.G, NB = k.NB, Sp = k.Sp;
  const int SSp = S * Sp, CSSp = C * SSp, CSS = C * S * S, OO = O * O, PP = P * P, mP = 4 * G;
  float *dbs = sm;
  float *xs = dbs + 4 * kST;
  const int tid = threadIdx.x;
  reinterpret_cast<float4 *>(dbs)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = tid; t < k.gx; t += kST) xs[NB * CSSp + t] = 0.f;
  const int T = k.T;
  const bool wact = tid < T * k.nsl;
  const int slice = tid / T, combo = tid - slice * T;
  const int mg = combo / C, c = combo - mg * C;
  float acc[4][FF];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < FF; ++e) acc[q][e] = 0.f;
  float dbq[4] = {0.f, 0.f, 0.f, 0.f};
  const int hp = Sp >> 1;
  TN_PHASE(1);

  int stage_no = 0;
  for (int grp = blockIdx.x; grp * NB < k.B; grp += gridDim.x, ++stage_no) {
    const int b0 = grp * NB, nb = min(NB, k.B - b0);
    const int ph0 = min(stage_no, 5) * 8;
    if (Sp == S) stage_contig(xs, k.x + (size_t)b0 * CSS, nb * CSS);
    else stage_pitched(xs, k.x + (size_t)b0 * CSS, nb * CSS, S, Sp, k.dS);
    if (nb < NB)       // ragged last group: zeros, not leftovers, behind its last image (0 * NaN)
      for (int t = tid; t < k.gx; t += kST) xs[nb * CSSp + t] = 0.f;
    __syncthreads();
    TN_PHASE(ph0 + 2);
    if (wact) {
      const int ncell = nb * PP;
      // the three global values per map of the NEXT cell are fetched while this one is multiplied
      float npo[4], nd[4];
      unsigned nmk[4];
      auto fetch = [&](int un) {
        const int b = (int)k.dPP.div((uint32_t)un);
        const int i0 = ((b0 + b) * M + 4 * mg) * PP + (un - b * PP);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ok = 4 * mg + q < M;
          npo[q] = ok ? __ldg(k.pooled + i0 + q * PP) : 0.f;
          nd[q] = ok ? __ldg(k.dtop + i0 + q * PP) : 0.f;
          nmk[q] = (ok && k.tie) ? __ldg(k.tie + i0 + q * PP) : 0u;
        }
      };
      if (slice < ncell) fetch(slice);
      for (int un = slice; un < ncell; un += k.nsl) {
        const int b = (int)k.dPP.div((uint32_t)un);
        const int cell = un - b * PP;
        const int pi = (int)k.dP.div((uint32_t)cell), pj = cell - pi * P;
        // dL/dz of the 2x2 window, maps 4mg..4mg+3: g[q][2*dy+dx]
        float g[4][4];
        const bool r1 = 2 * pi + 1 < O, c1 = 2 * pj + 1 < O;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float po = npo[q];
          const float gg = nd[q] * act_bwd_t<false>(k.ak, po);      // 0 for maps past M
          if (k.tie) {
            const unsigned mk = nmk[q];
            g[q][0] = (mk & 1u) ? gg : 0.f;
            g[q][1] = (mk & 2u) ? gg : 0.f;
            g[q][2] = (mk & 4u) ? gg : 0.f;
            g[q][3] = (mk & 8u) ? gg : 0.f;
          } else if (4 * mg + q < M) {
            const float *ar = k.a + ((size_t)(b0 + b) * M + 4 * mg + q) * OO + (2 * pi) * O + 2 * pj;
            g[q][0] = __ldg(ar) == po ? gg : 0.f;
            g[q][1] = (c1 && __ldg(ar + 1) == po) ? gg : 0.f;
            g[q][2] = (r1 && __ldg(ar + O) == po) ? gg : 0.f;
            g[q][3] = (r1 && c1 && __ldg(ar + O + 1) == po) ? gg : 0.f;
          } else {
            g[q][0] = g[q][1] = g[q][2] = g[q][3] = 0.f;
          }
        }
        if (un + k.nsl < ncell) fetch(un + k.nsl);
        if (c == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) dbq[q] += (g[q][0] + g[q][1]) + (g[q][2] + g[q][3]);
        }
        // input patch under the window (rows / columns past the image meet zero g only)
        const float2 *xb = reinterpret_cast<const float2 *>(xs + b * CSSp + c * SSp + (2 * pi) * Sp + 2 * pj);
        float p[F + 1][F + 1];
#pragma unroll
        for (int r = 0; r <= F; ++r)
#pragma unroll
          for (int e = 0; e <= F; e += 2) {
            const float2 t2 = xb[r * hp + (e >> 1)];
            p[r][e] = t2.x;
            if (e + 1 <= F) p[r][e + 1] = t2.y;
          }
#pragma unroll
        for (int u = 0; u < F; ++u)
#pragma unroll
          for (int v = 0; v < F; ++v)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q][u * F + v] = fmaf(g[q][3], p[u + 1][v + 1],
                                  fmaf(g[q][2], p[u + 1][v],
                                  fmaf(g[q][1], p[u][v + 1], fmaf(g[q][0], p[u][v], acc[q][u * F + v]))));
      }
    }
    TN_PHASE(ph0 + 5);
    __syncthreads();
    TN_PHASE(ph0 + 6);
  }
  float *red = xs;
  const int nW = T * 4 * FF;
  if (wact) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < FF; ++e) red[slice * (nW | 1) + (q * FF + e) * T + combo] = acc[q][e];
    if (c == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) dbs[slice * mP + 4 * mg + q] = dbq[q];
    }
  }
  __syncthreads();
  small_finish<F>(k, red, dbs, &s_flag);
}

// ---------------------------------------------------------------------------------------------
// host side: geometry, work split, launch
// ---------------------------------------------------------------------------------------------
static long long *g_small_dbg = nullptr;   // tn_convpool_debug_timestamps

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int guard_x(int S) { return (S + 8 + 3) / 4 * 4; }

bool small_conv_ok(int C, int S, int M, int f, int pad_lo, int O, int act, int pool, int P) {
  if (env_int("TN_CONV_SMALL", 1) == 0) return false;
  if (f != kSF || pad_lo != 0 || pool != 2 || O != S - f + 1 || O < 2) return false;
  if (P != (O + 1) / 2 && P != O / 2) return false;
  if (!act_is_fast(act)) return false;
  const int G = (M + 3) / 4;
  if (G * C > kST || M > 64 || C > 64) return false;
  // one image must fit: bordered channel-last dL/dz + the image + dgrad taps
  const int Hp = O + 2 * (f - 1);
  int ps = 4 * G;
  if (((ps / 4) & 1) == 0) ps += 4;
  const size_t one = ((size_t)4 * kST + (size_t)Hp * Hp * ps + 4 * ps + (size_t)C * S * S + guard_x(S) + 4 +
                      (size_t)4 * G * f * f * 4 * ((C + 3) / 4)) * sizeof(float);
  return one <= 160 * 1024 && (size_t)M * O * O < (1u << 24) && (size_t)C * S * S < (1u << 24);
}

struct SmallPlan {
  int NB, nt, grid, L, nseg, nsl, npl, team;
  size_t smem;
};

static double eff(int items, int lanes) {
  return (double)items / ((double)ceil_div(items, lanes) * lanes);
}

static SmallPlan plan_fprop(int B, int C, int S, int M, int O) {
  // (images per CTA, threads per CTA): items = one pooled cell x 4 maps each.  Small CTAs that
  // are all resident at once (many warps per SM, one item per thread) beat big ones: the batch
  // is only ~7 images per SM, there is nothing to amortise a big CTA over.
  const int G = (M + 3) / 4, Pc = (O + 1) / 2, Sp = (S + 1) & ~1;
  SmallPlan best{};
  double best_cost = 1e30;
  const int forced = env_int("TN_SMALL_NB_F", 0), forced_nt = env_int("TN_SMALL_NT_F", 0);
  for (int NB = 1; NB <= 64; ++NB) {
    const size_t smem = ((size_t)C * kSF * kSF * 4 * G + 4 * G + (size_t)NB * C * S * Sp + guard_x(S)) *
                        sizeof(float);
    if (NB > 1 && smem > 64 * 1024) break;
    const int items = G * NB * Pc * Pc;
    if (items >= (1 << 24)) break;
    const int groups = ceil_div(B, NB);
    for (int nt = 64; nt <= kST; nt += 32) {
      if (forced_nt && nt != forced_nt) continue;
      const int by_smem = (int)std::max<size_t>(1, (200 * 1024) / std::max<size_t>(smem, 1));
      const int resident = std::max(1, std::min(std::min(2048 / nt, 65536 / (48 * nt)), std::min(by_smem, 16)));
      const int per_sm = ceil_div(groups, kNumSM);
      const int live = std::min(per_sm, resident);
      const int waves = ceil_div(per_sm, resident);
      // time ~ CTAs an SM works through x (work per CTA / lane efficiency + a fixed per-CTA part)
      double cost = (double)per_sm * ((double)NB / eff(items, nt) + 0.15) * (waves > 1 ? 1.1 : 1.0);
      if (live * (nt / 32) < 16) cost *= 1.3;      // too few warps to hide latency
      if ((forced == 0 && cost < best_cost) || (forced == NB && cost < best_cost)) {
        best_cost = cost;
        best.NB = NB;
        best.nt = nt;
        best.grid = std::min(groups, kNumSM * resident);
        best.smem = smem;
      }
    }
    if (forced == NB) break;
  }
  return best;
}

static SmallPlan plan_bwd(int B, int C, int S, int M, int O, bool need_dx) {
  const int f = kSF;
  const int G = (M + 3) / 4, CG = (C + 3) / 4, Hp = O + 2 * (f - 1);
  int ps = 4 * G;
  if (((ps / 4) & 1) == 0) ps += 4;
  const int T = G * C, nsl = std::max(1, kST / T);
  const int RP = (O + 1) / 2, strips = ceil_div(S, 4);
  SmallPlan best{};
  double best_cost = 1e30;
  const int forced = env_int("TN_SMALL_NB_B", 0);
  for (int NB = 1; NB <= 64; ++NB) {
    const size_t lay = ((size_t)4 * kST + (size_t)NB * Hp * Hp * ps + 4 * ps +
                        ((size_t)NB * C * S * S + guard_x(S) + 3) / 4 * 4 +
                        (need_dx ? (size_t)4 * G * f * f * 4 * CG : 0)) * sizeof(float);
    const size_t red = ((size_t)4 * kST + (size_t)nsl * (T * 4 * f * f + 1)) * sizeof(float);
    const size_t smem = std::max(lay, red);
    if (NB > 1 && lay > 100 * 1024) break;
    // weight gradient: units = (image, row pair, segment); the longest segment that keeps the
    // slices busy wins (every segment re-primes its sliding window)
    int bestL = O;
    double best_e = -1.0;
    for (int nseg = 1; nseg <= O; ++nseg) {
      const int L = ceil_div(O, nseg);
      if (ceil_div(O, L) != nseg) continue;
      const double e = eff(NB * RP * nseg, nsl) * ((double)L / (L + 1.0));
      if (e > best_e + 1e-9) { best_e = e; bestL = L; }
    }
    const double w_work = (double)T * 36.0 * O * O / best_e;
    const double d_work = need_dx ? (double)G * CG * 192.0 * 3.0 * S * strips / eff(CG * NB * S * strips, kST) : 0.0;
    const double s_work = 30.0 * M * ((O + 1) / 2) * ((O + 1) / 2);
    const int groups = ceil_div(B, NB);
    const int resident = (int)std::min<size_t>(4, std::max<size_t>(1, (200 * 1024) / smem));
    const int per_sm = ceil_div(groups, kNumSM);
    double cost = (double)per_sm * NB * (w_work + d_work + s_work);
    if (std::min(per_sm, resident) * (kST / 32) < 12) cost *= 1.2;
    if ((forced == 0 && cost < best_cost) || forced == NB) {
      best_cost = forced == NB ? -1.0 : cost;
      best.NB = NB;
      best.grid = std::min(groups, kNumSM * resident);
      if (env_int("TN_SMALL_GRID_B", 0) > 0) best.grid = std::min(groups, env_int("TN_SMALL_GRID_B", 0));
      best.smem = smem;
      best.L = bestL;
      best.nseg = ceil_div(O, bestL);
      best.nsl = nsl;
      best.npl = std::max(1, kST / G);
    }
  }
  int team = 1;
  while (team * team < best.grid) ++team;
  best.team = std::min(team, 32);
  return best;
}

static void fill_small(SmallArgs &k, int B, int C, int S, int M, int O, int P, int act, int act_nn) {
  k.B = B; k.C = C; k.S = S; k.M = M; k.O = O; k.P = P; k.Pc = (O + 1) / 2;
  k.G = (M + 3) / 4; k.CG = (C + 3) / 4;
  k.Hp = O + 2 * (kSF - 1);
  k.ps = 4 * k.G;
  if (((k.ps / 4) & 1) == 0) k.ps += 4;
  k.Sp = (S + 1) & ~1;
  k.gx = guard_x(S);
  k.gg = 4 * k.ps;
  k.T = k.G * C;
  k.strips = ceil_div(S, 4);
  k.ak = make_actk(act, act_nn);
  k.dPcPc = FastDiv32((uint32_t)(k.Pc * k.Pc)); k.dPc = FastDiv32((uint32_t)k.Pc);
  k.dPP = FastDiv32((uint32_t)(P * P)); k.dP = FastDiv32((uint32_t)P); k.dM = FastDiv32((uint32_t)M);
  k.dStrips = FastDiv32((uint32_t)k.strips); k.dS = FastDiv32((uint32_t)S);
  k.dT = FastDiv32((uint32_t)k.T);
  k.dHH = FastDiv32((uint32_t)(k.Hp * k.Hp)); k.dH = FastDiv32((uint32_t)k.Hp);
  k.dCFF = FastDiv32((uint32_t)(C * kSF * kSF)); k.dNW = FastDiv32((uint32_t)(k.T * 4 * kSF * kSF));
}

template <typename K>
static int small_smem_attr(K kernel, size_t smem, const char *who) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  }
  return TN_OK;
}

// 1 when `grid` CTAs of `kernel` (kST threads, `smem` dynamic bytes) fit on the device at once,
// i.e. a grid-wide barrier inside the kernel cannot deadlock (TN_SMALL_COOP=0: never)
template <typename K>
static int small_coresident(K kernel, int grid, size_t smem) {
  if (env_int("TN_SMALL_COOP", 1) == 0) return 0;
  int per_sm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kST, smem) != cudaSuccess ||
      cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (long long)per_sm * sms >= grid ? 1 : 0;
}

int small_fprop(const float *x, const float *W, const float *bias, float *a, float *pooled,
                uint8_t *tie, int B, int C, int S, int M, int O, int act, int act_nn, int P,
                cudaStream_t st) {
  const char *who = "tn_convpool_fprop(small)";
  SmallArgs k{};
  fill_small(k, B, C, S, M, O, P, act, act_nn);
  const SmallPlan pl = plan_fprop(B, C, S, M, O);
  k.NB = pl.NB;
  k.dNB = FastDiv32((uint32_t)pl.NB);
  k.x = x; k.W = W; k.bias = bias; k.a = a; k.pooled = pooled; k.tie = tie;
  int rc = small_smem_attr(small_fprop_kernel<kSF>, pl.smem, who);
  if (rc) return rc;
  launch_pdl(small_fprop_kernel<kSF>, dim3(pl.grid), dim3(pl.nt), pl.smem, st, k);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

static size_t small_bwd_workspace_bytes_for(const SmallPlan &pl, int C, int M) {
  const int G = (M + 3) / 4;
  const int nout4 = (G * C * 4 * kSF * kSF + 4 * G + 3) / 4 * 4;
  const int nteam = ceil_div(pl.grid, pl.team);
  return ((size_t)pl.grid * nout4 + (size_t)nteam * nout4) * sizeof(float) +
         (size_t)(nteam + 1 + 2 + 3) / 4 * 4 * sizeof(unsigned);   // tickets + grid barrier
}

// weight gradient without dx (small_wgrad_kernel): only the images are staged
static SmallPlan plan_wgrad(int B, int C, int S, int M, int O, int P) {
  const int G = (M + 3) / 4, T = G * C, nsl = std::max(1, kST / T), Sp = (S + 1) & ~1;
  SmallPlan best{};
  double best_cost = 1e30;
  const int forced = env_int("TN_SMALL_NB_W", 0);
  for (int NB = 1; NB <= 64; ++NB) {
    const size_t lay = ((size_t)4 * kST + (size_t)NB * C * S * Sp + guard_x(S)) * sizeof(float);
    const size_t red = ((size_t)4 * kST + (size_t)nsl * (T * 4 * kSF * kSF + 1)) * sizeof(float);
    const size_t smem = std::max(lay, red);
    if (NB > 1 && lay > 64 * 1024) break;
    const int groups = ceil_div(B, NB);
    const int resident = (int)std::min<size_t>(2, std::max<size_t>(1, (200 * 1024) / smem));
    const int per_sm = ceil_div(groups, kNumSM);
    double cost = (double)per_sm * NB / eff(NB * P * P, nsl);
    if (std::min(per_sm, resident) * (kST / 32) < 12) cost *= 1.25;
    if ((forced == 0 && cost < best_cost) || forced == NB) {
      best_cost = forced == NB ? -1.0 : cost;
      best.NB = NB;
      best.grid = std::min(groups, kNumSM * resident);
      best.smem = smem;
      best.nsl = nsl;
      best.npl = nsl;
    }
  }
  int team = 1;
  while (team * team < best.grid) ++team;
  best.team = std::min(team, 32);
  return best;
}

static bool direct_wgrad(bool need_dx) { return !need_dx && env_int("TN_SMALL_DIRECT_WGRAD", 1) != 0; }

size_t small_bwd_workspace_bytes(int B, int C, int S, int M, int O, int P, bool need_dx) {
  const SmallPlan pl = direct_wgrad(need_dx) ? plan_wgrad(B, C, S, M, O, P) : plan_bwd(B, C, S, M, O, need_dx);
  return small_bwd_workspace_bytes_for(pl, C, M);
}

int small_bwd(const float *x, const float *a, const uint8_t *tie, const float *pooled,
              const float *dtop, const float *W, float *dW, float *db, float *dx,
              const float *below, void *workspace, int B, int C,
              int S, int M, int O, int act, int act_nn, int P, int act_below, int nn_below,
              cudaStream_t st) {
  const char *who = "tn_convpool_bwd(small)";
  SmallArgs k{};
  fill_small(k, B, C, S, M, O, P, act, act_nn);
  const bool direct = direct_wgrad(dx != nullptr);
  const SmallPlan pl = direct ? plan_wgrad(B, C, S, M, O, P) : plan_bwd(B, C, S, M, O, dx != nullptr);
  k.NB = pl.NB;
  k.dNB = FastDiv32((uint32_t)pl.NB);
  k.nsl = pl.nsl; k.L = pl.L; k.nseg = pl.nseg; k.npl = pl.npl; k.team = pl.team;
  k.upi = ((O + 1) / 2) * pl.nseg;
  k.dUPI = FastDiv32((uint32_t)k.upi); k.dNSEG = FastDiv32((uint32_t)pl.nseg);
  k.nout = k.T * 4 * kSF * kSF + 4 * k.G;
  k.nout4 = (k.nout + 3) / 4 * 4;
  k.x = x; k.a = const_cast<float *>(a); k.tie = const_cast<uint8_t *>(tie);
  k.pooled = const_cast<float *>(pooled); k.dtop = dtop;

  k.W = W; k.dW = dW; k.db = db; k.dx = dx; k.below = below;
  k.akb = make_actk(act_below, nn_below);
  k.dbg = g_small_dbg;
  const int nteam = ceil_div(pl.grid, pl.team);
  k.partial = (float *)workspace;
  k.teampart = k.partial + (size_t)pl.grid * k.nout4;
  k.tickets = reinterpret_cast<unsigned *>(k.teampart + (size_t)nteam * k.nout4);
  k.gbar = k.tickets + nteam + 1;
  if (direct) {
    int rc = small_smem_attr(small_wgrad_kernel<kSF>, pl.smem, who);
    if (rc) return rc;
    k.coop = small_coresident(small_wgrad_kernel<kSF>, pl.grid, pl.smem);
    launch_pdl(small_wgrad_kernel<kSF>, dim3(pl.grid), dim3(kST), pl.smem, st, k);
    TN_LAUNCH_CHECK(who);
    return TN_OK;
  }
  int rc = small_smem_attr(small_bwd_kernel<kSF>, pl.smem, who);
  if (rc) return rc;
  k.coop = small_coresident(small_bwd_kernel<kSF>, pl.grid, pl.smem);
  launch_pdl(small_bwd_kernel<kSF>, dim3(pl.grid), dim3(kST), pl.smem, st, k);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

}  // namespace tn

using namespace tn;

extern "C" int tn_convpool_small_supported(int C, int S, int M, int f, int pad_lo, int out_sz,
                                           int act, int pool, int pool_out_sz) {
  return small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz) ? 1 : 0;
}

extern "C" size_t tn_convpool_bwd_workspace_bytes(int B, int C, int S, int M, int f, int pad_lo,
                                                  int out_sz, int act, int pool, int pool_out_sz,
                                                  int need_dx) {
  if (!small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz)) return 0;
  return small_bwd_workspace_bytes(B, C, S, M, out_sz, pool_out_sz, need_dx != 0);
}

extern "C" int tn_convpool_fprop_train(const float *x, const float *W, const float *bias, float *a,
                                       float *pooled, uint8_t *tie, int B, int C, int S, int M,
                                       int f, int pad_lo, int out_sz, int act, int act_nn,
                                       int pool, int pool_out_sz, void *stream) {
  const char *who = "tn_convpool_fprop_train";
  TN_REQUIRE(x && W && bias && pooled, TN_ERR_ARG, "%s: null argument", who);
  TN_REQUIRE(B > 0 && (int64_t)B * std::max(M * out_sz * out_sz, C * S * S) < (1ll << 31),
             TN_ERR_SHAPE, "%s: bad batch size %d", who, B);
  TN_REQUIRE(small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz), TN_ERR_UNSUPPORTED,
             "%s: needs filter_sz 3, mode 'valid', pool 2 and a ReLU-family activation (got f=%d "
             "pad=%d pool=%d act=%d); use tn_convpool_fprop", who, f, pad_lo, pool, act);
  return small_fprop(x, W, bias, a, pooled, tie, B, C, S, M, out_sz, act, act_nn, pool_out_sz,
                     (cudaStream_t)stream);
}

extern "C" int tn_convpool_bwd(const float *x, const float *a, const uint8_t *tie,
                               const float *pooled, const float *dtop, const float *W, float *dW,
                               float *db, float *dx, const float *below, void *workspace, int B,
                               int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                               int act_nn, int pool, int pool_out_sz, int act_below, int nn_below,
                               void *stream) {
  const char *who = "tn_convpool_bwd";
  TN_REQUIRE(x && (a || tie) && pooled && dtop && W && dW && db && workspace, TN_ERR_ARG,
             "%s: null argument", who);
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0, TN_ERR_SHAPE, "%s: bad shape", who);
  TN_REQUIRE(small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz), TN_ERR_UNSUPPORTED,
             "%s: needs filter_sz 3, mode 'valid', pool 2 and a ReLU-family activation (got f=%d "
             "pad=%d pool=%d act=%d); use tn_convpool_bwd_weights / tn_convpool_bwd_data", who, f,
             pad_lo, pool, act);
  TN_REQUIRE((int64_t)B * std::max(M * out_sz * out_sz, C * S * S) < (1ll << 31), TN_ERR_SHAPE,
             "%s: batch %d too large for 32-bit offsets", who, B);
  TN_REQUIRE(!below || act_is_fast(act_below), TN_ERR_UNSUPPORTED,
             "%s: activation %d of the layer below is not on this path", who, act_below);
  return small_bwd(x, a, tie, pooled, dtop, W, dW, db, dx, below, workspace, B, C, S, M, out_sz, act,
                   act_nn, pool_out_sz, act_below, nn_below, (cudaStream_t)stream);
}

/* Debug aid (tools/phase_times.py): when set, thread 0 of every CTA of tn_convpool_bwd stores
 * clock64() at its phase boundaries into buf[cta * 16 + phase]; NULL switches it off. */
extern "C" int tn_convpool_debug_timestamps(long long *buf) {
  g_small_dbg = buf;
  return TN_OK;
}
