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

struct SmallArgs {
  const float *x, *W, *bias, *dtop, *below;
  float *a, *pooled, *dx, *dW, *db;
  float *partial, *teampart;
  unsigned *tickets;
  int B, C, S, M, O, P, Pc, NB;
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
  FastDiv32 dPcPc, dPc, dNB, dPP, dP, dM, dUPI, dNSEG, dStrips, dS;
};

// dst (16-byte aligned shared memory) <- n consecutive floats at src
__device__ __forceinline__ void stage_contig(float *dst, const float *__restrict__ src, int n) {
  const int tid = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = n >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int t = tid; t < n4; t += kST) d4[t] = __ldg(s4 + t);
    for (int t = 4 * n4 + tid; t < n; t += kST) dst[t] = __ldg(src + t);
  } else {
    for (int t = tid; t < n; t += kST) dst[t] = __ldg(src + t);
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// smem: ws[(c*F+u)*F+v][4G] (taps flipped: true convolution) | xs[NB][C][S][S] + guard
template <int F>
__global__ void __launch_bounds__(kST) small_fprop_kernel(const SmallArgs k) {
  extern __shared__ __align__(16) float sm[];
  const int C = k.C, S = k.S, M = k.M, O = k.O, P = k.P, Pc = k.Pc, G = k.G, NB = k.NB;
  const int coP = 4 * G, SS = S * S, CSS = C * SS, OO = O * O, PP = P * P, PcPc = Pc * Pc;
  float *ws = sm;
  float *xs = ws + ((C * F * F * coP + 3) & ~3);
  const int tid = threadIdx.x;
  for (int t = tid; t < C * F * F * coP; t += kST) {
    const int co = t % coP;
    int r = t / coP;
    const int v = r % F; r /= F;
    const int u = r % F;
    const int c = r / F;
    ws[t] = co < M ? k.W[((co * C + c) * F + (F - 1 - u)) * F + (F - 1 - v)] : 0.f;
  }
  for (int t = tid; t < k.gx; t += kST) xs[NB * CSS + t] = 0.f;
  const int items = G * NB * PcPc;   // (g, b, pi, pj), pj fastest: a warp shares its taps

  for (int grp = blockIdx.x; grp * NB < k.B; grp += gridDim.x) {
    const int b0 = grp * NB, nb = min(NB, k.B - b0);
    stage_contig(xs, k.x + (size_t)b0 * CSS, nb * CSS);
    __syncthreads();
    for (int it = tid; it < items; it += kST) {
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
      const float *xb = xs + b * CSS + (2 * pi) * S + 2 * pj;
      const float4 *w4 = reinterpret_cast<const float4 *>(ws) + g;
      for (int c = 0; c < C; ++c) {
        float p[F + 1][F + 1];   // rows/cols past the image only feed outputs that are dropped
#pragma unroll
        for (int r = 0; r <= F; ++r)
#pragma unroll
          for (int e = 0; e <= F; ++e) p[r][e] = xb[r * S + e];
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
        xb += SS;
      }
      const int i0 = 2 * pi, j0 = 2 * pj;
      const bool r1 = i0 + 1 < O, c1 = j0 + 1 < O;
      const bool pool_ok = pi < P && pj < P;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = 4 * g + q;
        if (m >= M) break;
        const float bm = __ldg(k.bias + m);
        float *ar = k.a + ((size_t)(b0 + b) * M + m) * OO + i0 * O + j0;
        const float v00 = act_fwd_t<false>(k.ak, acc[0][q] + bm);
        float mx = v00;
        ar[0] = v00;
        if (c1) {
          const float v01 = act_fwd_t<false>(k.ak, acc[1][q] + bm);
          ar[1] = v01;
          mx = fmaxf(mx, v01);
        }
        if (r1) {
          const float v10 = act_fwd_t<false>(k.ak, acc[2][q] + bm);
          ar[O] = v10;
          mx = fmaxf(mx, v10);
          if (c1) {
            const float v11 = act_fwd_t<false>(k.ak, acc[3][q] + bm);
            ar[O + 1] = v11;
            mx = fmaxf(mx, v11);
          }
        }
        if (pool_ok) k.pooled[((size_t)(b0 + b) * M + m) * PP + pi * P + pj] = mx;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward: dW, db (and dx when the layer below needs it) in one launch
// ---------------------------------------------------------------------------------------------
// smem: gz[NB][Hp][Hp][ps] + guard | xs[NB][C][S][S] + guard | wd[(m*F+u)*F+v][4CG]
template <int F>
__global__ void __launch_bounds__(kST) small_bwd_kernel(const SmallArgs k) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_flag;
  constexpr int FF = F * F;
  constexpr int pd = F - 1;
  const int C = k.C, S = k.S, M = k.M, O = k.O, P = k.P, G = k.G, CG = k.CG, NB = k.NB;
  const int Hp = k.Hp, ps = k.ps, cP = 4 * CG, mP = 4 * G;
  const int SS = S * S, CSS = C * SS, OO = O * O, PP = P * P;
  const int gzimg = Hp * Hp * ps;
  float *gz = sm;
  float *xs = gz + NB * gzimg + k.gg;
  float *wd = xs + ((NB * CSS + k.gx + 3) & ~3);
  const int tid = threadIdx.x;
  const bool need_dx = k.dx != nullptr;

  for (int t = tid; t < NB * gzimg + k.gg; t += kST) gz[t] = 0.f;   // borders stay zero
  for (int t = tid; t < k.gx; t += kST) xs[NB * CSS + t] = 0.f;
  if (need_dx) {
    for (int t = tid; t < mP * FF * cP; t += kST) {
      const int co = t % cP;
      int r = t / cP;
      const int v = r % F; r /= F;
      const int u = r % F;
      const int m = r / F;
      wd[t] = (co < C && m < M) ? k.W[((m * C + co) * F + u) * F + v] : 0.f;
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
  float dba[4] = {0.f, 0.f, 0.f, 0.f};
  const int grow = Hp * ps;
  __syncthreads();

  for (int grp = blockIdx.x; grp * NB < k.B; grp += gridDim.x) {
    const int b0 = grp * NB, nb = min(NB, k.B - b0);
    stage_contig(xs, k.x + (size_t)b0 * CSS, nb * CSS);
    {  // dL/dz of the conv layer from (a, pooled, dL/dpooled): one thread per pooled cell
      const float *pp = k.pooled + (size_t)b0 * M * PP;
      const float *dp = k.dtop + (size_t)b0 * M * PP;
      const float *ap = k.a + (size_t)b0 * M * OO;
      const int ncell = nb * M * PP;
      for (int t = tid; t < ncell; t += kST) {
        const float po = __ldg(pp + t), d = __ldg(dp + t);
        const int bm = (int)k.dPP.div(t);
        const int p = t - bm * PP;
        const int pi = (int)k.dP.div(p), pj = p - pi * P;
        const int b = (int)k.dM.div(bm), m = bm - b * M;
        const float gg = d * act_bwd_t<false>(k.ak, po);
        const float *ar = ap + (size_t)bm * OO + (2 * pi) * O + 2 * pj;
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
    __syncthreads();

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
    if (dbact) {   // the border pixels are zero: one linear pass over the whole bordered maps
      const int npx = nb * Hp * Hp;
      for (int px = pl; px < npx; px += k.npl) {
        const float4 g = *reinterpret_cast<const float4 *>(gz + px * ps + 4 * mgd);
        dba[0] += g.x; dba[1] += g.y; dba[2] += g.z; dba[3] += g.w;
      }
    }

    if (need_dx) {
      // dx[c,y,x] = sum_{m,u,v} gzb[m,y+u,x+v] W[m,c,u,v]; item = (channel group, image, row, strip)
      const int nit = CG * NB * S * k.strips;
      for (int it = tid; it < nit; it += kST) {
        const int r = (int)k.dStrips.div(it);
        const int s = it - r * k.strips;
        const int r2 = (int)k.dS.div(r);
        const int y = r - r2 * S;
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
          const size_t o = ((size_t)(b0 + b) * C + ch) * SS + y * S + x0;
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
    __syncthreads();
  }

  // ---- per-CTA partial: row slices combined in a fixed order -------------------------------
  float *red = sm;   // [nsl][T][4*FF], aliases the staging buffers (the loop ended on a barrier)
  if (wact) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < FF; ++e) red[(slice * T + combo) * 4 * FF + q * FF + e] = acc[q][e];
  }
  __syncthreads();
  const int nW = T * 4 * FF;
  float *pout = k.partial + (size_t)blockIdx.x * k.nout4;
  for (int t = tid; t < nW; t += kST) {
    float s = 0.f;
    for (int sl = 0; sl < k.nsl; ++sl) s += red[sl * nW + t];
    pout[t] = s;
  }
  __syncthreads();
  if (dbact) {
#pragma unroll
    for (int q = 0; q < 4; ++q) red[pl * mP + 4 * mgd + q] = dba[q];
  }
  __syncthreads();
  if (tid < mP) {
    float s = 0.f;
    for (int p = 0; p < k.npl; ++p) s += red[p * mP + tid];
    pout[nW + tid] = s;
  }
  for (int t = nW + mP + tid; t < k.nout4; t += kST) pout[t] = 0.f;

  // ---- two-level ticket: last CTA of a team sums the team, last team sums the teams --------
  const int team = blockIdx.x / k.team;
  const int nteam = (gridDim.x + k.team - 1) / k.team;
  const int tsize = min(k.team, (int)gridDim.x - team * k.team);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = atomicAdd(&k.tickets[team], 1u) == (unsigned)(tsize - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const int n4 = k.nout4 >> 2;
  {
    const float4 *src = reinterpret_cast<const float4 *>(k.partial) + (size_t)team * k.team * n4;
    float4 *dst = reinterpret_cast<float4 *>(k.teampart) + (size_t)team * n4;
    for (int t = tid; t < n4; t += kST) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
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
#pragma unroll 8
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
}

// ---------------------------------------------------------------------------------------------
// host side: geometry, work split, launch
// ---------------------------------------------------------------------------------------------
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
  const size_t one = ((size_t)Hp * Hp * ps + 4 * ps + (size_t)C * S * S + guard_x(S) + 4 +
                      (size_t)4 * G * f * f * 4 * ((C + 3) / 4)) * sizeof(float);
  return one <= 160 * 1024 && (size_t)M * O * O < (1u << 24);
}

struct SmallPlan {
  int NB, grid, L, nseg, nsl, npl, team;
  size_t smem;
};

static double eff(int items, int lanes) {
  return (double)items / ((double)ceil_div(items, lanes) * lanes);
}

static SmallPlan plan_fprop(int B, int C, int S, int M, int O) {
  const int G = (M + 3) / 4, Pc = (O + 1) / 2;
  SmallPlan best{};
  double best_cost = 1e30;
  const int forced = env_int("TN_SMALL_NB_F", 0);
  for (int NB = 1; NB <= 64; ++NB) {
    const size_t smem = (((size_t)C * kSF * kSF * 4 * G + 3) / 4 * 4 + (size_t)NB * C * S * S + guard_x(S)) *
                        sizeof(float);
    if (NB > 1 && smem > 64 * 1024) break;
    if ((int64_t)G * NB * Pc * Pc >= (1 << 24)) break;
    const int groups = ceil_div(B, NB);
    const int resident = (int)std::min<size_t>(8, std::max<size_t>(1, (200 * 1024) / std::max<size_t>(smem, 1)));
    const int per_sm = ceil_div(groups, kNumSM);               // CTAs the busiest SM works through
    double cost = (double)per_sm * NB / eff(G * NB * Pc * Pc, kST);
    if (std::min(per_sm, resident) * (kST / 32) < 12) cost *= 1.25;   // too few warps to hide latency
    if ((forced == 0 && cost < best_cost) || forced == NB) {
      best_cost = forced == NB ? -1.0 : cost;
      best.NB = NB;
      best.grid = std::min(groups, kNumSM * resident);
      best.smem = smem;
    }
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
    const size_t lay = ((size_t)NB * Hp * Hp * ps + 4 * ps + ((size_t)NB * C * S * S + guard_x(S) + 3) / 4 * 4 +
                        (need_dx ? (size_t)4 * G * f * f * 4 * CG : 0)) * sizeof(float);
    const size_t red = (size_t)nsl * T * 4 * f * f * sizeof(float);
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
  k.gx = guard_x(S);
  k.gg = 4 * k.ps;
  k.T = k.G * C;
  k.strips = ceil_div(S, 4);
  k.ak = make_actk(act, act_nn);
  k.dPcPc = FastDiv32((uint32_t)(k.Pc * k.Pc)); k.dPc = FastDiv32((uint32_t)k.Pc);
  k.dPP = FastDiv32((uint32_t)(P * P)); k.dP = FastDiv32((uint32_t)P); k.dM = FastDiv32((uint32_t)M);
  k.dStrips = FastDiv32((uint32_t)k.strips); k.dS = FastDiv32((uint32_t)S);
}

template <typename K>
static int small_smem_attr(K kernel, size_t smem, const char *who) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  }
  return TN_OK;
}

int small_fprop(const float *x, const float *W, const float *bias, float *a, float *pooled, int B,
                int C, int S, int M, int O, int act, int act_nn, int P, cudaStream_t st) {
  const char *who = "tn_convpool_fprop(small)";
  SmallArgs k{};
  fill_small(k, B, C, S, M, O, P, act, act_nn);
  const SmallPlan pl = plan_fprop(B, C, S, M, O);
  k.NB = pl.NB;
  k.dNB = FastDiv32((uint32_t)pl.NB);
  k.x = x; k.W = W; k.bias = bias; k.a = a; k.pooled = pooled;
  int rc = small_smem_attr(small_fprop_kernel<kSF>, pl.smem, who);
  if (rc) return rc;
  small_fprop_kernel<kSF><<<pl.grid, kST, pl.smem, st>>>(k);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

size_t small_bwd_workspace_bytes(int B, int C, int S, int M, int O, bool need_dx) {
  const SmallPlan pl = plan_bwd(B, C, S, M, O, need_dx);
  const int G = (M + 3) / 4;
  const int nout4 = (G * C * 4 * kSF * kSF + 4 * G + 3) / 4 * 4;
  const int nteam = ceil_div(pl.grid, pl.team);
  return ((size_t)pl.grid * nout4 + (size_t)nteam * nout4) * sizeof(float) +
         (size_t)(nteam + 1 + 3) / 4 * 4 * sizeof(unsigned);
}

int small_bwd(const float *x, const float *a, const float *pooled, const float *dtop, const float *W,
              float *dW, float *db, float *dx, const float *below, void *workspace, int B, int C,
              int S, int M, int O, int act, int act_nn, int P, int act_below, int nn_below,
              cudaStream_t st) {
  const char *who = "tn_convpool_bwd(small)";
  SmallArgs k{};
  fill_small(k, B, C, S, M, O, P, act, act_nn);
  const SmallPlan pl = plan_bwd(B, C, S, M, O, dx != nullptr);
  k.NB = pl.NB;
  k.dNB = FastDiv32((uint32_t)pl.NB);
  k.nsl = pl.nsl; k.L = pl.L; k.nseg = pl.nseg; k.npl = pl.npl; k.team = pl.team;
  k.upi = ((O + 1) / 2) * pl.nseg;
  k.dUPI = FastDiv32((uint32_t)k.upi); k.dNSEG = FastDiv32((uint32_t)pl.nseg);
  k.nout = k.T * 4 * kSF * kSF + 4 * k.G;
  k.nout4 = (k.nout + 3) / 4 * 4;
  k.x = x; k.a = const_cast<float *>(a); k.pooled = const_cast<float *>(pooled); k.dtop = dtop;
  k.W = W; k.dW = dW; k.db = db; k.dx = dx; k.below = below;
  k.akb = make_actk(act_below, nn_below);
  const int nteam = ceil_div(pl.grid, pl.team);
  k.partial = (float *)workspace;
  k.teampart = k.partial + (size_t)pl.grid * k.nout4;
  k.tickets = reinterpret_cast<unsigned *>(k.teampart + (size_t)nteam * k.nout4);
  int rc = small_smem_attr(small_bwd_kernel<kSF>, pl.smem, who);
  if (rc) return rc;
  small_bwd_kernel<kSF><<<pl.grid, kST, pl.smem, st>>>(k);
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
  return small_bwd_workspace_bytes(B, C, S, M, out_sz, need_dx != 0);
}

extern "C" int tn_convpool_bwd(const float *x, const float *a, const float *pooled,
                               const float *dtop, const float *W, float *dW, float *db, float *dx,
                               const float *below, void *workspace, int B, int C, int S, int M,
                               int f, int pad_lo, int out_sz, int act, int act_nn, int pool,
                               int pool_out_sz, int act_below, int nn_below, void *stream) {
  const char *who = "tn_convpool_bwd";
  TN_REQUIRE(x && a && pooled && dtop && W && dW && db && workspace, TN_ERR_ARG, "%s: null argument",
             who);
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0, TN_ERR_SHAPE, "%s: bad shape", who);
  TN_REQUIRE(small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz), TN_ERR_UNSUPPORTED,
             "%s: needs filter_sz 3, mode 'valid', pool 2 and a ReLU-family activation (got f=%d "
             "pad=%d pool=%d act=%d); use tn_convpool_bwd_weights / tn_convpool_bwd_data", who, f,
             pad_lo, pool, act);
  TN_REQUIRE(!below || act_is_fast(act_below), TN_ERR_UNSUPPORTED,
             "%s: activation %d of the layer below is not on this path", who, act_below);
  return small_bwd(x, a, pooled, dtop, W, dW, db, dx, below, workspace, B, C, S, M, out_sz, act,
                   act_nn, pool_out_sz, act_below, nn_below, (cudaStream_t)stream);
}
