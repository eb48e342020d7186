// ConvLayer (+ PoolLayer) fused kernels for small channel counts (theanet/layer/convpool.py:14-127).
//
// theanet's shipped networks run conv -> leaky-ReLU -> max-pool with 1..20 maps of 3x3 filters
// (params/mnist.prms:13-27): a few kFLOP per pixel, HBM/latency-bound.  These kernels keep a whole
// image resident in shared memory and touch HBM once per tensor:
//
//   tn_convpool_fprop       x -> a = act(conv(x) + b) and pooled = maxpool(a)          (1 launch)
//   tn_convpool_bwd_weights (x, a, pooled, dL/dpooled) -> dW, db                        (2 launches)
//   tn_convpool_bwd_data    (a, pooled, dL/dpooled, W) -> dL/dx [* act'(x) of a conv below]
//
// The backward kernels rebuild dL/dz = [a == pooled(window)] * dL/dpooled * act'(a) (Theano's
// MaxPoolGrad: every tied maximum receives the gradient) while staging the image, so the
// un-pooled gradient tensor is never written to or read from HBM.  Inner loops are register-tiled
// (4 output channels x a strip of 4 pixels with a sliding window over the filter row) so that
// shared-memory loads per FMA stay below the LDS issue rate.  Reductions use fixed orders: results
// are deterministic run to run.
#include <algorithm>

#include "common.cuh"
#include "conv_small.cuh"

namespace tn {

constexpr int kFT = 256;  // threads per CTA
constexpr int kL = 4;     // pixels per register strip

// n / d for n, d < 65536: one multiply-high
struct FastDiv {
  uint32_t d, magic;
  __host__ FastDiv() : d(1), magic(0) {}
  __host__ explicit FastDiv(int dd) : d((uint32_t)dd), magic((uint32_t)((0x100000000ull + dd - 1) / dd)) {}
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : __umulhi(n, magic); }
};

struct FusedArgs {
  const float *x;       // layer input (B, C, S, S)
  const float *W;       // OIHW (M, C, f, f)
  const float *bias;    // fprop
  float *a;             // layer output act(conv + b), (B, M, O, O): written by fprop, read by bwd
  float *pooled;        // (B, M, P, P): written by fprop, read by bwd (null: no pool layer)
  const float *dtop;    // bwd: dL/dpooled (B, M, P, P), or dL/da (B, M, O, O) without a pool layer
  float *dx;            // bwd-data output (B, C, S, S)
  const float *below;   // bwd-data: output of the conv layer directly below (multiply by its act')
  float *partial;       // bwd-weights: per-CTA partial sums
  int B, C, S, M, f, pad_lo, O, P, pool;  // pool = window (0: none)
  int act, act_below;
  float nn, nn_below;
  ActK ak, akb;         // activation of this layer / of the conv below, leaky slopes hoisted
  FastDiv divO, divOO, divP, divPP, divS, divSp, divstrO, divstrS;
};

// gs[(m*ldm + i + pd)*ldw + j + pd] = dL/dz of image element (m, i, j) (see header comment).  With
// a pool layer the loop runs over pooled cells (coalesced reads of pooled / dtop, index math
// amortised over the window); elements outside every window (ignore_border) are never written, so
// gs must have been zero-filled once.
template <bool GEN>
__device__ __forceinline__ void stage_gz(const FusedArgs &k, const float *__restrict__ a_img,
                                         const float *__restrict__ p_img,
                                         const float *__restrict__ d_img, float *gs, int ldm,
                                         int ldw, int pd) {
  const int tid = threadIdx.x;
  if (k.pool) {
    const int PP = k.P * k.P;
    for (int t = tid; t < k.M * PP; t += kFT) {
      const int m = (int)k.divPP.div(t);
      const int p = t - m * PP;
      const int oi = (int)k.divP.div(p), oj = p - oi * k.P;
      const float po = p_img[t], d = d_img[t];
      const int y0 = oi * k.pool, x0 = oj * k.pool;
      const int y1 = min(y0 + k.pool, k.O), x1 = min(x0 + k.pool, k.O);
      for (int yy = y0; yy < y1; ++yy) {
        const float *ar = a_img + (m * k.O + yy) * k.O;
        float *gr = gs + (m * ldm + yy + pd) * ldw + pd;
        for (int xx = x0; xx < x1; ++xx) {
          const float av = ar[xx];
          gr[xx] = av == po ? d * act_bwd_t<GEN>(k.ak, av) : 0.f;
        }
      }
    }
  } else {
    const int OO = k.O * k.O;
    for (int t = tid; t < k.M * OO; t += kFT) {
      const int m = (int)k.divOO.div(t);
      const int p = t - m * OO;
      const int i = (int)k.divO.div(p), j = p - i * k.O;
      gs[(m * ldm + i + pd) * ldw + j + pd] = d_img[t] * act_bwd_t<GEN>(k.ak, a_img[t]);
    }
  }
}

// xs[(c*Sp + Y)*ldw + X] = x[c, Y - pad, X - pad] or 0: one padded row per warp pass
__device__ __forceinline__ void stage_x(const FusedArgs &k, const float *__restrict__ img,
                                        float *xs, int Sp, int ldw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < k.C * Sp; row += kFT / 32) {
    const int c = (int)k.divSp.div(row);
    const int y = row - c * Sp - k.pad_lo;
    const bool yok = y >= 0 && y < k.S;
    const float *src = img + (c * k.S + y) * k.S - k.pad_lo;
    float *dst = xs + row * ldw;
    for (int X = lane; X < ldw; X += 32) {
      const int x = X - k.pad_lo;
      dst[X] = (yok && x >= 0 && x < k.S) ? src[X] : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward: conv + bias + activation (+ max-pool)
// ---------------------------------------------------------------------------------------------
// smem: ws[(c*F+u)*F+v][coP] | xs[C][Sp][Wp] (zero padded) | as[M][O][O]
template <int F, bool GEN>
__global__ void __launch_bounds__(kFT) convpool_fprop_kernel(const FusedArgs k) {
  extern __shared__ __align__(16) float sm[];
  const int G = (k.M + 3) >> 2, coP = 4 * G;
  const int strips = (k.O + kL - 1) / kL;
  const int Sp = k.O + F - 1;            // rows of the padded input
  const int Wp = strips * kL + F - 1;    // padded row pitch (strip overhang reads zeros)
  float *ws = sm;
  float *xs = ws + k.C * F * F * coP;
  float *as = xs + ((k.C * Sp * Wp + 3) & ~3);  // 16-byte aligned for the float4 copy-out
  const int tid = threadIdx.x;

  for (int t = tid; t < k.C * F * F * coP; t += kFT) {
    const int co = t % coP;
    int r = t / coP;
    const int v = r % F; r /= F;
    const int u = r % F;
    const int c = r / F;
    ws[t] = co < k.M ? k.W[((co * k.C + c) * F + (F - 1 - u)) * F + (F - 1 - v)] : 0.f;
  }
  const int OO = k.O * k.O;
  const bool vec_a = (((size_t)k.M * OO) & 3) == 0;

  for (int b = blockIdx.x; b < k.B; b += gridDim.x) {
    stage_x(k, k.x + (size_t)b * k.C * k.S * k.S, xs, Sp, Wp);
    __syncthreads();
    // items: (channel group g, output row i, strip s)
    const int items = G * k.O * strips;
    for (int it = tid; it < items; it += kFT) {
      const int r = (int)k.divstrO.div(it);
      const int s = it - r * strips;
      const int g = (int)k.divO.div(r);
      const int i = r - g * k.O;
      const int j0 = s * kL;
      float acc[kL][4];
#pragma unroll
      for (int l = 0; l < kL; ++l)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
      for (int c = 0; c < k.C; ++c) {
#pragma unroll
        for (int u = 0; u < F; ++u) {
          const float *xr = xs + (c * Sp + i + u) * Wp + j0;
          float xv[kL + F - 1];
#pragma unroll
          for (int e = 0; e < kL + F - 1; ++e) xv[e] = xr[e];
          const float4 *wr = reinterpret_cast<const float4 *>(ws + ((c * F + u) * F) * coP) + g;
#pragma unroll
          for (int v = 0; v < F; ++v) {
            const float4 w = wr[v * G];
#pragma unroll
            for (int l = 0; l < kL; ++l) {
              acc[l][0] = fmaf(xv[l + v], w.x, acc[l][0]);
              acc[l][1] = fmaf(xv[l + v], w.y, acc[l][1]);
              acc[l][2] = fmaf(xv[l + v], w.z, acc[l][2]);
              acc[l][3] = fmaf(xv[l + v], w.w, acc[l][3]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = 4 * g + q;
        if (m < k.M) {
          const float bm = k.bias[m];
#pragma unroll
          for (int l = 0; l < kL; ++l)
            if (j0 + l < k.O) as[(m * k.O + i) * k.O + j0 + l] = act_fwd_t<GEN>(k.ak, acc[l][q] + bm);
        }
      }
    }
    __syncthreads();
    // a -> HBM (coalesced), pooled -> HBM
    float *a_img = k.a + (size_t)b * k.M * OO;
    if (vec_a) {
      const float4 *s4 = reinterpret_cast<const float4 *>(as);
      float4 *d4 = reinterpret_cast<float4 *>(a_img);
      for (int t = tid; t < (k.M * OO) >> 2; t += kFT) d4[t] = s4[t];
    } else {
      for (int t = tid; t < k.M * OO; t += kFT) a_img[t] = as[t];
    }
    if (k.pool) {
      const int PP = k.P * k.P;
      float *p_img = k.pooled + (size_t)b * k.M * PP;
      for (int t = tid; t < k.M * PP; t += kFT) {
        const int m = (int)k.divPP.div(t);
        const int p = t - m * PP;
        const int oi = (int)k.divP.div(p), oj = p - oi * k.P;
        const int y0 = oi * k.pool, x0 = oj * k.pool;
        const int y1 = min(y0 + k.pool, k.O), x1 = min(x0 + k.pool, k.O);
        float mx = -INFINITY;
        for (int yy = y0; yy < y1; ++yy)
          for (int xx = x0; xx < x1; ++xx) mx = fmaxf(mx, as[(m * k.O + yy) * k.O + xx]);
        p_img[t] = mx;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward-weights: dW[m,c,.,.] and db[m]; per-CTA partials in the layout conv_wgrad_finish reads
// ---------------------------------------------------------------------------------------------
// thread = (slice, mg, c, u): 4 maps x F filter columns in registers, sliding window along a row
// smem: xs[C][Sp][Wx] | gs[mP][O][O]   (aliased by the slice reduction at the end)
template <int F, bool GEN>
__global__ void __launch_bounds__(kFT) convpool_wgrad_kernel(const FusedArgs k) {
  extern __shared__ __align__(16) float sm[];
  const int G = (k.M + 3) >> 2, mP = 4 * G;
  const int Sp = k.O + F - 1;
  const int OO = k.O * k.O;
  float *xs = sm;
  float *gs = xs + k.C * Sp * Sp;
  const int T = G * k.C * F;                       // threads per slice
  const int nsl = max(1, min(kFT / T, k.O));       // row slices
  const int tid = threadIdx.x;
  const bool active = tid < T * nsl;
  const int slice = tid / T, r0 = tid - slice * T;
  const int u = r0 % F, c = (r0 / F) % k.C, mg = r0 / (F * k.C);
  const bool is_db = active && c == 0 && u == 0;
  float acc[4][F], dba[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int v = 0; v < F; ++v) acc[q][v] = 0.f;
  for (int t = tid; t < mP * OO; t += kFT) gs[t] = 0.f;  // padded maps / uncovered borders stay 0
  __syncthreads();

  const int PP = k.P * k.P;
  for (int b = blockIdx.x; b < k.B; b += gridDim.x) {
    stage_x(k, k.x + (size_t)b * k.C * k.S * k.S, xs, Sp, Sp);
    stage_gz<GEN>(k, k.a + (size_t)b * k.M * OO, k.pool ? k.pooled + (size_t)b * k.M * PP : nullptr,
             k.dtop + (size_t)b * k.M * (k.pool ? PP : OO), gs, k.O, k.O, 0);
    __syncthreads();
    if (active) {
      for (int i = slice; i < k.O; i += nsl) {
        const float *xr = xs + (c * Sp + i + u) * Sp;
        const float *g0 = gs + ((4 * mg) * k.O + i) * k.O;
        float xw[F];
#pragma unroll
        for (int v = 1; v < F; ++v) xw[v] = xr[v - 1];
        for (int j = 0; j < k.O; ++j) {
#pragma unroll
          for (int v = 0; v < F - 1; ++v) xw[v] = xw[v + 1];
          xw[F - 1] = xr[j + F - 1];
          float gq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) gq[q] = g0[q * OO + j];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int v = 0; v < F; ++v) acc[q][v] = fmaf(gq[q], xw[v], acc[q][v]);
          if (is_db) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dba[q] += gq[q];
          }
        }
      }
    }
    __syncthreads();
  }
  // combine the row slices in a fixed order; partial[blk][og*4+q], og = ((mg*C+c)*F+u)*F+v
  float *red = sm;  // [nsl][T][4*F]
  const int OG = G * k.C * F * F;
  if (active) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int v = 0; v < F; ++v) red[(slice * T + r0) * 4 * F + q * F + v] = acc[q][v];
  }
  __syncthreads();
  float *pout = k.partial + (size_t)blockIdx.x * (OG * 4 + mP);
  for (int t = tid; t < T * 4 * F; t += kFT) {
    float s = 0.f;
    for (int sl = 0; sl < nsl; ++sl) s += red[sl * T * 4 * F + t];
    const int rr = t / (4 * F), e = t - rr * 4 * F;   // rr = (mg, c, u), e = q*F + v
    const int q = e / F, v = e - q * F;
    pout[(rr * F + v) * 4 + q] = s;
  }
  __syncthreads();
  if (is_db) {
#pragma unroll
    for (int q = 0; q < 4; ++q) red[slice * mP + 4 * mg + q] = dba[q];
  }
  __syncthreads();
  if (tid < mP) {
    float s = 0.f;
    for (int sl = 0; sl < nsl; ++sl) s += red[sl * mP + tid];
    pout[OG * 4 + tid] = s;
  }
}

// final[o] = sum over CTAs in a fixed order; scatter into dW (OIHW, flipped back) and db
__global__ void __launch_bounds__(1024)
fused_wgrad_finish_kernel(const float *__restrict__ partial, int nblk, int C, int M, int f,
                          float *__restrict__ dW, float *__restrict__ db) {
  __shared__ float red[32][33];
  const int G = (M + 3) >> 2, mP = 4 * G;
  const int OG = G * C * f * f;
  const int stride = OG * 4 + mP;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (o < stride)
    for (int kk = ty; kk < nblk; kk += 32) s += partial[(size_t)kk * stride + o];
  red[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || o >= stride) return;
  s = red[0][tx];
#pragma unroll
  for (int q = 1; q < 32; ++q) s += red[q][tx];
  if (o < OG * 4) {
    const int q = o & 3, og = o >> 2;
    const int v = og % f;
    const int u = (og / f) % f;
    const int c = (og / (f * f)) % C;
    const int m = 4 * (og / (f * f * C)) + q;
    if (m < M) dW[((m * C + c) * f + (f - 1 - u)) * f + (f - 1 - v)] = s;
  } else {
    const int m = o - OG * 4;
    if (m < M) db[m] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// backward-data: dx[c,y,x] = sum_{m,u,v} gzpad[m,y+u,x+v] W[m,c,u,v]
// ---------------------------------------------------------------------------------------------
// thread = (map group mg of NG, channel group cg, row y, strip): partial over its maps, 4 input
// channels x 4 pixels in registers; partials combined through smem in a fixed order
// smem: ws[(m*F+u)*F+v][cP] | gs[M][Hp][Wg] zero-bordered | part[NG][cP][S][S]
template <int F, bool GEN>
__global__ void __launch_bounds__(kFT) convpool_dgrad_kernel(const FusedArgs k, int NG) {
  extern __shared__ __align__(16) float sm[];
  const int CG = (k.C + 3) >> 2, cP = 4 * CG;
  const int pd = F - 1 - k.pad_lo;                 // zero border of the gradient map
  const int strips = (k.S + kL - 1) / kL;
  const int Hp = k.S + F - 1;
  const int Wg = strips * kL + F - 1;
  const int OO = k.O * k.O, SS = k.S * k.S, PP = k.P * k.P;
  float *ws = sm;
  float *gs = ws + k.M * F * F * cP;
  float *part = gs + k.M * Hp * Wg;
  const int tid = threadIdx.x;
  for (int t = tid; t < k.M * F * F * cP; t += kFT) {
    const int co = t % cP;
    int r = t / cP;
    const int v = r % F; r /= F;
    const int u = r % F;
    const int m = r / F;
    ws[t] = co < k.C ? k.W[((m * k.C + co) * F + u) * F + v] : 0.f;
  }
  for (int t = tid; t < k.M * Hp * Wg; t += kFT) gs[t] = 0.f;  // borders stay zero
  __syncthreads();
  const int mper = (k.M + NG - 1) / NG;
  const int items = NG * CG * k.S * strips;

  for (int b = blockIdx.x; b < k.B; b += gridDim.x) {
    stage_gz<GEN>(k, k.a + (size_t)b * k.M * OO, k.pool ? k.pooled + (size_t)b * k.M * PP : nullptr,
             k.dtop + (size_t)b * k.M * (k.pool ? PP : OO), gs, Hp, Wg, pd);
    __syncthreads();
    for (int it = tid; it < items; it += kFT) {
      int r = (int)k.divstrS.div(it);
      const int s = it - r * strips;
      const int r2 = (int)k.divS.div(r);
      const int y = r - r2 * k.S;
      const int cg = r2 % CG;
      const int mg = r2 / CG;
      const int x0 = s * kL;
      float acc[kL][4];
#pragma unroll
      for (int l = 0; l < kL; ++l)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
      const int m1 = min(k.M, (mg + 1) * mper);
      for (int m = mg * mper; m < m1; ++m) {
#pragma unroll
        for (int u = 0; u < F; ++u) {
          const float *gr = gs + (m * Hp + y + u) * Wg + x0;
          float gv[kL + F - 1];
#pragma unroll
          for (int e = 0; e < kL + F - 1; ++e) gv[e] = gr[e];
          const float4 *wr = reinterpret_cast<const float4 *>(ws + ((m * F + u) * F) * cP) + cg;
#pragma unroll
          for (int v = 0; v < F; ++v) {
            const float4 w = wr[v * CG];
#pragma unroll
            for (int l = 0; l < kL; ++l) {
              acc[l][0] = fmaf(gv[l + v], w.x, acc[l][0]);
              acc[l][1] = fmaf(gv[l + v], w.y, acc[l][1]);
              acc[l][2] = fmaf(gv[l + v], w.z, acc[l][2]);
              acc[l][3] = fmaf(gv[l + v], w.w, acc[l][3]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int l = 0; l < kL; ++l)
          if (x0 + l < k.S) part[((mg * cP + 4 * cg + q) * k.S + y) * k.S + x0 + l] = acc[l][q];
    }
    __syncthreads();
    float *dx_img = k.dx + (size_t)b * k.C * SS;
    const float *bl_img = k.below ? k.below + (size_t)b * k.C * SS : nullptr;
    for (int t = tid; t < k.C * SS; t += kFT) {
      float s = 0.f;
      for (int mg = 0; mg < NG; ++mg) s += part[mg * cP * SS + t];
      if (bl_img) s *= act_bwd_t<GEN>(k.akb, bl_img[t]);  // rare: conv on conv
      dx_img[t] = s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
static int images_grid(int B) {
  const int cap = 4 * kNumSM;
  const int per = ceil_div(B, cap);
  return ceil_div(B, per);
}

template <typename K>
static int set_smem(K kernel, size_t smem, const char *who) {
  TN_REQUIRE(smem <= 220 * 1024, TN_ERR_UNSUPPORTED,
             "%s: fused small-channel path needs %zu B of shared memory; use the general path", who,
             smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
  }
  return TN_OK;
}

static void set_act(FusedArgs &k, int act, int act_nn) {
  k.act = act;
  k.nn = (float)act_nn;
  k.ak = make_actk(act, act_nn);
}

static int fill_geom(FusedArgs &k, int B, int C, int S, int M, int f, int pad_lo, int O, int pool,
                     int P, const char *who) {
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0 && O > 0 && pad_lo >= 0 && pad_lo <= f - 1 &&
                 O + f - 1 >= S + pad_lo && S < 32768,
             TN_ERR_SHAPE, "%s: bad shape B=%d C=%d S=%d M=%d f=%d pad=%d out=%d", who, B, C, S, M,
             f, pad_lo, O);
  TN_REQUIRE(f == 3 || f == 5, TN_ERR_UNSUPPORTED, "%s: fused path supports filter_sz 3 and 5 (got %d)",
             who, f);
  TN_REQUIRE(pool == 0 || (pool > 0 && P > 0 && P * pool < O + pool), TN_ERR_SHAPE,
             "%s: bad pool geometry O=%d p=%d P=%d", who, O, pool, P);
  k.B = B; k.C = C; k.S = S; k.M = M; k.f = f; k.pad_lo = pad_lo; k.O = O; k.P = pool ? P : 0;
  k.pool = pool;
  k.divO = FastDiv(O); k.divOO = FastDiv(O * O); k.divS = FastDiv(S);
  k.divP = FastDiv(pool ? P : 1); k.divPP = FastDiv(pool ? P * P : 1);
  k.divSp = FastDiv(O + f - 1);
  k.divstrO = FastDiv(ceil_div(O, kL)); k.divstrS = FastDiv(ceil_div(S, kL));
  TN_REQUIRE((int64_t)M * O * O < 65536 && (int64_t)C * S * S < 65536, TN_ERR_UNSUPPORTED,
             "%s: image too large for the fused small-channel path", who);
  return TN_OK;
}

}  // namespace tn

using namespace tn;

extern "C" int tn_convpool_fprop(const float *x, const float *W, const float *bias, float *a,
                                 float *pooled, int B, int C, int S, int M, int f, int pad_lo,
                                 int out_sz, int act, int act_nn, int pool, int pool_out_sz,
                                 void *stream) {
  const char *who = "tn_convpool_fprop";
  TN_REQUIRE(x && W && bias && a && (pooled || !pool), TN_ERR_ARG, "%s: null argument", who);
  if (B > 0 && small_conv_ok(C, S, M, f, pad_lo, out_sz, act, pool, pool_out_sz))   // conv_small.cu
    return small_fprop(x, W, bias, a, pooled, nullptr, B, C, S, M, out_sz, act, act_nn, pool_out_sz,
                       (cudaStream_t)stream);
  FusedArgs k{};
  int rc = fill_geom(k, B, C, S, M, f, pad_lo, out_sz, pool, pool_out_sz, who);
  if (rc) return rc;
  k.x = x; k.W = W; k.bias = bias; k.a = a; k.pooled = pooled;
  set_act(k, act, act_nn);
  const int G = (M + 3) / 4, strips = ceil_div(out_sz, kL);
  const size_t smem = ((size_t)C * f * f * 4 * G + (size_t)C * (out_sz + f - 1) * (strips * kL + f - 1) + 3 +
                       (size_t)M * out_sz * out_sz) * sizeof(float);
  const bool gen = !act_is_fast(act);
  void (*kern)(FusedArgs) = f == 3 ? (gen ? convpool_fprop_kernel<3, true> : convpool_fprop_kernel<3, false>)
                                    : (gen ? convpool_fprop_kernel<5, true> : convpool_fprop_kernel<5, false>);
  rc = set_smem(kern, smem, who);
  if (rc) return rc;
  kern<<<images_grid(B), kFT, smem, (cudaStream_t)stream>>>(k);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}

extern "C" size_t tn_convpool_bwd_weights_workspace_bytes(int B, int C, int M, int f) {
  const int G = (M + 3) / 4;
  return (size_t)images_grid(B) * ((size_t)G * C * f * f * 4 + 4 * G) * sizeof(float);
}

extern "C" int tn_convpool_bwd_weights(const float *x, const float *a, const float *pooled,
                                       const float *dtop, float *dW, float *db, void *workspace,
                                       int B, int C, int S, int M, int f, int pad_lo, int out_sz,
                                       int act, int act_nn, int pool, int pool_out_sz,
                                       void *stream) {
  const char *who = "tn_convpool_bwd_weights";
  TN_REQUIRE(x && a && dtop && dW && db && workspace && (pooled || !pool), TN_ERR_ARG,
             "%s: null argument", who);
  FusedArgs k{};
  int rc = fill_geom(k, B, C, S, M, f, pad_lo, out_sz, pool, pool_out_sz, who);
  if (rc) return rc;
  const int G = (M + 3) / 4, mP = 4 * G;
  TN_REQUIRE(G * C * f <= kFT, TN_ERR_UNSUPPORTED,
             "%s: ceil(M/4)*C*f = %d exceeds %d threads; use the general path", who, G * C * f, kFT);
  k.x = x; k.a = const_cast<float *>(a); k.pooled = const_cast<float *>(pooled); k.dtop = dtop;
  k.partial = (float *)workspace;
  set_act(k, act, act_nn);
  const int Sp = out_sz + f - 1;
  size_t smem = ((size_t)C * Sp * Sp + (size_t)mP * out_sz * out_sz) * sizeof(float);
  const int T = G * C * f, nsl = std::max(1, std::min(kFT / T, out_sz));
  smem = std::max(smem, (size_t)nsl * T * 4 * f * sizeof(float));
  const bool gen = !act_is_fast(act);
  void (*kern)(FusedArgs) = f == 3 ? (gen ? convpool_wgrad_kernel<3, true> : convpool_wgrad_kernel<3, false>)
                                    : (gen ? convpool_wgrad_kernel<5, true> : convpool_wgrad_kernel<5, false>);
  rc = set_smem(kern, smem, who);
  if (rc) return rc;
  const int grid = images_grid(B);
  cudaStream_t st = (cudaStream_t)stream;
  kern<<<grid, kFT, smem, st>>>(k);
  TN_LAUNCH_CHECK(who);
  const int n = G * C * f * f * 4 + mP;
  fused_wgrad_finish_kernel<<<ceil_div(n, 32), 1024, 0, st>>>((const float *)workspace, grid, C, M,
                                                               f, dW, db);
  TN_LAUNCH_CHECK("tn_convpool_bwd_weights(finish)");
  return TN_OK;
}

extern "C" int tn_convpool_bwd_data(const float *a, const float *pooled, const float *dtop,
                                    const float *W, float *dx, const float *below, int B, int C,
                                    int S, int M, int f, int pad_lo, int out_sz, int act,
                                    int act_nn, int pool, int pool_out_sz, int act_below,
                                    int nn_below, void *stream) {
  const char *who = "tn_convpool_bwd_data";
  TN_REQUIRE(a && dtop && W && dx && (pooled || !pool), TN_ERR_ARG, "%s: null argument", who);
  FusedArgs k{};
  int rc = fill_geom(k, B, C, S, M, f, pad_lo, out_sz, pool, pool_out_sz, who);
  if (rc) return rc;
  k.a = const_cast<float *>(a); k.pooled = const_cast<float *>(pooled); k.dtop = dtop; k.W = W;
  k.dx = dx; k.below = below; k.act_below = act_below;
  k.akb = make_actk(act_below, nn_below);
  set_act(k, act, act_nn);
  k.nn_below = (float)nn_below;
  const int CG = (C + 3) / 4, cP = 4 * CG, strips = ceil_div(S, kL);
  int NG = kFT / std::max(1, CG * S * strips);
  NG = std::max(1, std::min(NG, M));
  const size_t smem = ((size_t)M * f * f * cP + (size_t)M * (S + f - 1) * (strips * kL + f - 1) +
                       (size_t)NG * cP * S * S) * sizeof(float);
  const bool gen = !act_is_fast(act) || (below && !act_is_fast(act_below));
  void (*kern)(FusedArgs, int) = f == 3 ? (gen ? convpool_dgrad_kernel<3, true> : convpool_dgrad_kernel<3, false>)
                                         : (gen ? convpool_dgrad_kernel<5, true> : convpool_dgrad_kernel<5, false>);
  rc = set_smem(kern, smem, who);
  if (rc) return rc;
  kern<<<images_grid(B), kFT, smem, (cudaStream_t)stream>>>(k, NG);
  TN_LAUNCH_CHECK(who);
  return TN_OK;
}
