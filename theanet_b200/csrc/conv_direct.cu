// ConvLayer kernels, small-channel direct path (theanet/layer/convpool.py:42-72; Theano conv2d
// with filter_flip=True, stride 1).  For the shapes theanet's shipped networks use (C*f*f <= 36,
// M <= 20: params/mnist.prms:13-27) the GEMM view has N=4..20 and K=9..36, far below any tcgen05
// tile, and the layers are HBM-bound: one CTA stages a whole (padded) image and the
// re-laid-out filter bank in shared memory, every thread produces 4 output channels of one pixel
// from one scalar + one broadcast 128-bit shared load per tap, and HBM sees each activation once.
//
//   fprop : out = act(bias + xpad (*) flip(W))
//   dgrad : the same correlation with the roles of the channel axes swapped and an unflipped W
//   wgrad : per-CTA register accumulation over a grid-strided set of images, then a fixed-order
//           cross-CTA reduction (deterministic, no atomics)
#include "common.cuh"

namespace tn {

constexpr int kConvThreads = 256;

struct CorrArgs {
  const float *in;      // (B, Cin, S_in, S_in)
  const float *W;       // OIHW (M, C, f, f) of the conv layer
  const float *bias;    // fprop only
  const float *mul_in;  // dgrad only: tensor whose act' multiplies the result (or null)
  float *out;           // (B, Cout, S_out, S_out)
  int B, Cin, S_in, Cout, S_out, f, pad_lo;
  int M, C;             // layer's filter dims (for indexing W)
  int dgrad;            // 0: fprop (flip), 1: dgrad (no flip, channel axes swapped)
  int act;
  float act_nn;
};

// ws[(ci*f+u)*f+v][co] (co padded to a multiple of 4); in_s[ci][Y][X] zero padded, Sp = S_out+f-1
template <int F>
__global__ void __launch_bounds__(kConvThreads) conv_corr_kernel(CorrArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int f = F > 0 ? F : a.f;
  const int Sp = a.S_out + f - 1;
  const int G = (a.Cout + 3) >> 2, coP = G * 4;
  const int nW = a.Cin * f * f * coP;
  float *ws = smem;
  float *in_s = smem + nW;
  const int b = blockIdx.x;

  for (int t = threadIdx.x; t < nW; t += blockDim.x) {
    const int co = t % coP;
    int r = t / coP;
    const int v = r % f; r /= f;
    const int u = r % f;
    const int ci = r / f;
    float w = 0.f;
    if (co < a.Cout) {
      if (!a.dgrad)  // ci = c, co = m : W[m, c, f-1-u, f-1-v]
        w = a.W[((co * a.C + ci) * f + (f - 1 - u)) * f + (f - 1 - v)];
      else           // ci = m, co = c : W[m, c, u, v]
        w = a.W[((ci * a.C + co) * f + u) * f + v];
    }
    ws[t] = w;
  }
  const float *img = a.in + (size_t)b * a.Cin * a.S_in * a.S_in;
  const int nIn = a.Cin * Sp * Sp;
  for (int t = threadIdx.x; t < nIn; t += blockDim.x) {
    const int X = t % Sp;
    int r = t / Sp;
    const int Y = r % Sp;
    const int ci = r / Sp;
    const int y = Y - a.pad_lo, x = X - a.pad_lo;
    float val = 0.f;
    if (y >= 0 && y < a.S_in && x >= 0 && x < a.S_in) val = img[(ci * a.S_in + y) * a.S_in + x];
    in_s[t] = val;
  }
  __syncthreads();

  const int npix = a.S_out * a.S_out;
  const int items = npix * G;
  const float4 *ws4 = reinterpret_cast<const float4 *>(ws);
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int g = it / npix, pix = it - g * npix;
    const int i = pix / a.S_out, j = pix - i * a.S_out;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ci = 0; ci < a.Cin; ++ci) {
      const float *ip = in_s + (ci * Sp + i) * Sp + j;
      const float4 *wp = ws4 + (ci * f * f) * G + g;
#pragma unroll
      for (int u = 0; u < f; ++u) {
#pragma unroll
        for (int v = 0; v < f; ++v) {
          const float x = ip[u * Sp + v];
          const float4 w = wp[(u * f + v) * G];
          acc.x = fmaf(x, w.x, acc.x);
          acc.y = fmaf(x, w.y, acc.y);
          acc.z = fmaf(x, w.z, acc.z);
          acc.w = fmaf(x, w.w, acc.w);
        }
      }
    }
    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = 4 * g + q;
      if (co < a.Cout) {
        const size_t o = ((size_t)b * a.Cout + co) * npix + pix;
        float val = r[q];
        if (!a.dgrad) {
          val = act_fwd(val + a.bias[co], a.act, a.act_nn);
        } else if (a.mul_in) {
          val *= act_bwd_from_out(a.mul_in[o], a.act, a.act_nn);
        }
        a.out[o] = val;
      }
    }
  }
}

static int launch_corr(const CorrArgs &a, const char *name, cudaStream_t st) {
  const int Sp = a.S_out + a.f - 1;
  const int coP = ((a.Cout + 3) / 4) * 4;
  const size_t smem = ((size_t)a.Cin * a.f * a.f * coP + (size_t)a.Cin * Sp * Sp) * sizeof(float);
  TN_REQUIRE(smem <= 220 * 1024, TN_ERR_UNSUPPORTED,
             "%s: direct path needs %zu B of shared memory (Cin=%d S=%d f=%d Cout=%d); "
             "use the implicit-GEMM path",
             name, smem, a.Cin, a.S_in, a.f, a.Cout);
  void (*k)(CorrArgs) = a.f == 3 ? conv_corr_kernel<3> : (a.f == 5 ? conv_corr_kernel<5> : conv_corr_kernel<0>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  }
  k<<<a.B, kConvThreads, smem, st>>>(a);
  TN_LAUNCH_CHECK(name);
  return TN_OK;
}

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const float *x;   // (B, C, S, S) layer input
  const float *gz;  // (B, M, out, out) dL/dz
  float *partial;   // [grid][OG*4 + mP]
  int B, C, S, M, f, pad_lo, out_sz;
  int RB;           // output rows staged per pass (a band of the image)
};

template <int F>
__global__ void __launch_bounds__(1024) conv_wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int f = F > 0 ? F : a.f;
  const int Sp = a.out_sz + f - 1;
  const int G = (a.M + 3) >> 2, mP = 4 * G;
  const int npix = a.out_sz * a.out_sz;
  const int OG = G * a.C * f * f;
  const int PS = max(1, (int)blockDim.x / OG);
  const int RB = a.RB;
  float *gs = smem;                               // [RB*out_sz][mP]
  float *xs = smem + (size_t)RB * a.out_sz * mP;  // [C][RB+f-1][Sp]
  const int tid = threadIdx.x;
  const bool active = tid < OG * PS;
  const int slice = tid / OG, og = tid % OG;
  // og -> (mg, c, u, v)
  const int v = og % f;
  const int u = (og / f) % f;
  const int c = (og / (f * f)) % a.C;
  const int mg = og / (f * f * a.C);
  const bool is_db = active && c == 0 && u == 0 && v == 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), dba = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 *gs4 = reinterpret_cast<const float4 *>(gs);

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float *gimg = a.gz + (size_t)b * a.M * npix;
    const float *img = a.x + (size_t)b * a.C * a.S * a.S;
    for (int r0 = 0; r0 < a.out_sz; r0 += RB) {   // bands of output rows: large images fit in smem
      const int nr = min(RB, a.out_sz - r0);
      const int bpix = nr * a.out_sz, xr = nr + f - 1;
      for (int t = tid; t < bpix * mP; t += blockDim.x) {
        const int m = t / bpix, p = t - m * bpix;  // coalesced global read, transposed shared store
        gs[p * mP + m] = m < a.M ? gimg[m * npix + r0 * a.out_sz + p] : 0.f;
      }
      for (int t = tid; t < a.C * xr * Sp; t += blockDim.x) {
        const int X = t % Sp;
        int r = t / Sp;
        const int Y = r % xr;
        const int ci = r / xr;
        const int y = r0 + Y - a.pad_lo, x = X - a.pad_lo;
        xs[t] = (y >= 0 && y < a.S && x >= 0 && x < a.S) ? img[(ci * a.S + y) * a.S + x] : 0.f;
      }
      __syncthreads();
      if (active) {
        const float *xp = xs + (c * xr + u) * Sp + v;
        for (int p = slice; p < bpix; p += PS) {
          const int i = p / a.out_sz, j = p - i * a.out_sz;
          const float xv = xp[i * Sp + j];
          const float4 g = gs4[p * G + mg];
          acc.x = fmaf(xv, g.x, acc.x);
          acc.y = fmaf(xv, g.y, acc.y);
          acc.z = fmaf(xv, g.z, acc.z);
          acc.w = fmaf(xv, g.w, acc.w);
          if (is_db) {
            dba.x += g.x; dba.y += g.y; dba.z += g.z; dba.w += g.w;
          }
        }
      }
      __syncthreads();
    }
  }
  // reduce the PS pixel slices in a fixed order (shared memory reused)
  float4 *red = reinterpret_cast<float4 *>(smem);  // [PS][OG] then [PS][G]
  if (active) red[slice * OG + og] = acc;
  __syncthreads();
  float *pout = a.partial + (size_t)blockIdx.x * (OG * 4 + mP);
  if (tid < OG) {
    float4 s = red[tid];
    for (int q = 1; q < PS; ++q) {
      const float4 t = red[q * OG + tid];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4 *>(pout)[tid] = s;
  }
  __syncthreads();
  if (is_db) red[slice * G + mg] = dba;
  __syncthreads();
  if (tid < G) {
    float4 s = red[tid];
    for (int q = 1; q < PS; ++q) {
      const float4 t = red[q * G + tid];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4 *>(pout + OG * 4)[tid] = s;
  }
}

// final[o] = sum over CTAs (fixed order); scatter into dW (OIHW, flipped back) and db.
// 32 outputs x 32 slices of the CTA partials per block, slices combined in a fixed order.
__global__ void __launch_bounds__(1024)
conv_wgrad_finish_kernel(const float *__restrict__ partial, int nblk, int C, int M, int f,
                         float *__restrict__ dW, float *__restrict__ db) {
  __shared__ float red[32][33];
  const int G = (M + 3) >> 2, mP = 4 * G;
  const int OG = G * C * f * f;
  const int stride = OG * 4 + mP;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (o < stride)
    for (int k = ty; k < nblk; k += 32) s += partial[(size_t)k * stride + o];
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

static int wgrad_grid(int B) { return min(B, 2 * kNumSM); }

}  // namespace tn

using namespace tn;

extern "C" int tn_conv2d_fprop(const float *x, const float *W, const float *bias, float *out,
                               int B, int C, int S, int M, int f, int pad_lo, int out_sz, int act,
                               int act_nn, void *stream) {
  TN_REQUIRE(x && W && bias && out, TN_ERR_ARG, "tn_conv2d_fprop: null argument");
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0 && f > 0 && out_sz > 0 && pad_lo >= 0 &&
                 out_sz + f - 1 >= S + pad_lo,
             TN_ERR_SHAPE, "tn_conv2d_fprop: bad shape B=%d C=%d S=%d M=%d f=%d pad=%d out=%d", B,
             C, S, M, f, pad_lo, out_sz);
  CorrArgs a{x, W, bias, nullptr, out, B, C, S, M, out_sz, f, pad_lo, M, C, 0, act, (float)act_nn};
  return launch_corr(a, "tn_conv2d_fprop", (cudaStream_t)stream);
}

extern "C" int tn_conv2d_dgrad(const float *gz, const float *W, float *dx, const float *x_in,
                               int B, int C, int S, int M, int f, int pad_lo, int out_sz,
                               int act_prev, int nn_prev, void *stream) {
  TN_REQUIRE(gz && W && dx, TN_ERR_ARG, "tn_conv2d_dgrad: null argument");
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0 && f > 0 && out_sz > 0 && pad_lo >= 0 &&
                 pad_lo <= f - 1,
             TN_ERR_SHAPE, "tn_conv2d_dgrad: bad shape");
  // dx[y] = sum_{u'} gzpad[y+u'] W[u'],  gzpad padded by f-1-pad_lo  (see DESIGN.md "conv dgrad")
  CorrArgs a{gz, W, nullptr, x_in, dx, B, M, out_sz, C, S, f, f - 1 - pad_lo, M, C, 1, act_prev,
             (float)nn_prev};
  return launch_corr(a, "tn_conv2d_dgrad", (cudaStream_t)stream);
}

extern "C" size_t tn_conv2d_wgrad_workspace_bytes(int B, int C, int S, int M, int f) {
  (void)S;
  const int G = (M + 3) / 4;
  return (size_t)wgrad_grid(B) * ((size_t)G * C * f * f * 4 + 4 * G) * sizeof(float);
}

extern "C" int tn_conv2d_wgrad(const float *x, const float *gz, float *dW, float *db,
                               void *workspace, int B, int C, int S, int M, int f, int pad_lo,
                               int out_sz, void *stream) {
  TN_REQUIRE(x && gz && dW && db && workspace, TN_ERR_ARG, "tn_conv2d_wgrad: null argument");
  TN_REQUIRE(B > 0 && C > 0 && S > 0 && M > 0 && f > 0 && out_sz > 0, TN_ERR_SHAPE,
             "tn_conv2d_wgrad: bad shape");
  const int G = (M + 3) / 4, mP = 4 * G;
  const int OG = G * C * f * f;
  TN_REQUIRE(OG <= 1024, TN_ERR_UNSUPPORTED,
             "tn_conv2d_wgrad: direct path supports ceil(M/4)*C*f*f <= 1024 (got %d); use the "
             "implicit-GEMM path", OG);
  const int PS = max(1, kConvThreads / OG);
  const int threads = ((OG * PS + 31) / 32) * 32;
  const int Sp = out_sz + f - 1;
  // rows per band: the whole image when it fits in ~96 KB, else as many rows as do
  int RB = out_sz;
  auto band_bytes = [&](int rb) {
    return ((size_t)rb * out_sz * mP + (size_t)C * (rb + f - 1) * Sp) * sizeof(float);
  };
  while (RB > 1 && band_bytes(RB) > 96 * 1024) RB = (RB + 1) / 2;
  size_t smem = band_bytes(RB);
  const size_t red = (size_t)PS * OG * 16;
  if (red > smem) smem = red;
  TN_REQUIRE(smem <= 220 * 1024, TN_ERR_UNSUPPORTED,
             "tn_conv2d_wgrad: direct path needs %zu B of shared memory", smem);
  void (*k)(WgradArgs) = f == 3 ? conv_wgrad_kernel<3> : (f == 5 ? conv_wgrad_kernel<5> : conv_wgrad_kernel<0>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_conv2d_wgrad: %s", cudaGetErrorString(e));
  }
  const int grid = wgrad_grid(B);
  WgradArgs a{x, gz, (float *)workspace, B, C, S, M, f, pad_lo, out_sz, RB};
  cudaStream_t st = (cudaStream_t)stream;
  k<<<grid, threads, smem, st>>>(a);
  TN_LAUNCH_CHECK("tn_conv2d_wgrad");
  const int n = OG * 4 + mP;
  conv_wgrad_finish_kernel<<<ceil_div(n, 32), 1024, 0, st>>>((const float *)workspace, grid, C, M,
                                                              f, dW, db);
  TN_LAUNCH_CHECK("tn_conv2d_wgrad(finish)");
  return TN_OK;
}
