// ColorLayer (theanet/layer/color.py:9-52): per (sample, map) random white balance and two gamma
// curves.  out = maxval * (1 - (1 - clip(x/maxval * e1, 0, 1)^e2)^e3) with
// e1 = exp(ln(balance) * u1), e2 = exp(ln(gamma) * u2), e3 = exp(ln(gamma) * u3), u ~ U(-1,1)
// drawn per (sample, map) (color.py:32-42; pos_rand is called three times = three draws).
// The draws come from the (seed, step, global sample) Philox stream, block = map, words x, y, z,
// or from u_inj (B*C*3 float32, already in (-1,1)).  Pure HBM: one read, one write per pixel.
#include "common.cuh"

namespace tn {

__device__ __forceinline__ float sym_uniform(uint32_t w) {   // oracle/philox.py: 2*uniform01(w) - 1
  return (float)(2.0 * (((double)w + 0.5) * (1.0 / 4294967296.0)) - 1.0);
}

__global__ void color_jitter_kernel(const float *__restrict__ x, float *__restrict__ out, int B,
                                    int C, int hw, float log_balance, float log_gamma,
                                    float maxval, uint64_t seed, const int32_t *__restrict__ ctl,
                                    const float *__restrict__ u_inj) {
  // one CTA per (sample, map) plane: the three exponents are uniform over the block
  for (int p = blockIdx.x; p < B * C; p += gridDim.x) {
    const int b = p / C, c = p - b * C;
    float u1, u2, u3;
    if (u_inj) {
      u1 = u_inj[3 * p], u2 = u_inj[3 * p + 1], u3 = u_inj[3 * p + 2];
    } else {
      const Philox4 r = philox_block(seed, TN_RNG_COLOR, (uint32_t)ctl[TN_CTL_STEP],
                                     (uint32_t)(ctl[TN_CTL_SAMPLE0] + b), (uint32_t)c);
      u1 = sym_uniform(r.x), u2 = sym_uniform(r.y), u3 = sym_uniform(r.z);
    }
    const float e1 = expf(__fmul_rn(log_balance, u1));
    const float e2 = expf(__fmul_rn(log_gamma, u2));
    const float e3 = expf(__fmul_rn(log_gamma, u3));
    const float *src = x + (size_t)p * hw;
    float *dst = out + (size_t)p * hw;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {
      float v = __fmul_rn(__fdiv_rn(src[i], maxval), e1);
      v = fminf(fmaxf(v, 0.f), 1.f);
      v = powf(v, e2);
      v = __fsub_rn(1.f, powf(__fsub_rn(1.f, v), e3));
      dst[i] = __fmul_rn(v, maxval);
    }
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_color_jitter(const float *x, float *out, int B, int C, int S, float log_balance,
                               float log_gamma, float maxval, uint64_t seed, const int32_t *ctl,
                               const float *u_inj, void *stream) {
  TN_REQUIRE(x && out && (ctl || u_inj), TN_ERR_ARG, "tn_color_jitter: null argument");
  TN_REQUIRE(B > 0 && C > 0 && S > 0, TN_ERR_SHAPE, "tn_color_jitter: bad shape");
  TN_REQUIRE(maxval > 0.f, TN_ERR_ARG, "tn_color_jitter: maxval must be positive");
  const int hw = S * S;
  const int threads = hw >= 1024 ? 256 : (hw >= 256 ? 128 : 64);
  const int blocks = (int)min64((int64_t)B * C, (int64_t)kNumSM * 32);
  color_jitter_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
      x, out, B, C, hw, log_balance, log_gamma, maxval, seed, ctl, u_inj);
  TN_LAUNCH_CHECK("tn_color_jitter");
  return TN_OK;
}
