// Library plumbing (version, thread-local error text, device gate) and the small elementwise
// entry points: Philox word dump, standalone dropout (theanet/layer/dropout.py:9-31) and the
// activation backward used when a conv layer feeds a dense layer directly.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace tn {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("TN_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void philox_words_kernel(uint32_t *__restrict__ words, int n_samples, int n,
                                    uint64_t seed, int purpose, uint32_t step, uint32_t sample0) {
  const int nblk = (n + 3) >> 2;
  const int64_t total = (int64_t)n_samples * nblk;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(t / nblk), blk = (int)(t % nblk);
    const Philox4 r = philox_block(seed, purpose, step, sample0 + (uint32_t)s, (uint32_t)blk);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * blk + q < n) words[(int64_t)s * n + 4 * blk + q] = w[q];
  }
}

__global__ void dropout_apply_kernel(const float *x, float *out, int B,  // may alias (in place)
                                     int n, int mask_on, uint32_t thr, uint64_t seed,
                                     const int32_t *__restrict__ ctl,
                                     const float *__restrict__ mask_inj, float scale) {
  const int nblk = (n + 3) >> 2;
  const int64_t total = (int64_t)B * nblk;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / nblk), blk = (int)(t % nblk);
    float mk[4] = {1.f, 1.f, 1.f, 1.f};
    if (mask_on == 1) {
      const Philox4 r = philox_block(seed, TN_RNG_DROPOUT, (uint32_t)ctl[TN_CTL_STEP],
                                     (uint32_t)(ctl[TN_CTL_SAMPLE0] + b), (uint32_t)blk);
      mk[0] = r.x < thr ? 1.f : 0.f;
      mk[1] = r.y < thr ? 1.f : 0.f;
      mk[2] = r.z < thr ? 1.f : 0.f;
      mk[3] = r.w < thr ? 1.f : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = 4 * blk + q;
      if (j < n) {
        const int64_t i = (int64_t)b * n + j;
        float v = x[i];
        if (mask_on == 2) v *= mask_inj[i];
        else if (mask_on == 1) v *= mk[q];
        if (scale != 1.f) v *= scale;
        out[i] = v;
      }
    }
  }
}

__global__ void act_bwd_kernel(const float *g, const float *__restrict__ a,
                               float *gz,  /* g and gz may alias */ int64_t n, int act, float nn) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    gz[i] = g[i] * act_bwd_from_out(a[i], act, nn);
}

__global__ void set_ctl_kernel(int32_t *ctl, int step, int sample0, int row0, int lr_bits) {
  if (threadIdx.x == 0) {
    ctl[TN_CTL_STEP] = step;
    ctl[TN_CTL_SAMPLE0] = sample0;
    ctl[TN_CTL_ROW0] = row0;
    ctl[TN_CTL_LR_BITS] = lr_bits;
  }
}

}  // namespace tn

using namespace tn;

// The per-step scalars travel as kernel arguments: they are copied when the launch is enqueued, so
// a host that runs several steps ahead of the device (lazy returns) can never overwrite values a
// queued step has yet to read -- which a pinned staging buffer read by an in-graph memcpy could.
extern "C" int tn_set_ctl(int32_t *ctl, int step, int sample0, int row0, int lr_bits,
                          void *stream) {
  TN_REQUIRE(ctl, TN_ERR_ARG, "tn_set_ctl: null control block");
  set_ctl_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ctl, step, sample0, row0, lr_bits);
  TN_LAUNCH_CHECK("tn_set_ctl");
  return TN_OK;
}

extern "C" int tn_version(void) { return TN_VERSION; }

extern "C" const char *tn_last_error(void) { return g_err; }

extern "C" uint64_t tn_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int tn_device_check(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  TN_REQUIRE(e == cudaSuccess, TN_ERR_CUDA, "tn_device_check: %s", cudaGetErrorString(e));
  TN_REQUIRE(p.major == 10, TN_ERR_UNSUPPORTED,
             "tn_device_check: device %d is sm_%d%d; this library is built for sm_100a only",
             device, p.major, p.minor);
  return TN_OK;
}

extern "C" int tn_philox_words(uint32_t *words, int n_samples, int n_per_sample, uint64_t seed,
                               int purpose, int step, int sample0, void *stream) {
  TN_REQUIRE(words && n_samples > 0 && n_per_sample > 0, TN_ERR_ARG, "tn_philox_words: bad argument");
  const int64_t total = (int64_t)n_samples * ((n_per_sample + 3) / 4);
  const int blocks = (int)min64(ceil_div64(total, 256), (int64_t)kNumSM * 16);
  philox_words_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      words, n_samples, n_per_sample, seed, purpose, (uint32_t)step, (uint32_t)sample0);
  TN_LAUNCH_CHECK("tn_philox_words");
  return TN_OK;
}

extern "C" int tn_dropout_apply(const float *x, float *out, int B, int n, double pkeep,
                                uint64_t seed, const int32_t *ctl, const float *mask_inj,
                                float scale, void *stream) {
  TN_REQUIRE(x && out && B > 0 && n > 0, TN_ERR_ARG, "tn_dropout_apply: bad argument");
  const int mask_on = mask_inj ? 2 : (pkeep < 1.0 ? 1 : 0);
  TN_REQUIRE(mask_on != 1 || ctl, TN_ERR_ARG, "tn_dropout_apply: dropout needs ctl");
  const int64_t total = (int64_t)B * ((n + 3) / 4);
  const int blocks = (int)min64(ceil_div64(total, 256), (int64_t)kNumSM * 16);
  dropout_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      x, out, B, n, mask_on, bernoulli_threshold(pkeep), seed, ctl, mask_inj, scale);
  TN_LAUNCH_CHECK("tn_dropout_apply");
  return TN_OK;
}

extern "C" int tn_act_bwd(const float *g, const float *a, float *gz, int64_t n, int act,
                          int act_nn, void *stream) {
  TN_REQUIRE(g && a && gz && n > 0, TN_ERR_ARG, "tn_act_bwd: bad argument");
  const int blocks = (int)min64(ceil_div64(n, 256), (int64_t)kNumSM * 16);
  act_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, a, gz, n, act, (float)act_nn);
  TN_LAUNCH_CHECK("tn_act_bwd");
  return TN_OK;
}
