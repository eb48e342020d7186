// Narrow-output dense kernels (dense_small.cu), dispatched from the tn_dense_* entry points.
#pragma once
#include "common.cuh"

namespace tn {

constexpr int kSmallN = 32;

struct SmallArgs {
  const float *x;      // fwd: (B, n_in);           bwd-data: g (B, n_out)
  const float *W;      // (n_in, n_out)
  const float *bias;   // fwd
  const float *aux;    // bwd-data: prev_out (B, n_in) or null
  const float *mask_inj;
  float *out;          // fwd: (B, n_out);          bwd-data: dx (B, n_in)
  const int32_t *ctl;
  uint64_t seed;
  uint32_t thr;
  int mask_on, act;
  float act_nn, scale;
  int B, n_in, n_out;
};

bool dense_small_ok(int n_in, int n_out);
int dense_fwd_small(const SmallArgs &a, cudaStream_t st);
int dense_bwd_data_small(const SmallArgs &a, cudaStream_t st);
int dense_bwd_weights_small(const float *x, const float *g, float *dW, float *db, int B, int n_in,
                            int n_out, cudaStream_t st);

}  // namespace tn
