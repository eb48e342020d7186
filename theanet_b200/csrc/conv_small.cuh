// Second-generation small-channel ConvLayer + 2x2 PoolLayer kernels (conv_small.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tn {

// filter 3x3, mode 'valid', 2x2 pool, ReLU-family / linear activation, everything of one image
// resident in shared memory
bool small_conv_ok(int C, int S, int M, int f, int pad_lo, int O, int act, int pool, int P);

// a == NULL: the un-pooled activations are not written; tie != NULL: per pooled cell, which of
// the window's elements equal the maximum (what the backward kernel needs instead of a)
int small_fprop(const float *x, const float *W, const float *bias, float *a, float *pooled,
                uint8_t *tie, int B, int C, int S, int M, int O, int act, int act_nn, int P,
                cudaStream_t st);

size_t small_bwd_workspace_bytes(int B, int C, int S, int M, int O, int P, bool need_dx);

int small_bwd(const float *x, const float *a, const uint8_t *tie, const float *pooled,
              const float *dtop, const float *W, float *dW, float *db, float *dx,
              const float *below, void *workspace, int B, int C, int S, int M, int O, int act,
              int act_nn, int P, int act_below, int nn_below, cudaStream_t st);

}  // namespace tn
